"""Kernel-shape tuning on the GPU box (not part of the tests).

  python tools/tune_shapes.py build   # here: builds partitionedhashjoin_b200/variants/*.so (they travel with gpurun)
  python tools/tune_shapes.py         # on the box: every variant x {separate, fused} pass-2 histogram

Each variant is the same source compiled with other tile shapes (-DPHJ_SCAT_TPB=.. etc.)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARDIR = os.path.join(ROOT, "partitionedhashjoin_b200", "variants")
VARIANTS = {
    "s512x8m2": "",
    "s512x8m2_plain3": "-DPHJ_SCAT_MINB_PLAIN=3",  # three CTAs/SM for the scatter without the fused histogram
                                                   # (40 registers: ~20 spilled; 71 KB of shared memory each)
    "s384x8m3": "-DPHJ_SCAT_TPB=384 -DPHJ_SCAT_MINB=3",  # 56 registers, no spills, three CTAs/SM (also fused)
    "s256x8m5": "-DPHJ_SCAT_TPB=256 -DPHJ_SCAT_MINB=5",
    "s512x6m2": "-DPHJ_SCAT_IPT=6",
    "s512x4m3": "-DPHJ_SCAT_IPT=4 -DPHJ_SCAT_MINB=3",
    "s256x8m4": "-DPHJ_SCAT_TPB=256 -DPHJ_SCAT_MINB=4",
    "s384x8m2": "-DPHJ_SCAT_TPB=384 -DPHJ_SCAT_MINB=2",
    "s512x8m2_ju6": "-DPHJ_JOIN_U=6",   # more probe loads in flight per thread (ncu: 11.5 % of the join's
    "s512x8m2_ju8": "-DPHJ_JOIN_U=8",   # stall samples wait on the first use of a loaded key)
    "s512x8m2_j256": "-DPHJ_JOIN_TPB=256",
    "s512x8m2_j1024": "-DPHJ_JOIN_TPB=1024",
    "s512x8m2_h1024": "-DPHJ_HIST_TPB=1024 -DPHJ_HIST_IPT=4",
}


def build():
    os.makedirs(VARDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(VARDIR, f"libphj_{name}.so")
        cmd = (f"nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-pthread {flags} "
               f"-shared -o {out} {ROOT}/partitionedhashjoin_b200/csrc/phj_engine.cu {ROOT}/partitionedhashjoin_b200/csrc/phj_datagen.cpp")
        procs.append((name, subprocess.Popen(cmd, shell=True)))
        if len(procs) % 4 == 0:
            for _, p in procs[-4:]:
                p.wait()
    for name, p in procs:
        assert p.wait() == 0, name


def child():
    import numpy as np
    sys.path.insert(0, ROOT)
    os.environ["PHJ_KERNEL_TIMES"] = "1"
    import partitionedhashjoin_b200 as phj
    nr, ns = 10_000_000, 200_000_000
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    for alpha in [float(a) for a in os.environ.get("ALPHAS", "0.01").split(",")]:
        phj.fill_zipf(S, alpha, 1, nr, 12345, 64)
        for flags in (phj.FLAG_NO_FUSE_HIST2, 0):
            with phj.Engine("radix-partitioning", partitions=4096, radix_bits=(6, 6), flags=flags) as e:
                e.upload(R, S)
                runs = []
                for _ in range(8):
                    res = e.join()
                    runs.append((res["total_ns"], e.kernel_times()))
                assert res["matches"] == ns
                runs.sort(key=lambda r: r[0])
                med, kt = runs[len(runs) // 2]
                print(f"  alpha={alpha} fuse2={int(not flags)}: median {med / 1e6:.3f} ms best {runs[0][0] / 1e6:.3f} | " +
                      " ".join(f"{n.replace('radix_', '')}={t / 1e3:.0f}" for n, t in kt if t > 15000), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    elif len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        only = sys.argv[1:]
        for name in VARIANTS:
            if only and name not in only:
                continue
            print(f"== {name}", flush=True)
            env = dict(os.environ, PHJ_LIB=os.path.join(VARDIR, f"libphj_{name}.so"))
            subprocess.run([sys.executable, __file__, "child"], env=env)
