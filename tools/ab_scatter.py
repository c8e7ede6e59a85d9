"""A/B of radix_scatter builds on the GPU box (not part of the tests).

  python tools/ab_scatter.py build [names...]   # here: builds partitionedhashjoin_b200/variants/libphj_<name>.so
  python tools/ab_scatter.py [names...]         # on the box: the two-pass 10 M x 200 M join with every variant

Every variant is the same source with other compile-time knobs (-DPHJ_SCAT_*). A child process per variant
(PHJ_LIB selects the library) joins device-generated relations 8 times and prints the median step and the
per-kernel device times; `check` in the name list also runs the partition-layout parity tests with that build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARDIR = os.path.join(ROOT, "partitionedhashjoin_b200", "variants")
VARIANTS = {
    "peers6": "-DPHJ_SCAT_PEERS4=0",                        # warp_peers as plain C++: six ALU instructions per digit bit
    "base": "",                                             # the default build: 16 segments per SM, leader-free store
    "segs8": "-DPHJ_SEGS_PER_SM=8",                         # round 1's segment length
    "segs24": "-DPHJ_SEGS_PER_SM=24",
    "segs32": "-DPHJ_SEGS_PER_SM=32",
    "leader": "-DPHJ_SCAT_ALLWRITE=0",                      # elected leader stores the warp counter
    "joinw8": "-DPHJ_JOIN_WAVES=8",                         # shorter probe slices in join_partitions
    "joinw12": "-DPHJ_JOIN_WAVES=12",
    "pscan0": "-DPHJ_SCAT_PSCAN=0",                         # round 1's two-warp counter scan, four barriers per tile
    "t1024x4": "-DPHJ_SCAT_TPB=1024 -DPHJ_SCAT_IPT=4",      # same 4096-tuple tile, 64 warps per SM (32 registers)
    "t768x5": "-DPHJ_SCAT_TPB=768 -DPHJ_SCAT_IPT=5",        # 3840-tuple tile, 48 warps per SM
}


def build(only):
    os.makedirs(VARDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        if only and name not in only:
            continue
        out = os.path.join(VARDIR, f"libphj_{name}.so")
        cmd = (f"nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-pthread {flags} "
               f"-shared -o {out} {ROOT}/partitionedhashjoin_b200/csrc/phj_engine.cu {ROOT}/partitionedhashjoin_b200/csrc/phj_datagen.cpp")
        procs.append((name, subprocess.Popen(cmd, shell=True)))
    for name, p in procs:
        assert p.wait() == 0, name


def child():
    sys.path.insert(0, ROOT)
    os.environ["PHJ_KERNEL_TIMES"] = "1"
    import partitionedhashjoin_b200 as phj
    nr, ns = 10_000_000, 200_000_000
    for alpha in [float(a) for a in os.environ.get("ALPHAS", "0.01").split(",")]:
        dR = phj.DeviceTuples(nr, 0).fill_sequential(1)
        dS = phj.DeviceTuples(ns, 0).fill_zipf(alpha, 1, nr, 12345, 1 << 14)
        for parts, bits in ((4096, (6, 6)), (64, (6, 0))):
            with phj.Engine("radix-partitioning", partitions=parts, radix_bits=bits) as e:
                e.bind_device(dR.ptr, nr, dS.ptr, ns, keepalive=(dR, dS))
                runs = []
                for _ in range(8):
                    res = e.join()
                    assert res["matches"] == ns, res
                    runs.append((res["total_ns"], e.kernel_times()))
                runs.sort(key=lambda r: r[0])
                med, kt = runs[len(runs) // 2]
                print(f"  alpha={alpha} P={parts}: median {med / 1e6:.3f} ms best {runs[0][0] / 1e6:.3f} | " +
                      " ".join(f"{n.replace('radix_', '')}={t / 1e3:.0f}" for n, t in kt if t > 5000), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        only = [a for a in sys.argv[1:] if a != "check"]
        for name in VARIANTS:
            if only and name not in only:
                continue
            lib = os.path.join(VARDIR, f"libphj_{name}.so")
            if not os.path.exists(lib):
                continue
            print(f"== {name} ({VARIANTS[name] or 'default'})", flush=True)
            env = dict(os.environ, PHJ_LIB=lib)
            subprocess.run([sys.executable, __file__, "child"], env=env)
            if "check" in sys.argv[1:] and "pscan0" not in name:
                r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                                    "-m", "gpu", "-k", "partition_layout or prepartitioned or golden"], env=env,
                                   capture_output=True, text=True)
                print("  parity:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
