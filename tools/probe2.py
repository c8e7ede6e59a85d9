"""Variant probe: PHJ_RANK in {default, match, lanes, stable} x a few fan-outs, per-kernel times."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["PHJ_KERNEL_TIMES"] = "1"
    import partitionedhashjoin_b200 as phj
    nr, ns = 10_000_000, 200_000_000
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE); S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1); phj.fill_zipf(S, float(os.environ.get("ALPHA", "1.05")), 1, nr, 12345, 64)
    for P, bits in [(4096, (6, 6)), (2048, (6, 5)), (65536, (8, 8))]:
        with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits) as e:
            e.upload(R, S)
            best = None
            for _ in range(4):
                res = e.join()
                if best is None or res["total_ns"] < best["total_ns"]:
                    best, kt = res, e.kernel_times()
            assert best["matches"] == ns
            print(f"  P={P} {bits}: total {best['total_ns']/1e6:.3f} ms | " + " ".join(f"{n.split('[')[0][6:] if n.startswith('radix_') else n}{n[n.index('['):] if '[' in n else ''}={t/1e3:.0f}" for n, t in kt if t > 15000), flush=True)
else:
    for rank in ["", "match", "lanes", "stable"]:
        env = dict(os.environ)
        if rank: env["PHJ_RANK"] = rank
        print(f"PHJ_RANK={rank or 'default'}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
