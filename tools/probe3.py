import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PHJ_KERNEL_TIMES"] = "1"
import partitionedhashjoin_b200 as phj
nr, ns = 10_000_000, 200_000_000
R = np.empty(nr, dtype=phj.TUPLE_DTYPE); S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1)
for alpha in (0.01, 1.05):
    phj.fill_zipf(S, alpha, 1, nr, 12345, 64)
    for algo, P, bits in [("no-partitioning", 0, (0, 0)), ("radix-partitioning", 16, (4, 0)), ("radix-partitioning", 64, (6, 0)), ("radix-partitioning", 256, (8, 0)), ("radix-partitioning", 4096, (6, 6))]:
        with phj.Engine(algo, partitions=P, radix_bits=bits) as e:
            e.upload(R, S)
            best = None
            for _ in range(4):
                res = e.join()
                if best is None or res["total_ns"] < best["total_ns"]:
                    best, kt = res, e.kernel_times()
            assert best["matches"] == ns
            print(f"alpha={alpha} {algo} P={P} {bits}: total {best['total_ns']/1e6:.3f} ms ({(nr+ns)/best['total_ns']:.1f} G/s) fallback={best['fallback_partitions']} | " + " ".join(f"{n}={t/1e3:.0f}" for n, t in kt if t > 15000), flush=True)
