"""Ad-hoc GPU probe (not part of the tests): per-kernel times of the joins at full size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PHJ_KERNEL_TIMES", "1")
import partitionedhashjoin_b200 as phj  # noqa: E402


def main():
    nr = int(os.environ.get("NR", 10_000_000))
    ns = int(os.environ.get("NS", 200_000_000))
    alpha = float(os.environ.get("ALPHA", 1.05))
    print(json.dumps(phj.device_info(0)))
    t = time.time()
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, alpha, 1, nr, 12345, 64)
    print(f"generated in {time.time() - t:.1f}s", flush=True)
    configs = [("no-partitioning", {}),
               ("radix-partitioning", {"partitions": 4096}),
               ("radix-partitioning", {"partitions": 4096, "flags": phj.FLAG_NO_TMA_STORE}),
               ("radix-partitioning", {"partitions": 2048}),
               ("radix-partitioning", {"partitions": 1024, "radix_bits": (5, 5)}),
               ("radix-partitioning", {"partitions": 8192, "radix_bits": (7, 6)}),
               ("radix-partitioning", {"partitions": 65536, "radix_bits": (8, 8)}),
               ("radix-partitioning", {"partitions": 256, "radix_bits": (8, 0)}),
               ]
    if os.environ.get("PROBE") == "fanout":
        # one pass of 1..8 bits: how the scatter's time depends on the length of the runs it writes
        # (tile of 4096 tuples / digits); the join behind it does not matter here
        configs = [("radix-partitioning", {"partitions": 1 << b, "radix_bits": (b, 0)}) for b in (1, 2, 3, 4, 5, 6, 7, 8)]
    if os.environ.get("PROBE") == "l2join":
        # one partitioning pass + L2-resident partition tables (PHJ_FLAG_L2_TABLES) against the default two
        # passes + shared-memory tables
        configs = [("radix-partitioning", {"partitions": 1 << b, "radix_bits": (b, 0), "flags": phj.FLAG_L2_TABLES})
                   for b in (3, 4, 5, 6, 7, 8)]
        configs += [("radix-partitioning", {"partitions": 4096, "flags": phj.FLAG_L2_TABLES}),
                    ("radix-partitioning", {"partitions": 4096})]
    for algo, kw in configs:
        with phj.Engine(algo, **kw) as e:
            e.upload(R, S)
            best = None
            for it in range(4):
                res = e.join()
                if best is None or res["total_ns"] < best["total_ns"]:
                    best, kt = res, e.kernel_times()
            assert best["matches"] == ns, best
            tput = (nr + ns) / (best["total_ns"] / 1e9) / 1e9
            print(f"== {algo} {kw}: total {best['total_ns'] / 1e6:.3f} ms  part {best['partition_ns'] / 1e6:.3f}  "
                  f"join {best['join_ns'] / 1e6:.3f}  -> {tput:.2f} G tuples/s  fallback={best['fallback_partitions']}")
            for name, ns_ in kt:
                print(f"     {name:24s} {ns_ / 1e3:10.1f} us")
            sys.stdout.flush()


if __name__ == "__main__":
    main()
