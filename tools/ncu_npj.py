"""Profiling driver for the no-partitioning join (gt_clear / gt_build / gt_probe): 10 M x 200 M keys
generated on the device, a few joins, nothing else -- short enough to run under `ncu --set full`.

    python tools/ncu_npj.py [alpha] [joins]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PHJ_KERNEL_TIMES", "1")
import partitionedhashjoin_b200 as phj  # noqa: E402


def main():
    alpha = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
    joins = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    nr, ns = 10_000_000, 200_000_000
    dR = phj.DeviceTuples(nr, 0).fill_sequential(1)
    dS = phj.DeviceTuples(ns, 0).fill_zipf(alpha, 1, nr, 12345, 1 << 14)
    flags = phj.FLAG_CHAINED_TABLE if os.environ.get("CHAINED") == "1" else 0
    flags |= phj.FLAG_COOP_PROBE if os.environ.get("COOP") == "1" else 0
    with phj.Engine("no-partitioning", flags=flags) as e:
        e.bind_device(dR.ptr, nr, dS.ptr, ns, keepalive=(dR, dS))
        for _ in range(joins):
            res = e.join()
            assert res["matches"] == ns, res
            print(f"npj alpha={alpha}: {res['total_ns'] / 1e6:.3f} ms build {res['build_ns'] / 1e6:.3f} "
                  f"probe {res['probe_ns'] / 1e6:.3f}", e.kernel_times(), flush=True)


if __name__ == "__main__":
    main()
