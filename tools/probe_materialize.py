"""Full-size timing of the materialising join (10 M x 200 M, 4096 partitions): not part of the tests."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PHJ_KERNEL_TIMES"] = "1"
import partitionedhashjoin_b200 as phj  # noqa: E402

nr, ns = 10_000_000, 200_000_000
R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1)
for alpha in (0.01, 1.25):
    phj.fill_zipf(S, alpha, 1, nr, 12345, 64)
    with phj.Engine("radix-partitioning", partitions=4096) as e:
        e.upload(R, S)
        for _ in range(3):
            res = e.join_materialize()
        assert res["joined_tuples"] == ns and res["matches"] == ns
        head = e.read_joined(0, 1_000_000)
        assert (head["payloadA"] == head["id"] - 1).all()           # build payload = index = id - 1
        assert (S["id"][head["payloadB"]] == head["id"]).all()      # probe payload = index into S
        print(f"alpha={alpha}: total {res['total_ns'] / 1e6:.3f} ms, materialize {res['materialize_ns'] / 1e6:.3f} ms, "
              f"{(nr + ns) / res['total_ns']:.1f} G tuples/s | " +
              " ".join(f"{n}={t / 1e3:.0f}" for n, t in e.kernel_times() if t > 15000), flush=True)
