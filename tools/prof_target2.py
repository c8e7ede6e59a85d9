import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import partitionedhashjoin_b200 as phj
nr, ns = 10_000_000, 200_000_000
R = np.empty(nr, dtype=phj.TUPLE_DTYPE); S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1); phj.fill_zipf(S, 0.01, 1, nr, 12345, 64)
with phj.Engine("radix-partitioning", partitions=64, radix_bits=(6, 0)) as e:
    e.upload(R, S)
    for _ in range(2):
        res = e.join()
    assert res["matches"] == ns
    print(res)
