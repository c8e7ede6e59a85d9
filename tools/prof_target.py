"""Profiling target: one radix join configuration, a few joins (for ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import partitionedhashjoin_b200 as phj
nr, ns = int(os.environ.get("NR", 10_000_000)), int(os.environ.get("NS", 200_000_000))
algo = os.environ.get("ALGO", "radix-partitioning")
P = int(os.environ.get("P", 4096))
R = np.empty(nr, dtype=phj.TUPLE_DTYPE); S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1); phj.fill_zipf(S, float(os.environ.get("ALPHA", 1.05)), 1, nr, 12345, 64)
kw = {"partitions": P} if algo != "no-partitioning" else {}
with phj.Engine(algo, **kw) as e:
    e.upload(R, S)
    for _ in range(int(os.environ.get("JOINS", 3))):
        res = e.join()
    assert res["matches"] == ns
    print(res)
