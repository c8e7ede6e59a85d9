import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import partitionedhashjoin_b200 as phj
import _oracle
o = _oracle.Oracle()
which = sys.argv[1]
n_r, n_s = int(sys.argv[2]), int(sys.argv[3])
R = np.empty(n_r, dtype=phj.TUPLE_DTYPE); S = np.empty(n_s, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1); phj.fill_zipf(S, 1.05, 1, 2 * n_r, 12345, 16)
want = o.count_by_sort(R, S)
P, b1, b2, flags = [int(x) for x in which.split(",")]
with phj.Engine("radix-partitioning", partitions=P, radix_bits=(b1, b2), flags=flags) as e:
    e.upload(R, S); res = e.join()
    print(which, n_r, n_s, "matches", res["matches"], "want", want, "fallback", res["fallback_partitions"], flush=True)
