import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import partitionedhashjoin_b200 as phj
import _oracle
o = _oracle.Oracle()
n_r, n_s = 200_000, 3_000_000
R = np.empty(n_r, dtype=phj.TUPLE_DTYPE); S = np.empty(n_s, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1); phj.fill_zipf(S, 1.05, 1, 2 * n_r, 12345, 16)
want = o.count_by_sort(R, S)
for P, bits, flags in [(128,(0,0),0),(128,(0,0),2),(128,(4,3),0),(256,(8,0),0),(64,(6,0),0),(128,(7,0),0)]:
    with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, flags=flags) as e:
        e.upload(R, S); res = e.join()
        msg = f"P={P} bits={bits} flags={flags} matches={res['matches']} want={want}"
        for which, rel in ((0, R), (1, S)):
            got, gb = e.read_partitions(which, P)
            w, wb = o.radix_partition(rel, P, 0, 0x9E3779B97F4A7C15, workers=1)
            okb = gb[:-1].tolist() == wb[:, 0].tolist()
            okd = bool((got["id"] == w["id"]).all() and (got["payload"] == w["payload"]).all())
            msg += f" | rel{which} bounds_ok={okb} data_ok={okd}"
            if not okd:
                bad = np.nonzero((got["id"] != w["id"]) | (got["payload"] != w["payload"]))[0]
                msg += f" nbad={bad.size} first={bad[:5].tolist()}"
        # recount per partition on host using the GPU's layout
        print(msg, flush=True)
