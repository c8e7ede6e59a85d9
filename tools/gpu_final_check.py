"""A few seconds of GPU time for the final build of round 2: one single-GPU radix join and the library's sharded join
with ONE rank (dist_plan / NCCL communicator / windows / dist_layout / L2-table probes) against a numpy count, without
importing torch (its import alone can take longer than the GPU time that was left). Prints one JSON line."""
import json
import os
import sys
import time
import types

t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# multigpu imports torch only so that the library binds the NCCL torch has loaded; the system libnccl.so.2 serves too
stub = types.ModuleType("torch")
stub.distributed = types.ModuleType("torch.distributed")
sys.modules.setdefault("torch", stub)
sys.modules.setdefault("torch.distributed", stub.distributed)
import numpy as np  # noqa: E402

import partitionedhashjoin_b200 as phj  # noqa: E402
from partitionedhashjoin_b200 import multigpu  # noqa: E402

out = {"device": phj.device_info(0)["name"], "lib": os.path.basename(phj.LIB_PATH)}
n_r, n_s = int(os.environ.get("N_R", 1_000_000)), int(os.environ.get("N_S", 20_000_000))
R = np.empty(n_r, dtype=phj.TUPLE_DTYPE)
S = np.empty(n_s, dtype=phj.TUPLE_DTYPE)
phj.fill_sequential(R, 1)
phj.fill_zipf(S, 1.25, 1, 2 * n_r, 4242, 16)  # half of the key range misses R, key 1 is a heavy hitter
S["id"][::max(1, n_s // 20)] = -(2 ** 63)
want = int(((S["id"] >= 1) & (S["id"] <= n_r)).sum())
with phj.Engine("radix-partitioning", partitions=4096) as e:
    e.upload(R, S)
    res = e.join()
out["radix_2pass"] = {"matches": res["matches"], "want": want, "ms": res["total_ns"] / 1e6}
job = multigpu.ShardedJoin(None, 0, 1, 0)
job.upload(R, S)
res = [job.join() for _ in range(3)][-1]
out["sharded_1_rank"] = {"matches": res["matches"], "want": want, "ms": res["total_ns"] / 1e6, "layout": job.info()}
job.close()
out["ok"] = out["radix_2pass"]["matches"] == want and out["sharded_1_rank"]["matches"] == want
out["seconds"] = round(time.time() - t0, 1)
print(json.dumps(out))
sys.exit(0 if out["ok"] else 1)
