"""Worker of tests/test_distributed.py: one rank of the sharded join over gloo, with a CPU stand-in
for the device backend (hashing / counting by the oracle). TEST INFRASTRUCTURE."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _cases  # noqa: E402
import _oracle  # noqa: E402
from partitionedhashjoin_b200 import multigpu  # noqa: E402

SEED = 0x9E3779B97F4A7C15


class OracleBackend:
    """Same contract as multigpu.GpuBackend, computed on the CPU by the oracle."""

    def __init__(self, world, oracle):
        self.world, self.oracle, self.launches = world, oracle, 0

    def upload(self, R, S):
        self.rel = [np.ascontiguousarray(R), np.ascontiguousarray(S)]

    def split(self):
        sends, counts = [], []
        for rel in self.rel:
            owner = ((self.oracle.hash_batch(0, SEED, rel["id"]) >> np.uint64(multigpu.SHARD_SHIFT))
                     & np.uint64(self.world - 1)).astype(np.int64)
            order = np.argsort(owner, kind="stable")
            part = rel[order]
            sends.append(torch.from_numpy(part.view("<i8").reshape(-1, 2).copy()))
            counts.append(np.bincount(owner, minlength=self.world))
        return sends, np.stack(counts), 0

    def recv_buffer(self, which, rows):
        return torch.empty((rows, 2), dtype=torch.int64)

    def local_join(self, recv_R, recv_S):
        R = recv_R.numpy().copy().view(_cases.TUPLE).reshape(-1)
        S = recv_S.numpy().copy().view(_cases.TUPLE).reshape(-1)
        self.received = (R, S)
        return self.oracle.count_by_sort(R, S), {"kernel_launches": 0}

    def count_tensor(self, value):
        return torch.tensor([value], dtype=torch.int64)

    def int_tensor(self, array):
        return torch.from_numpy(np.ascontiguousarray(array, dtype=np.int64))

    def close(self):
        pass


def main():
    case = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    oracle = _oracle.Oracle()
    if case == "random":
        keys_r = _cases.splitmix64(40000, 5).astype(np.int64) % 30011
        keys_s = _cases.splitmix64(300000, 6).astype(np.int64) % 45007
    elif case == "skewed":  # one key carries most of the probe side: a single owner receives it
        keys_r = np.arange(1, 20001)
        keys_s = np.where(_cases.splitmix64(200000, 7) % np.uint64(10) < 7, 1234, _cases.splitmix64(200000, 8) % np.uint64(40000)).astype(np.int64)
    else:  # "tiny": fewer tuples than ranks on one side, an empty shard on the last rank
        keys_r = np.array([5, 6, 7])
        keys_s = np.array([7, 7, 8, 5])
    R, S = _cases.tuples(keys_r), _cases.tuples(keys_s)
    want = oracle.count_by_sort(R, S)

    def shard(rel):
        n = rel.shape[0]
        per = n // world
        lo, hi = rank * per, (n if rank == world - 1 else (rank + 1) * per)
        if case == "tiny" and rank == world - 1:
            lo = hi = n if world > 1 else lo  # empty shard
            return rel[n:n]
        if case == "tiny" and rank == world - 2:
            hi = n
        return rel[lo:hi]

    backend = OracleBackend(world, oracle)
    job = multigpu.ShardedRadixJoin(dist if world > 1 else None, rank, world, backend)
    job.upload(shard(R), shard(S))
    res = job.join()
    # every received tuple belongs to this rank, and nothing was lost or duplicated
    got_R, got_S = backend.received
    for rel in (got_R, got_S):
        if rel.shape[0]:
            owner = (oracle.hash_batch(0, SEED, rel["id"]) >> np.uint64(multigpu.SHARD_SHIFT)) & np.uint64(world - 1)
            assert (owner == rank).all()
    rows = torch.tensor([got_R.shape[0], got_S.shape[0]], dtype=torch.int64)
    if world > 1:
        dist.all_reduce(rows)
    assert rows.tolist() == [R.shape[0], S.shape[0]], rows.tolist()
    assert res["matches"] == want, (res["matches"], want)
    res2 = job.join()  # the job is reusable
    assert res2["matches"] == want
    if rank == 0:
        print(json.dumps({"case": case, "world": world, "matches": res["matches"], "want": want}))
    if world > 1:
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
