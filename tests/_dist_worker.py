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


class OracleNpjBackend:
    """CPU stand-in for multigpu.NpjGpuBackend: the gathered build relation is joined with the local
    probe shard by the oracle's no-partitioning join."""

    def __init__(self, world, oracle):
        self.world, self.oracle, self.launches = world, oracle, 0

    def upload(self, R, S):
        self.R = torch.from_numpy(np.ascontiguousarray(R).view("<i8").reshape(-1, 2).copy())
        self.S = np.ascontiguousarray(S)

    def build_shard(self):
        return self.R

    def gather_buffer(self, rows):
        return torch.full((rows, 2), -1, dtype=torch.int64)

    def local_join(self, full_R):
        R = full_R.numpy().copy().view(_cases.TUPLE).reshape(-1)
        self.gathered = R
        count = self.oracle.join_npj(R, self.S) if R.shape[0] else 0  # the reference's table rejects n = 0
        assert count == self.oracle.count_by_sort(R, self.S)
        return count, {"kernel_launches": 0}

    def count_tensor(self, value):
        return torch.tensor([value], dtype=torch.int64)

    def int_tensor(self, array):
        return torch.from_numpy(np.ascontiguousarray(array, dtype=np.int64))

    def close(self):
        pass


class OracleFusedBackend(OracleBackend):
    """CPU stand-in for multigpu.FusedGpuBackend: the receive windows are POSIX shared memory that
    the peers map by name (the role CUDA IPC plays on the device) and write their pieces into."""

    def __init__(self, world, oracle, partitions_local, pass1_in_shuffle):
        super().__init__(world, oracle)
        self.win, self.peer = [None] * 3, [[None] * world for _ in range(3)]
        self.allocs = 0
        self.b1, self.b2, self.ndig = multigpu.split_plan(world, partitions_local, pass1_in_shuffle)
        self.d1 = self.ndig // world

    def digit(self, ids):
        h = self.oracle.hash_batch(0, SEED, ids)
        shift = self.b2 if self.b1 else multigpu.SHARD_SHIFT
        return ((h >> np.uint64(shift)) & np.uint64(self.ndig - 1)).astype(np.int64)

    def win_alloc(self, which, rows):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=max(rows, 1) * 16)
        self.win[which] = shm
        self.allocs += 1
        return shm.name.encode()

    def win_free(self, which):
        if self.win[which] is not None:
            self.win[which].close()
            self.win[which].unlink()
            self.win[which] = None

    def peer_open(self, which, src, handle, own):
        from multiprocessing import shared_memory
        self.peer[which][src] = self.win[which] if own else shared_memory.SharedMemory(name=handle.decode())

    def peer_close(self, which, src, own):
        if self.peer[which][src] is not None and not own:
            self.peer[which][src].close()
        self.peer[which][src] = None

    def count(self):
        self.pieces = []
        counts = []
        for rel in self.rel:
            dig = self.digit(rel["id"])
            self.pieces.append([rel[dig == d] for d in range(self.ndig)])
            counts.append(np.bincount(dig, minlength=self.ndig))
        return np.stack(counts).astype(np.int64)

    def scatter(self, offsets, owner_of):
        owner_of = np.asarray(owner_of)
        if owner_of.ndim == 1:
            owner_of = np.stack([owner_of, owner_of])
        self.owner_of = owner_of[0]
        for which in (0, 1):
            for d in range(self.ndig):
                piece = self.pieces[which][d]
                if piece.shape[0]:
                    shm = self.peer[which][int(owner_of[which][d])]
                    win = np.ndarray((shm.size // 16,), dtype=_cases.TUPLE, buffer=shm.buf)
                    o = int(offsets[which][d])
                    win[o:o + piece.shape[0]] = piece
                    del win
        return 0

    def pull(self, which, owner, src_row, dst_row, rows):
        if rows:
            src = np.ndarray((self.peer[which][owner].size // 16,), dtype=_cases.TUPLE, buffer=self.peer[which][owner].buf)
            dst = np.ndarray((self.win[which].size // 16,), dtype=_cases.TUPLE, buffer=self.win[which].buf)
            dst[dst_row:dst_row + rows] = src[src_row:src_row + rows]
            del src, dst
            self.pulled = getattr(self, "pulled", 0) + rows

    def local_join_window(self, rows, bounds, first_digit=0, parent_digits=None, windows=(0, 1)):
        got = []
        for which in (0, 1):
            n = rows[which]
            shm = self.win[windows[which]]
            win = np.ndarray((n,), dtype=_cases.TUPLE, buffer=shm.buf) if n else np.empty(0, _cases.TUPLE)
            got.append(win.copy())
            del win
            # the window is partitioned by split digit exactly as `bounds` says, starting at first_digit
            b = np.asarray(bounds[which])
            assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
            if n:
                digits = first_digit + np.arange(b.shape[0] - 1) if parent_digits is None else np.asarray(parent_digits)
                assert (self.digit(got[which]["id"]) == np.repeat(digits, np.diff(b))).all()
        self.received = tuple(got)
        return self.oracle.count_by_sort(*got), {"kernel_launches": 0}


def main():
    case = sys.argv[1]
    fused = len(sys.argv) > 2 and sys.argv[2] in ("fused", "pass1")
    pass1 = fused and sys.argv[2] == "pass1"
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

    if len(sys.argv) > 2 and sys.argv[2] == "npj":
        # no-partitioning join: the build shards are gathered on every rank, the probe shards stay
        job = multigpu.ReplicatedNoPartitioningJoin(dist if world > 1 else None, rank, world, OracleNpjBackend(world, oracle))
        job.upload(shard(R), shard(S))
        for _ in range(2):  # reusable
            res = job.join()
            assert res["matches"] == want, (res["matches"], want)
            got = job.backend.gathered
            assert got.shape[0] == R.shape[0] == res["build_rows"]
            assert (got["id"] == R["id"]).all() and (got["payload"] == R["payload"]).all()  # shards in rank order
        if rank == 0:
            print(json.dumps({"case": case, "world": world, "matches": res["matches"], "want": want, "fused": False}))
        if world > 1:
            dist.barrier()
        dist.destroy_process_group()
        return
    if fused:
        backend = OracleFusedBackend(world, oracle, 256, pass1)
        assert (backend.b1 > 0) == pass1
        job = multigpu.FusedShardedRadixJoin(dist if world > 1 else None, rank, world, backend)
    else:
        backend = OracleBackend(world, oracle)
        job = multigpu.ShardedRadixJoin(dist if world > 1 else None, rank, world, backend)
    job.upload(shard(R), shard(S))
    res = job.join()
    # every received tuple belongs to this rank, and nothing was lost or duplicated
    got_R, got_S = backend.received
    for rel in (got_R, got_S):
        if rel.shape[0]:
            if fused and res.get("hot_digits"):
                continue  # heavy-hitter digits are replicated (build) / stay where they are (probe)
            if fused:
                owner = backend.owner_of[backend.digit(rel["id"])]
            else:
                owner = (oracle.hash_batch(0, SEED, rel["id"]) >> np.uint64(multigpu.SHARD_SHIFT)) & np.uint64(world - 1)
            assert (owner == rank).all()
    rows = torch.tensor([got_R.shape[0] - getattr(backend, "pulled", 0), got_S.shape[0]], dtype=torch.int64)
    if world > 1:
        dist.all_reduce(rows)
    assert rows.tolist() == [R.shape[0], S.shape[0]], rows.tolist()  # nothing lost; only replicas added
    if case == "skewed" and pass1 and world > 1:
        assert res["hot_digits"], "the 70 % key must be recognised as a heavy hitter"
    assert res["matches"] == want, (res["matches"], want)
    res2 = job.join()  # the job is reusable
    assert res2["matches"] == want
    if fused:
        assert res["regrown"] and not res2["regrown"] and backend.allocs <= 3  # windows are kept
        # a bigger probe shard: the windows grow collectively, the count follows
        job.upload(shard(R), np.concatenate([shard(S)] * 3))
        res3 = job.join()
        assert res3["matches"] == 3 * want
        job.upload(shard(R), shard(S))
        assert job.join()["matches"] == want
    if rank == 0:
        print(json.dumps({"case": case, "world": world, "matches": res["matches"], "want": want, "fused": fused}))
    if fused:
        job.close()
    if world > 1:
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
