"""Worker of the gpu-marked multi-rank tests: one rank per GPU over NCCL, the product backends
(multigpu.GpuBackend / FusedGpuBackend) on the device, the oracle only as the checker.
TEST INFRASTRUCTURE."""
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


def check_window(job, oracle, rel_global, which, rank, world, hash_seed=0x9E3779B97F4A7C15, hash_id=0):
    """What landed in this rank's windows, tuple by tuple. This rank's partitions are the digits it owns, then the hot digits of other owners. A build
    partition holds ALL tuples of its digit (a foreign hot one was copied from its owner); a probe partition of a hot
    digit holds this rank's OWN tuples only (they never travelled), any other one everybody's."""
    lay = job.info()
    got, bounds = job.read_window(which)
    dl, ndig, hot = lay["local_partitions"], lay["digits"], lay["hot_digits"]
    parts = list(range(rank * dl, rank * dl + dl)) + [h for h in hot if h // dl != rank]
    assert lay["partitions_here"] == len(parts)
    def digits_of(keys):  # a bit field of the hash for 2 / 4 / 8 ... ranks, hash % digits for any other count
        h = oracle.hash_batch(hash_id, hash_seed, keys)
        return (h % np.uint64(ndig) if ndig & (ndig - 1) else h & np.uint64(ndig - 1)).astype(np.int64)

    digit = digits_of(rel_global["id"])
    per = rel_global.shape[0] // world
    in_my_shard = np.zeros(rel_global.shape[0], dtype=bool)
    in_my_shard[rank * per:(rel_global.shape[0] if rank == world - 1 else (rank + 1) * per)] = True
    regions = bounds[None] if which == 0 else bounds
    for idx, d in enumerate(parts):
        pieces = []
        for c in range(regions.shape[0]):
            piece = got[regions[c][idx]:regions[c][idx + 1]]
            dg = digits_of(piece["id"])
            assert (dg == d).all(), (which, c, idx, d)
            assert (np.diff(piece["payload"]) > 0).all(), (which, c, idx, "order inside a region")
            pieces.append(piece)
        allp = np.concatenate(pieces)
        allp = allp[np.argsort(allp["payload"], kind="stable")]
        sel = digit == d
        if which == 1 and d in hot:
            sel &= in_my_shard
        expect = rel_global[sel]
        assert allp.shape[0] == expect.shape[0], (which, idx, d, allp.shape, expect.shape)
        assert (allp["id"] == expect["id"]).all() and (allp["payload"] == expect["payload"]).all()


def check_library_join(rank, world, local, oracle, R, S, want, shard, mode):
    dist_mod = dist if world > 1 else None
    # both count modes at every world size: piece-wise (later probe chunks counted while the earlier ones travel)
    # is the default from 4 GPUs on, up front below that
    from partitionedhashjoin_b200 import _lib
    for partitions, chunks, flags in ((0, 0, 0), (world * 4, 3, _lib.FLAG_COUNT_PIECEWISE), (256 // world * world, 1, 0),
                                      (world, 5, _lib.FLAG_COUNT_UPFRONT), (0, 4, _lib.FLAG_COUNT_PIECEWISE)):
        if partitions and partitions < world:
            continue
        job = multigpu.ShardedJoin(dist_mod, rank, world, local, partitions=partitions, chunks=chunks, flags=flags)
        job.upload(shard(R), shard(S))
        for _ in range(3):
            res = job.join()
            assert res["matches"] == want, (res["matches"], want, partitions, chunks)
        assert job.info()["resizes"] == 1, job.info()
        check_window(job, oracle, R, 0, rank, world)
        check_window(job, oracle, S, 1, rank, world)
        # bigger shards: the device-side overflow flag sends every rank through a collective re-size
        job.upload(shard(R), np.concatenate([shard(S)] * 2))
        res = job.join()
        assert res["matches"] == 2 * want, (res["matches"], 2 * want)
        res = job.join()
        assert res["matches"] == 2 * want and res["gpus"] == world
        # smaller again: the windows are kept
        job.upload(shard(R), shard(S)[: shard(S).shape[0] // 2])
        n_resizes = job.info()["resizes"]
        r1 = job.join()
        part = np.concatenate([shard_of(S, q, world)[: shard_of(S, q, world).shape[0] // 2] for q in range(world)])
        assert r1["matches"] == oracle.count_by_sort(R, part) and job.info()["resizes"] == n_resizes
        job.close()
    # heavy hitters (30 % of the probe keys are one key) are kept local by default: the runs above went that way.
    # Against the plain exchange (PHJ_FLAG_NO_HOT_DIGITS): same count, nothing marked hot, more bytes over NVLink
    job = multigpu.ShardedJoin(dist_mod, rank, world, local, chunks=3)
    job.upload(shard(R), shard(S))
    assert job.join()["matches"] == want
    plain = multigpu.ShardedJoin(dist_mod, rank, world, local, chunks=3, flags=_lib.FLAG_NO_HOT_DIGITS)
    plain.upload(shard(R), shard(S))
    assert plain.join()["matches"] == want
    if world > 1:
        assert job.info()["hot_digits"] and not plain.info()["hot_digits"], (job.info(), plain.info())
        assert job.info()["sent_remote_bytes"] <= plain.info()["sent_remote_bytes"]
    check_window(plain, oracle, R, 0, rank, world)
    check_window(plain, oracle, S, 1, rank, world)
    job.close()
    plain.close()
    if rank == 0:
        print(json.dumps({"mode": mode, "world": world, "matches": res["matches"] // 2, "want": want}))


def shard_of(rel, q, world):
    per = rel.shape[0] // world
    return rel[q * per:(rel.shape[0] if q == world - 1 else (q + 1) * per)]


def main():
    mode = sys.argv[1]  # lib | pass1 | fused | nccl | npj
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    oracle = _oracle.Oracle()
    # same global relations on every rank; each rank uploads its row range
    keys_r = _cases.splitmix64(400_000, 21).astype(np.int64) % 300_007
    keys_s = np.where(_cases.splitmix64(3_000_000, 22) % np.uint64(10) < 3, 4242,
                      _cases.splitmix64(3_000_000, 23) % np.uint64(450_001)).astype(np.int64)
    R, S = _cases.tuples(keys_r), _cases.tuples(keys_s)
    want = oracle.count_by_sort(R, S)

    def shard(rel, scale=1):
        per = rel.shape[0] // world
        lo, hi = rank * per, (rel.shape[0] if rank == world - 1 else (rank + 1) * per)
        return rel[lo:hi]

    if mode.startswith("lib"):  # the sharded join inside libphj_b200.so (phj_dist_*): ShardedJoin
        check_library_join(rank, world, local, oracle, R, S, want, shard, mode)
        dist.barrier()
        dist.destroy_process_group()
        return
    fused = mode in ("pass1", "fused")
    if mode == "npj":  # no-partitioning join, build side gathered on every rank (not yet in the pytest matrix)
        job = multigpu.ReplicatedNoPartitioningJoin(dist, rank, world, multigpu.NpjGpuBackend(world, local))
        job.upload(shard(R), shard(S))
        for _ in range(2):
            res = job.join()
            assert res["matches"] == want and res["build_rows"] == R.shape[0], (res, want)
        if rank == 0:
            print(json.dumps({"mode": mode, "world": world, "matches": res["matches"], "want": want}))
        job.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if fused:
        job = multigpu.FusedShardedRadixJoin(dist, rank, world, multigpu.FusedGpuBackend(
            world, local, partitions_local=256, pass1_in_shuffle=(mode == "pass1")))
    else:
        job = multigpu.ShardedRadixJoin(dist, rank, world, multigpu.GpuBackend(world, local, partitions_local=256))
    job.upload(shard(R), shard(S))
    res = job.join()
    assert res["matches"] == want, (res["matches"], want)
    res = job.join()
    assert res["matches"] == want, (res["matches"], want)
    first = res["matches"]
    if fused:
        assert not res["regrown"]
        # what landed in this rank's window is exactly the tuples it owns, pieces in source-rank order
        # with input order inside a piece (stable split)
        be = job.backend
        for which, rel in ((0, R), (1, S)):
            rows = res["recv_rows"][which]
            got = torch.as_tensor(multigpu._CudaView(be.win[which], rows), device=f"cuda:{local}").cpu().numpy() \
                if rows else np.empty((0, 2), np.int64)
            shift = be.b2 if be.b1 else multigpu.SHARD_SHIFT
            digit = ((oracle.hash_batch(0, 0x9E3779B97F4A7C15, rel["id"]) >> np.uint64(shift))
                     & np.uint64(be.ndig - 1)).astype(np.int64)
            lo, hi = res["first_digit"][rank], res["first_digit"][rank + 1]  # the digits this rank owns
            parents = res["parent_digits"] if res["parent_digits"] is not None else list(range(lo, hi))
            hot = set(res["hot_digits"])
            per = rel.shape[0] // world
            in_my_shard = np.zeros(rel.shape[0], dtype=bool)
            in_my_shard[rank * per:(rel.shape[0] if rank == world - 1 else (rank + 1) * per)] = True
            # parent by parent: the tuples of that digit in global input order (= source-rank order,
            # then input order); of a heavy-hitter digit the probe side holds this rank's own piece only
            pieces = []
            for d in parents:
                sel = digit == d
                if which == 1 and d in hot:
                    sel &= in_my_shard
                pieces.append(rel[sel])
            expect = np.concatenate(pieces) if pieces else rel[:0]
            assert np.diff(res["bounds"][which]).tolist() == [p.shape[0] for p in pieces]
            assert got.shape[0] == expect.shape[0], (got.shape, expect.shape)
            assert (got[:, 0] == expect["id"]).all() and (got[:, 1] == expect["payload"]).all()
        # bigger shards: windows regrow collectively
        job.upload(shard(R), np.concatenate([shard(S)] * 2))
        res = job.join()
        assert res["matches"] == 2 * want and res["regrown"]
    if rank == 0:
        print(json.dumps({"mode": mode, "world": world, "matches": first, "want": want}))
    job.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
