"""The multi-rank host logic of the sharded join (partitionedhashjoin_b200/multigpu.py) over gloo on
the CPU: split sizes, the size + data all-to-all, the local joins and the count all-reduce, with
the oracle standing in for the device (tests/_dist_worker.py). The same orchestration class runs on
NCCL with GpuBackend in bench.py --gpus N and in the gpu-marked test below."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(world, script, *args, timeout=300):
    for attempt in range(3):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), script, *args]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        # the probed port can be taken again (an NCCL bootstrap socket of the run before) by the time torchrun binds it
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    return r


@pytest.mark.parametrize("world,case", [(2, "random"), (2, "skewed"), (2, "tiny"), (4, "random"), (1, "random")])
def test_sharded_join_over_gloo(world, case):
    r = torchrun(world, os.path.join(HERE, "_dist_worker.py"), case)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0


@pytest.mark.parametrize("world,case,mode", [(2, "random", "pass1"), (2, "skewed", "pass1"), (2, "tiny", "pass1"),
                                             (4, "random", "pass1"), (4, "skewed", "pass1"), (1, "random", "pass1"), (2, "random", "fused")])
def test_fused_shuffle_over_gloo(world, case, mode):
    """FusedShardedRadixJoin: sizes all-gather -> window offsets -> every rank writes its pieces into
    the owners' windows (shared memory stands in for the CUDA-IPC-mapped NVLink windows). pass1:
    the split digit is (owner : local pass-1 digit) and the window arrives pass-1 partitioned."""
    r = torchrun(world, os.path.join(HERE, "_dist_worker.py"), case, mode)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["fused"] and line["matches"] == line["want"] and line["want"] > 0


@pytest.mark.parametrize("world,ndig", [(2, 8), (4, 16), (2, 128), (8, 256), (1, 64)])
def test_window_layout_against_brute_force(world, ndig):
    """FusedShardedRadixJoin.ownership / layout: window sizes, write offsets (digit-major, then source
    rank) and local boundaries for equal and for count-balanced digit ownership, against a direct
    enumeration of where every piece lands."""
    import numpy as np

    from partitionedhashjoin_b200.multigpu import FusedShardedRadixJoin as F
    rng = np.random.default_rng(world * 1000 + ndig)
    M = rng.integers(0, 50, size=(world, 2, ndig))
    M[:, 1, 3] += 2000  # a heavy-hitter digit on the probe side
    for balance in (False, True):
        first = F.ownership(M, world, balance)
        assert first[0] == 0 and first[-1] == ndig and (np.diff(first) >= 0).all()
        if not balance:
            assert (np.diff(first) == ndig // world).all()
        for rank in range(world):
            need, offsets, bounds, owner = F.layout(M, world, rank, first)
            assert (owner == np.repeat(np.arange(world), np.diff(first))).all()
            for rel in (0, 1):
                for o in range(world):
                    pos = 0
                    for d in range(first[o], first[o + 1]):
                        if o == rank:
                            assert bounds[rel][d - first[o]] == pos
                        for src in range(world):
                            if src == rank:
                                assert offsets[rel][d] == pos
                            pos += M[src][rel][d]
                    assert need[rel][o] == pos
                assert bounds[rel][-1] == need[rel][rank]
        if balance and world > 1:  # the cut is no worse than the equal split
            load = lambda f: max(M[:, :, f[o]:f[o + 1]].sum() for o in range(world))
            assert load(first) <= load(F.ownership(M, world, False))


@pytest.mark.parametrize("world,case", [(2, "random"), (2, "skewed"), (2, "tiny"), (4, "random"), (1, "random")])
def test_replicated_no_partitioning_join_over_gloo(world, case):
    """ReplicatedNoPartitioningJoin (NPJ on several GPUs): every rank gathers the whole build
    relation (shards in rank order, uneven and empty shards included), probes it with its own probe
    shard, and the summed counts equal the oracle's."""
    r = torchrun(world, os.path.join(HERE, "_dist_worker.py"), case, "npj")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0


@pytest.mark.parametrize("world,ndig,hot_on", [(2, 8, "all"), (4, 16, "all"), (8, 256, "all"), (4, 128, "some"),
                                               (8, 64, "some"), (2, 128, "none")])
def test_hot_digit_layout_simulation(world, ndig, hot_on):
    """FusedShardedRadixJoin.hot_digits / layout_hot: a numpy simulation of the whole exchange -- every
    rank scatters its pieces to the windows layout_hot names, non-owners pull the replicated build side
    of the heavy-hitter digits -- after which every window is completely filled without overlap, every
    local parent holds the WHOLE build side of its digit (all sources, source order) and, on the probe
    side, all sources' pieces (normal digit, at its owner) or only the rank's own piece (hot digit, on
    every rank): each probe tuple ends up on exactly one rank, next to every build tuple it can match."""
    import numpy as np

    from partitionedhashjoin_b200.multigpu import FusedShardedRadixJoin as F
    rng = np.random.default_rng(7 * world + ndig)
    M = rng.integers(0, 40, size=(world, 2, ndig))
    M[:, 0, 5 % ndig] = 0                                  # a digit without build tuples
    hot_want = []
    if hot_on != "none":
        hot_want = sorted({3 % ndig, (ndig // 2 + 1) % ndig})
        for d in hot_want:
            M[:, 1, d] = rng.integers(3000, 6000, size=world)
        if hot_on == "some":                               # a heavy hitter that only some ranks hold
            M[1:, 1, hot_want[0]] = 0
            M[0, 1, hot_want[0]] = 20000
    hot = F.hot_digits(M, world)
    assert sorted(np.nonzero(hot)[0].tolist()) == hot_want
    Wt = M.copy()
    Wt[:, 1, hot] = 0
    first = F.ownership(Wt, world, True)
    plans = [F.layout_hot(M, world, r, first, hot) for r in range(world)]
    need = plans[0][0]
    for pl in plans:
        assert (pl[0] == need).all()                       # every rank derives the same window sizes

    def ids(rel, src, d):                                  # globally unique tuple ids of one piece
        return ((rel * world + src) * ndig + d) * 100000 + np.arange(M[src][rel][d])

    win = [[np.full(int(need[rel][o]), -1, dtype=np.int64) for o in range(world)] for rel in (0, 1)]
    for src, (_, offsets, dst, _, _, _) in enumerate(plans):
        for rel in (0, 1):
            for d in range(ndig):
                n, o, at = int(M[src][rel][d]), int(dst[rel][d]), int(offsets[rel][d])
                assert (win[rel][o][at:at + n] == -1).all(), "pieces overlap"
                win[rel][o][at:at + n] = ids(rel, src, d)
    for r, (_, _, _, _, _, pulls) in enumerate(plans):     # after the exchange barrier
        for owner, src_row, dst_row, rows in pulls:
            assert owner != r and (win[0][r][dst_row:dst_row + rows] == -1).all()
            win[0][r][dst_row:dst_row + rows] = win[0][owner][src_row:src_row + rows]
    seen_probe = []
    for r, (_, _, _, bounds, parents, _) in enumerate(plans):
        assert (win[0][r] != -1).all() and (win[1][r] != -1).all(), "window not completely filled"
        assert bounds[0][-1] == need[0][r] and bounds[1][-1] == need[1][r]
        owned = set(range(int(first[r]), int(first[r + 1])))
        assert set(parents.tolist()) == owned | set(hot_want)
        for i, d in enumerate(parents.tolist()):
            build = win[0][r][bounds[0][i]:bounds[0][i + 1]]
            assert (build == np.concatenate([ids(0, src, d) for src in range(world)])).all()
            probe = win[1][r][bounds[1][i]:bounds[1][i + 1]]
            want = ids(1, r, d) if hot[d] else np.concatenate([ids(1, src, d) for src in range(world)])
            assert (probe == want).all()
            seen_probe.append(probe)
    seen = np.sort(np.concatenate(seen_probe))
    everything = np.sort(np.concatenate([ids(1, src, d) for src in range(world) for d in range(ndig)]))
    assert (seen == everything).all()                      # every probe tuple on exactly one rank


def test_split_plan():
    """The split digit (owner rank : local pass-1 digit) never exceeds 256 values and the local plan
    keeps passes of <= 8 bits; otherwise the split falls back to owner-only."""
    from partitionedhashjoin_b200.multigpu import split_plan
    assert split_plan(2, 4096) == (6, 6, 128)
    assert split_plan(4, 4096) == (6, 6, 256)
    assert split_plan(8, 4096) == (5, 7, 256)
    assert split_plan(8, 8192) == (5, 8, 256)
    assert split_plan(1, 4096) == (6, 6, 64)
    assert split_plan(2, 256) == (4, 4, 32)
    assert split_plan(8, 16384) == (0, 0, 8)       # would need a 9-bit pass 2
    assert split_plan(2, 100) == (0, 0, 2)         # not a power of two
    assert split_plan(2, 4096, pass1_in_shuffle=False) == (0, 0, 2)


def test_world_must_be_power_of_two():
    from partitionedhashjoin_b200 import multigpu
    with pytest.raises(ValueError):
        multigpu.ShardedRadixJoin(None, 0, 3, backend=None)


@pytest.mark.gpu
def test_sharded_join_single_gpu_backend(phj, oracle):
    """GpuBackend with one rank: shard split (1 owner) -> zero-copy bind -> local radix join."""
    import numpy as np

    import _cases
    from partitionedhashjoin_b200 import multigpu
    R = _cases.tuples(_cases.splitmix64(60000, 5).astype(np.int64) % 50021)
    S = _cases.tuples(_cases.splitmix64(900000, 6).astype(np.int64) % 70001)
    job = multigpu.ShardedRadixJoin(None, 0, 1, multigpu.GpuBackend(1, 0, partitions_local=256))
    job.upload(R, S)
    assert job.join()["matches"] == oracle.count_by_sort(R, S)
    job.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_split_kernel_matches_oracle(phj, oracle, world):
    """PHJ_ALGO_SHARD_SPLIT on the device: piece r holds exactly the tuples whose owner digit is r,
    in input order (what every rank sends to rank r)."""
    import numpy as np

    import _cases
    from partitionedhashjoin_b200 import multigpu
    seed = 0x9E3779B97F4A7C15
    R = _cases.tuples(_cases.splitmix64(50000, 15).astype(np.int64))
    S = _cases.tuples(_cases.splitmix64(700001, 16).astype(np.int64) % 12345)
    with phj.Engine("shard-split", partitions=world, hash_seed=seed, shard_shift=multigpu.SHARD_SHIFT) as e:
        e.upload(R, S)
        res = e.join()
        assert res["matches"] == 0 and res["passes"] == 1
        for which, rel in ((0, R), (1, S)):
            got, bounds = e.read_partitions(which, world)
            owner = ((oracle.hash_batch(0, seed, rel["id"]) >> np.uint64(multigpu.SHARD_SHIFT)) & np.uint64(world - 1)).astype(np.int64)
            want = rel[np.argsort(owner, kind="stable")]
            assert np.diff(bounds.astype(np.int64)).tolist() == np.bincount(owner, minlength=world).tolist()
            assert (got["id"] == want["id"]).all() and (got["payload"] == want["payload"]).all()


@pytest.mark.gpu
def test_fused_shuffle_single_gpu(phj, oracle):
    """FusedGpuBackend with one rank: phj_shard_count + phj_shard_scatter into the rank's own
    shared window (phj_shared_alloc), then the local join bound to the window."""
    import numpy as np

    import _cases
    from partitionedhashjoin_b200 import multigpu
    R = _cases.tuples(_cases.splitmix64(60000, 5).astype(np.int64) % 50021)
    S = _cases.tuples(_cases.splitmix64(900000, 6).astype(np.int64) % 70001)
    want = oracle.count_by_sort(R, S)
    for pass1, parts in ((True, 256), (True, 4096), (True, 16), (False, 256)):
        job = multigpu.FusedShardedRadixJoin(None, 0, 1, multigpu.FusedGpuBackend(1, 0, partitions_local=parts,
                                                                                  pass1_in_shuffle=pass1))
        job.upload(R, S)
        assert job.join()["matches"] == want
        assert job.join()["matches"] == want
        job.close()


@pytest.mark.gpu
@pytest.mark.parametrize("hash", ["xxh3", "murmur3", "city"])
def test_library_sharded_join_single_rank(phj, oracle, hash):
    """The sharded join inside the library (phj_dist_*) with ONE rank on one GPU: the whole pipeline -- count, NCCL
    all-gather, device-side layout, scatter into the (own) window chunk by chunk with stream-ordered barriers, L2
    tables, all-reduce -- against the oracle's count, on the adversarial cases, with the window contents checked."""
    import numpy as np

    import _cases
    import _dist_gpu_worker as worker
    from partitionedhashjoin_b200 import multigpu
    hash_id = {"xxh3": 0, "murmur3": 1, "city": 2}[hash]
    cases = dict(_cases.adversarial_cases())
    cases["big_random"] = (_cases.tuples(_cases.splitmix64(300_000, 51).astype(np.int64) % 200_003),
                           _cases.tuples(_cases.splitmix64(2_000_000, 52).astype(np.int64) % 300_007))
    for partitions, chunks in ((0, 0), (8, 3), (256, 16), (1, 1)):
        job = multigpu.ShardedJoin(None, 0, 1, 0, partitions=partitions, chunks=chunks, hash=hash)
        for name, (R, S) in cases.items():
            R, S = R.copy(), S.copy()
            R["payload"], S["payload"] = np.arange(R.shape[0]), np.arange(S.shape[0])
            job.upload(R, S)
            want = oracle.count_by_sort(R, S)
            for _ in range(2):
                res = job.join()
                assert res["matches"] == want, (name, partitions, chunks, res["matches"], want)
            if R.shape[0] and S.shape[0]:
                worker.check_window(job, oracle, R, 0, 0, 1, hash_id=hash_id)
                worker.check_window(job, oracle, S, 1, 0, 1, hash_id=hash_id)
        job.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4, 8, 3, 6])
def test_one_process_several_gpus(phj, oracle, gpus):
    """phj_config.num_gpus: one process, one host thread per GPU, the same sharded join (phj_upload shards the rows,
    phj_join returns the global count) -- against the oracle on adversarial and larger inputs, incl. re-uploads that
    grow the windows. Skipped when the box has fewer GPUs."""
    import numpy as np

    import _cases
    if phj.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    cases = dict(_cases.adversarial_cases())
    cases["big_random"] = (_cases.tuples(_cases.splitmix64(300_000, 51).astype(np.int64) % 200_003),
                           _cases.tuples(_cases.splitmix64(2_000_000, 52).astype(np.int64) % 300_007))
    for partitions, chunks in ((0, 0), (gpus, 3), (256 // gpus * gpus, 1)):
        with phj.Engine("radix-partitioning", partitions=partitions, split_chunks=chunks, num_gpus=gpus) as e:
            for name, (R, S) in cases.items():
                want = oracle.count_by_sort(R, S)
                e.upload(R, S)
                for _ in range(2):
                    res = e.join()
                    assert res["matches"] == want and res["gpus"] == gpus, (name, partitions, chunks, res, want)
                res = e.join_host(R, S)
                assert res["matches"] == want and res["h2d_bytes"] == 16 * (R.shape[0] + S.shape[0])
    # the no-partitioning joiner over the same GPUs: every GPU builds the whole table and probes its row shard
    for flags in (0, phj.FLAG_CHAINED_TABLE):
        with phj.Engine("no-partitioning", num_gpus=gpus, flags=flags) as e:
            for name, (R, S) in cases.items():
                e.upload(R, S)
                res = e.join()
                assert res["matches"] == oracle.count_by_sort(R, S) and res["gpus"] == gpus, (name, flags, res)
            assert e.join_host(R, S)["matches"] == oracle.count_by_sort(R, S)
    with pytest.raises(phj.PhjError):
        phj.Engine("radix-partitioning", partitions=gpus // 2 or 3, num_gpus=gpus)
    with pytest.raises(phj.PhjError):
        phj.Engine("no-partitioning", partitions=64, num_gpus=gpus)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8, 3, 6])
def test_library_sharded_join_on_gpus(phj, world):
    """One rank per GPU: the library's sharded join (NVLink peer stores + NCCL) against the oracle's count with the
    contents of every rank's windows checked; skipped when the box has fewer GPUs."""
    if phj.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = torchrun(world, os.path.join(HERE, "_dist_gpu_worker.py"), "lib", timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,mode", [(2, "pass1"), (2, "fused"), (2, "nccl"), (2, "npj"), (4, "pass1"), (8, "pass1")])
def test_sharded_join_on_gpus(phj, world, mode):
    """One rank per GPU over NCCL: the fused NVLink-store shuffle (and the all-to-all variant)
    against the oracle's count; skipped when the box has fewer GPUs."""
    if phj.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = torchrun(world, os.path.join(HERE, "_dist_gpu_worker.py"), mode, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0
