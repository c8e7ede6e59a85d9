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
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), script, *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world,case", [(2, "random"), (2, "skewed"), (2, "tiny"), (4, "random"), (1, "random")])
def test_sharded_join_over_gloo(world, case):
    r = torchrun(world, os.path.join(HERE, "_dist_worker.py"), case)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0


@pytest.mark.parametrize("world,case,mode", [(2, "random", "pass1"), (2, "skewed", "pass1"), (2, "tiny", "pass1"),
                                             (4, "random", "pass1"), (1, "random", "pass1"), (2, "random", "fused"),
                                             (2, "random", "pipelined"), (2, "skewed", "pipelined"),
                                             (2, "tiny", "pipelined"), (4, "random", "pipelined"),
                                             (1, "random", "pipelined"), (2, "skewed", "pipelined-ce"),
                                             (4, "random", "pipelined-ce")])
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
@pytest.mark.parametrize("world,mode", [(2, "pipelined"), (2, "pipelined-sm"), (2, "pass1"), (2, "fused"), (2, "nccl"), (4, "pass1"),
                                        (8, "pass1"), (8, "pipelined")])
def test_sharded_join_on_gpus(phj, world, mode):
    """One rank per GPU over NCCL: the fused NVLink-store shuffle (and the all-to-all variant)
    against the oracle's count; skipped when the box has fewer GPUs."""
    if phj.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = torchrun(world, os.path.join(HERE, "_dist_gpu_worker.py"), mode, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == world and line["matches"] == line["want"] and line["want"] > 0
