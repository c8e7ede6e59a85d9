"""TEST INFRASTRUCTURE: parity checks that run the product's ENGINE and KERNEL SOURCE on the CPU (tests/emu: the
unmodified phj_engine.cu linked against a host stand-in for the CUDA runtime, every kernel launch executed by a fiber
scheduler over the g++-compiled kernel headers). Started by tests/test_emulated.py in a subprocess whose PHJ_LIB points
at the emulated build; compares with the oracle exactly as the `-m gpu` tests do on the B200, at sizes the emulator
finishes in seconds. Not a product path and not a source of numbers.

    python tests/emu/checks.py single | group | cli | failure | ranks <world>
"""
import json
import os
import sys
import threading
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.path.insert(0, os.path.dirname(TESTS))
SEED_P = 0x9E3779B97F4A7C15


def load():
    import partitionedhashjoin_b200 as phj

    import _oracle
    assert "emu" in os.path.basename(phj.LIB_PATH), f"these checks are for the emulated build, not {phj.LIB_PATH}"
    assert "emulated" in phj.device_info(0)["name"]
    return phj, _oracle.Oracle()


def same_partitioning(got, gb, want, wb, n):
    assert gb[:-1].tolist() == wb[:, 0].tolist() and int(gb[-1]) == n
    assert (got["id"] == want["id"]).all() and (got["payload"] == want["payload"]).all()


def check_single():
    """One emulated GPU: every join algorithm and plan against the oracle on the adversarial cases, partitions
    bit-identical to the oracle's stable radix partitioning, the joined table, the streamed host join."""
    import _cases
    phj, oracle = load()
    done = []
    for hash_id, name in enumerate(("xxh3", "murmur3", "city")):
        keys = _cases.splitmix64(5000, 5 + hash_id).astype(np.int64)
        assert (phj.hash_batch(hash_id, 7, keys) == oracle.hash_batch(hash_id, 7, keys)).all(), name
    done.append("3 device hashers bit-exact")
    t_sec = time.time()

    def lap(what):
        nonlocal t_sec
        if os.environ.get("PHJ_EMU_TIMES"):
            print(f"  [{time.time() - t_sec:6.1f} s] {what}", file=sys.stderr, flush=True)
        t_sec = time.time()
    plans = [("no-partitioning", {}), ("no-partitioning", {"flags": phj.FLAG_CHAINED_TABLE}),
             ("no-partitioning", {"flags": phj.FLAG_COOP_PROBE}),
             ("radix-partitioning", {"partitions": 0}), ("radix-partitioning", {"partitions": 64}),
             ("radix-partitioning", {"partitions": 100}), ("radix-partitioning", {"partitions": 1024}),
             ("radix-partitioning", {"partitions": 1024, "flags": phj.FLAG_NO_HIST12}),
             ("radix-partitioning", {"partitions": 32, "flags": phj.FLAG_L2_TABLES, "hash": "city"})]
    # a wide first pass makes 256 parents of a few tuples each, one CTA per pass-2 segment: slow here, two cases do
    few = [("radix-partitioning", {"partitions": 512, "radix_bits": (8, 1), "hash": "murmur3"})]
    cases = _cases.adversarial_cases()
    n = 0
    for algo, kw in plans + few:
        with phj.Engine(algo, hash_seed=SEED_P, **kw) as e:
            for name, (R, S) in cases.items():
                if (algo, kw) in few and name not in ("tiny_build", "all_build_equal"):
                    continue
                want = oracle.count_by_sort(R, S)
                e.upload(R, S)
                res = e.join()
                assert res["matches"] == want, (algo, kw, name, res["matches"], want)
                n += 1
        lap(f"{algo} {kw}")
    done.append(f"{n} counts == oracle ({len(plans)} plans x {len(cases)} adversarial cases + a wide first pass)")

    # generator-shaped data (Zipf 1.25: a heavy hitter), two passes: count, and both partitioned relations
    # bit-identical to the reference algorithm's stable partitioning -- with every way pass 2 gets its histogram,
    # with and without TMA-style bulk flushes, and for a fan-out that is not a power of two
    R = _cases.sequential(5000)
    S = np.empty(40_011, dtype=phj.TUPLE_DTYPE)
    phj.fill_zipf(S, 1.25, 1, 7500, 777, 5)
    want = oracle.count_by_sort(R, S)
    layouts = 0
    for P, bits, flags in ((4096, (6, 6), 0), (4096, (6, 6), phj.FLAG_NO_HIST12), (4096, (6, 6), phj.FLAG_NO_FUSE_HIST2),
                           (4096, (6, 6), phj.FLAG_NO_TMA_STORE), (256, (8, 0), 0), (1000, (0, 0), 0)):
        with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, hash_seed=SEED_P, flags=flags) as e:
            e.upload(R, S)
            assert e.join()["matches"] == want, (P, bits, flags)
            for which, rel in ((0, R), (1, S)):
                got, gb = e.read_partitions(which, P)
                wt, wb = oracle.radix_partition(rel, P, 0, SEED_P, workers=1)
                same_partitioning(got, gb, wt, wb, rel.shape[0])
                layouts += 1
        lap(f"layout {P} {bits} {flags:#x}")
    # the exchange step on its own (PHJ_ALGO_SHARD_SPLIT): a digit count that is a bit field of the hash, and two that
    # are not (what 3 and 6 GPUs x their local partitions give)
    for P in (64, 63, 60):
        with phj.Engine("shard-split", partitions=P, hash_seed=SEED_P) as e:
            e.upload(R, S)
            e.join()
            for which, rel in ((0, R), (1, S)):
                got, gb = e.read_partitions(which, P)
                wt, wb = oracle.radix_partition(rel, P, 0, SEED_P, workers=1)
                same_partitioning(got, gb, wt, wb, rel.shape[0])
                layouts += 1
    lap("shard-split layouts")
    done.append(f"{layouts} partitioned relations bit-identical to the oracle's")

    # the joined table against the oracle's GetAll join, as a multiset
    for name in ("dup_build_keys", "random64", "prime_sizes"):
        Rm, Sm = cases[name]
        with phj.Engine("radix-partitioning", partitions=64, hash_seed=SEED_P) as e:
            e.upload(Rm, Sm)
            res = e.join_materialize()
            rows = _cases.sorted_rows(e.read_joined())
        ref = _cases.sorted_rows(oracle.join_materialize(Rm, Sm))
        assert res["joined_tuples"] == ref.shape[0] and (rows == ref).all(), name
        lap(f"materialize {name}")
    done.append("joined table == oracle on 3 cases")

    # the streamed host join (helper thread uploads chunks, a child handle joins them as they land)
    with phj.Engine("radix-partitioning", partitions=256, upload_chunks=3, hash_seed=SEED_P) as e:
        res = e.join_host(R, S)
        assert res["matches"] == want and res["upload_chunks"] == 3, res
    with phj.Engine("no-partitioning", upload_chunks=2) as e:
        assert e.join_host(R, S)["matches"] == want
    # oversize build partitions go through the global table
    with phj.Engine("radix-partitioning", partitions=2, hash_seed=SEED_P) as e:
        Rb = _cases.tuples(_cases.splitmix64(40_000, 3).astype(np.int64) % 30_011)
        Sb = _cases.tuples(_cases.splitmix64(50_000, 4).astype(np.int64) % 45_007)
        e.upload(Rb, Sb)
        res = e.join()
        assert res["matches"] == oracle.count_by_sort(Rb, Sb) and res["fallback_partitions"] > 0, res
    lap("streamed + fallback")
    done.append("streamed host join, oversize fallback")
    return done


def check_group():
    """phj_config.num_gpus on 2 / 3 / 4 / 6 / 8 emulated GPUs (one process, the engine's own host thread per GPU, peer
    stores into the other threads' windows, collectives between the threads): the global count against the oracle,
    on uniform and heavy-hitter keys, re-uploads that grow the windows, the one-call host join."""
    import _cases
    phj, oracle = load()
    keep = ("dup_build_keys", "extreme_keys", "all_probe_equal_hit", "tiny_build", "build_smaller_than_fanout",
            "empty_probe", "random64")
    cases = {k: v for k, v in _cases.adversarial_cases().items() if k in keep}
    R = _cases.sequential(6000)
    S = np.empty(40_001, dtype=phj.TUPLE_DTYPE)
    phj.fill_zipf(S, 1.25, 1, 12_000, 4242, 7)   # half of the key range misses R; key 1 is 22 % of S
    cases["zipf125_half_miss"] = (R, S)
    cases["random_larger"] = (_cases.tuples(_cases.splitmix64(20_000, 51).astype(np.int64) % 15_013),
                              _cases.tuples(_cases.splitmix64(50_000, 52).astype(np.int64) % 20_011))
    done, n = [], 0
    for gpus in (2, 3, 4, 6, 8):  # 3 and 6: the split digit is hash % digits instead of a bit field
        for partitions, chunks, flags in ((0, 0, 0), (gpus, 3, phj.FLAG_COUNT_PIECEWISE),
                                          (256 // gpus * gpus, 1, phj.FLAG_NO_HOT_DIGITS)):
            with phj.Engine("radix-partitioning", partitions=partitions, split_chunks=chunks, num_gpus=gpus,
                            flags=flags) as e:
                for name, (Rc, Sc) in cases.items():
                    want = oracle.count_by_sort(Rc, Sc)
                    e.upload(Rc, Sc)
                    res = e.join()
                    assert res["matches"] == want and res["gpus"] == gpus, (gpus, partitions, chunks, name, res, want)
                    n += 1
                res = e.join_host(R, S)
                assert res["matches"] == oracle.count_by_sort(R, S)
        # the no-partitioning joiner over the same GPUs: whole build relation everywhere, probe rows sharded
        for flags in ((0, phj.FLAG_CHAINED_TABLE) if gpus in (2, 3) else (0,)):
            with phj.Engine("no-partitioning", num_gpus=gpus, flags=flags) as e:
                for name, (Rc, Sc) in cases.items():
                    e.upload(Rc, Sc)
                    res = e.join()
                    assert res["matches"] == oracle.count_by_sort(Rc, Sc) and res["gpus"] == gpus, (gpus, name, flags, res)
                    n += 1
                assert e.join_host(R, S)["matches"] == oracle.count_by_sort(R, S)
        done.append(f"{gpus} GPUs ok")
    done.append(f"{n} joins over several GPUs == oracle")
    return done


def check_failure():
    """An allocation that fails on ONE of the GPUs of a sharded join: every rank must come back with an error (the
    ranks agree on the failure before the next collective) instead of the others waiting for the one that gave up; the
    handle works again once memory is there."""
    import _cases
    phj, oracle = load()
    R = _cases.sequential(5000)
    S = np.empty(60_000, dtype=phj.TUPLE_DTYPE)
    phj.fill_zipf(S, 0.5, 1, 8000, 99, 4)
    want = oracle.count_by_sort(R, S)
    done = []
    for gpus, victim in ((4, 2), (3, 0)):
        with phj.Engine("radix-partitioning", num_gpus=gpus) as e:
            e.upload(R, S)
            os.environ["PHJ_EMU_FAIL_MALLOC"] = f"{victim}:100000"   # the probe window of that GPU (~ 270 KB)
            try:
                try:
                    e.join()
                    raise AssertionError("the join went through although a window could not be allocated")
                except phj.PhjError as exc:
                    assert exc.status == 4 and f"GPU {victim}" in str(exc) and "no memory for" in str(exc), str(exc)
            finally:
                del os.environ["PHJ_EMU_FAIL_MALLOC"]
            for _ in range(2):
                assert e.join()["matches"] == want
        done.append(f"{gpus} GPUs: a failed window allocation on GPU {victim} is an error on every rank, then recovers")
    return done


class ThreadDist:
    """What multigpu.ShardedJoin needs of torch.distributed, between threads: the NCCL id travels from rank 0."""

    def __init__(self, world):
        self.world, self.slot, self.barrier_, self.me = world, None, threading.Barrier(world), threading.local()

    def broadcast_object_list(self, box, src=0):
        if self.me.rank == src:
            self.slot = box[0]
        self.barrier_.wait()
        box[0] = self.slot
        self.barrier_.wait()

    def barrier(self):
        self.barrier_.wait()


def check_ranks(world):
    """One rank per THREAD through the one-process-per-GPU entry points (phj_dist_create / upload / join: the path
    `bench.py --gpus N` takes under torchrun), with the window contents checked tuple by tuple exactly as
    tests/_dist_gpu_worker.py does on real GPUs: both count modes, several digit / chunk shapes, a collective re-size,
    heavy-hitter digits kept local against the plain exchange."""
    # torch stays out of this process (it would bring the real NCCL): multigpu / the worker module import it only to
    # bind torch's NCCL first and for the torchrun launch, neither of which happens here
    torch_stub, dist_stub = types.ModuleType("torch"), types.ModuleType("torch.distributed")
    torch_stub.distributed = dist_stub
    sys.modules.setdefault("torch", torch_stub)
    sys.modules.setdefault("torch.distributed", dist_stub)
    import _cases
    phj, oracle = load()
    import _dist_gpu_worker as worker
    keys_r = _cases.splitmix64(6_000, 21).astype(np.int64) % 4_001
    keys_s = np.where(_cases.splitmix64(30_000, 22) % np.uint64(10) < 3, 4242,
                      _cases.splitmix64(30_000, 23) % np.uint64(6_007)).astype(np.int64)
    R, S = _cases.tuples(keys_r), _cases.tuples(keys_s)
    want = oracle.count_by_sort(R, S)
    dist = worker.dist = ThreadDist(world)   # the module-level name check_library_join hands to ShardedJoin
    errors = []

    def run(rank):
        try:
            dist.me.rank = rank
            worker.check_library_join(rank, world, rank, oracle, R, S, want, lambda rel: worker.shard_of(rel, rank, world),
                                      "lib-emulated")
        except BaseException as exc:  # noqa: BLE001 -- reported by the main thread
            errors.append((rank, repr(exc)))
            dist.barrier_.abort()
            raise

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return [f"{world} ranks: counts, windows tuple by tuple, re-size, hot digits vs plain exchange"]


def check_cli():
    """The C++ host mirror (host/main.cpp: the reference's CLI and joiner classes) linked against the emulated engine:
    every joiner, the streamed upload, the joined table, --gpus N for both joiners incl. a GPU count that is not a power
    of two. The CLI joins generator data (every probe key has a build match), so the logged count is |secondary|."""
    import subprocess
    exe = os.path.join(HERE, "_build", "phjoin_emu")
    out = os.path.join(HERE, "_build", "cli_result.json")
    runs = [("radix-partitioning", ["-p", "256"]), ("radix-partitioning", ["-p", "100", "--hash", "murmur3"]),
            ("no-partitioning", []), ("no-partitioning", ["--table", "separate-chaining"]),
            ("radix-partitioning", ["-p", "64", "--stream-upload"]), ("radix-partitioning", ["-p", "64", "--materialize"]),
            ("radix-partitioning", ["--gpus", "2"]), ("radix-partitioning", ["--gpus", "3"]),
            ("radix-partitioning", ["--gpus", "6", "-p", "12"]), ("no-partitioning", ["--gpus", "2"]),
            ("no-partitioning", ["--gpus", "5", "--table", "separate-chaining"])]
    for join, extra in runs:
        r = subprocess.run([exe, "--join", join, "--primary", "6000", "--secondary", "50000", "--skew", "1.25", "--log", "debug",
                            "-f", out, *extra], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Joined 50000 tuples." in r.stderr, (join, extra, r.returncode, r.stderr[-600:])
        d = json.load(open(out))
        want_type = "RadixParitioning" if join.startswith("radix") else "NoPartitioning"
        assert d["parameters"]["Type"] == want_type and int(d["results"]["probe"]) >= 0, d
    r = subprocess.run([exe, "--join", "radix-partitioning", "--gpus", "3", "-p", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "multiple of the 3 GPUs" in r.stderr, r.stderr[-400:]
    return [f"{len(runs)} CLI runs logged the expected count"]


def main():
    t0 = time.time()
    what = sys.argv[1]
    done = {"single": check_single, "group": check_group, "cli": check_cli, "failure": check_failure}[what]() if what != "ranks" \
        else check_ranks(int(sys.argv[2]))
    print(json.dumps({"check": " ".join(sys.argv[1:]), "ok": True, "seconds": round(time.time() - t0, 1), "done": done}))


if __name__ == "__main__":
    main()
