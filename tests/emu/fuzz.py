"""TEST INFRASTRUCTURE: randomised parity runs of the engine on the emulated build (tests/emu) against the oracle --
sizes around the tile boundaries, empty and one-tuple relations, duplicate and reserved keys, every plan / flag / hasher
combination, re-used handles, the streamed host join, the joined table, and the sharded join on 2 .. 8 emulated GPUs.
Not part of the default test run (it is open-ended); `python tests/emu/fuzz.py <first seed> <seeds> [minutes]` under the
environment tests/test_emulated.py sets up. Prints one line per failure and a summary; exit code 1 if anything failed.
"""
import os
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import checks  # noqa: E402

INT64_MIN = -(2 ** 63)
SIZES = [0, 1, 2, 31, 32, 33, 255, 256, 4095, 4096, 4097, 8191, 8192, 8193, 12288, 16384, 20000, 32768, 40000]


def relation(rng, n, key_space, cases):
    keys = rng.integers(1, max(2, key_space), size=n, dtype=np.int64)
    style = rng.integers(0, 6)
    if style == 0 and n:
        keys[:] = keys[0]                                   # all equal
    elif style == 1 and n:
        keys[rng.integers(0, n, size=max(1, n // 3))] = 7    # a heavy hitter
    elif style == 2 and n:
        keys[rng.integers(0, n, size=min(n, 3))] = INT64_MIN  # the reserved table marker as a key
    elif style == 3:
        keys = keys * np.int64(-7919)                        # negative / wide keys
    return cases.tuples(keys)


def one(seed, phj, oracle, cases):
    rng = np.random.default_rng(seed)
    n_r, n_s = int(rng.choice(SIZES)), int(rng.choice(SIZES)) * int(rng.choice([1, 1, 2]))
    space = int(rng.choice([50, 3000, 100_000, 2 ** 40]))
    R, S = relation(rng, n_r, space, cases), relation(rng, n_s, space, cases)
    want = oracle.count_by_sort(R, S)
    hash_name = str(rng.choice(["xxh3", "murmur3", "city"]))
    seed_p = int(rng.integers(0, 2 ** 63))
    kind = int(rng.integers(0, 10))
    desc = {"seed": seed, "n": (n_r, n_s), "space": space, "hash": hash_name}
    if kind <= 1:  # no-partitioning
        flags = int(rng.choice([0, phj.FLAG_CHAINED_TABLE, phj.FLAG_COOP_PROBE]))
        desc.update(algo="npj", flags=flags)
        with phj.Engine("no-partitioning", hash=hash_name, hash_seed=seed_p, flags=flags) as e:
            if rng.integers(0, 2):
                got = e.join_host(R, S)["matches"]
            else:
                e.upload(R, S)
                got = e.join()["matches"]
        assert got == want, (desc, got, want)
    elif kind <= 6:  # radix on one GPU
        P = int(rng.choice([0, 1, 2, 3, 7, 32, 64, 100, 256, 1000, 1024, 4096, 16384]))
        bits = (0, 0)
        if P and not P & (P - 1) and rng.integers(0, 2):
            total = P.bit_length() - 1
            b1 = int(rng.integers(max(1, total - 8), min(8, total) + 1)) if total else 0
            bits = (b1, total - b1)
        flags = 0
        for f in (phj.FLAG_NO_TMA_STORE, phj.FLAG_NO_HIST12, phj.FLAG_NO_FUSE_HIST2, phj.FLAG_L2_TABLES):
            if rng.integers(0, 4) == 0:
                flags |= f
        chunks = int(rng.integers(0, 6))
        desc.update(algo="radix", P=P, bits=bits, flags=flags, chunks=chunks)
        with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, hash=hash_name, hash_seed=seed_p,
                        flags=flags, upload_chunks=chunks) as e:
            op = int(rng.integers(0, 4))
            desc["op"] = op
            if op == 0:
                got = e.join_host(R, S)["matches"]
                assert got == want, (desc, got, want)
            else:
                e.upload(R, S)
                res = e.join()
                assert res["matches"] == want, (desc, res["matches"], want)
                if op == 1 and P and n_r + n_s:  # the partitioned relations, bit for bit
                    hid = {"xxh3": 0, "murmur3": 1, "city": 2}[hash_name]
                    for which, rel in ((0, R), (1, S)):
                        got_p, gb = e.read_partitions(which, P)
                        wt, wb = oracle.radix_partition(rel, P, hid, seed_p, workers=1)
                        checks.same_partitioning(got_p, gb, wt, wb, rel.shape[0])
                rows = 0
                if op == 2:  # rows of the joined table: skip the quadratic blow-ups (all keys equal on both sides)
                    ku, kc = np.unique(R["id"], return_counts=True)
                    pos = np.searchsorted(ku, S["id"])
                    hit = (pos < ku.shape[0]) & (ku[np.minimum(pos, max(ku.shape[0] - 1, 0))] == S["id"]) if ku.shape[0] else np.zeros(n_s, bool)
                    rows = int(kc[pos[hit]].sum()) if ku.shape[0] else 0
                if op == 2 and rows <= 300_000:  # the joined table
                    rm = e.join_materialize()
                    ref = cases.sorted_rows(oracle.join_materialize(R, S))
                    assert rm["joined_tuples"] == ref.shape[0] == rows, (desc, rm["joined_tuples"], ref.shape[0], rows)
                    assert (cases.sorted_rows(e.read_joined()) == ref).all(), desc
                if op == 3:  # the handle again, on other relations
                    R2, S2 = relation(rng, int(rng.choice(SIZES)), space, cases), relation(rng, int(rng.choice(SIZES)), space, cases)
                    e.upload(R2, S2)
                    got = e.join()["matches"]
                    assert got == oracle.count_by_sort(R2, S2), (desc, "second upload")
    else:  # sharded over emulated GPUs
        gpus = int(rng.integers(2, 9))
        local = int(rng.choice([0, 1, 2, 4, 8]))
        P = gpus * local
        if P and not gpus & (gpus - 1):  # power-of-two GPU counts want a power-of-two digit count
            P = gpus * (1 << int(rng.integers(0, 4)))
        chunks = int(rng.integers(0, 7))
        flags = 0
        if rng.integers(0, 3) == 0:
            flags |= phj.FLAG_COUNT_PIECEWISE
        if rng.integers(0, 3) == 0:
            flags |= phj.FLAG_NO_HOT_DIGITS
        desc.update(algo="sharded", gpus=gpus, P=P, chunks=chunks, flags=flags)
        if rng.integers(0, 5) == 0:  # the no-partitioning joiner over the same GPUs (build side replicated)
            table = int(rng.choice([0, phj.FLAG_CHAINED_TABLE]))
            desc.update(algo="npj replicas", flags=table)
            with phj.Engine("no-partitioning", num_gpus=gpus, hash=hash_name, hash_seed=seed_p, flags=table) as e:
                e.upload(R, S)
                res = e.join()
                assert res["matches"] == want and res["gpus"] == gpus, (desc, res["matches"], want)
                assert e.join_host(R, S)["matches"] == want, (desc, "join_host")
            return desc
        with phj.Engine("radix-partitioning", partitions=P, split_chunks=chunks, num_gpus=gpus, hash=hash_name,
                        hash_seed=seed_p, flags=flags) as e:
            e.upload(R, S)
            for _ in range(2):
                res = e.join()
                assert res["matches"] == want and res["gpus"] == gpus, (desc, res["matches"], want)
            if rng.integers(0, 2):
                R2, S2 = relation(rng, int(rng.choice(SIZES)), space, cases), relation(rng, 2 * int(rng.choice(SIZES)), space, cases)
                got = e.join_host(R2, S2)["matches"]
                assert got == oracle.count_by_sort(R2, S2), (desc, "second relations through join_host")
    return desc


def main():
    first, count = int(sys.argv[1]), int(sys.argv[2])
    minutes = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
    import _cases
    phj, oracle = checks.load()
    t0, done, failed = time.time(), 0, 0
    for seed in range(first, first + count):
        if time.time() - t0 > minutes * 60:
            break
        try:
            one(seed, phj, oracle, _cases)
        except Exception as exc:  # noqa: BLE001
            failed += 1
            print(f"FAIL seed {seed}: {type(exc).__name__}: {str(exc)[:600]}", flush=True)
            if os.environ.get("PHJ_FUZZ_TRACE"):
                traceback.print_exc()
        done += 1
        if done % 50 == 0:
            print(f"  ... {done} runs, {failed} failed, {time.time() - t0:.0f} s", flush=True)
    print(f"seeds {first}..{first + done - 1}: {done} runs, {failed} failed, {time.time() - t0:.0f} s", flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
