"""Deterministic join inputs shared by the golden-vector generator and the parity tests.

Nothing here depends on numpy's RNG streams: pseudo-random keys come from splitmix64 written with
uint64 arithmetic, generator-shaped relations from the product's / oracle's seeded generators.
Every case is a pair (R, S) of TUPLE arrays; names are stable because tests/golden/ keys on them.
"""
from __future__ import annotations

import numpy as np

TUPLE = np.dtype([("id", "<i8"), ("payload", "<i8")], align=True)
INT64_MIN, INT64_MAX = -(2**63), 2**63 - 1


def splitmix64(n: int, seed: int) -> np.ndarray:
    """n pseudo-random uint64 (Vigna's splitmix64), reproducible everywhere."""
    with np.errstate(over="ignore"):
        x = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def tuples(ids) -> np.ndarray:
    ids = np.asarray(ids).astype(np.int64)
    out = np.empty(ids.shape[0], dtype=TUPLE)
    out["id"] = ids
    out["payload"] = np.arange(ids.shape[0], dtype=np.int64)
    return out


def sequential(n, start=1):
    return tuples(np.arange(start, start + n, dtype=np.int64))


def adversarial_cases():
    """inputs whose answer is not trivially |S|."""
    cases = {}
    # (i) duplicate build keys: every probe tuple still counts once
    r = np.concatenate([np.arange(1, 501), np.arange(1, 501), np.full(300, 7)])
    s = (splitmix64(5000, 11) % np.uint64(1000)).astype(np.int64)  # half of them miss
    cases["dup_build_keys"] = (tuples(r), tuples(s))
    # (ii) probe keys outside the build domain, incl. 0, negatives and the int64 extremes
    r = np.concatenate([np.arange(-50, 50), [INT64_MIN, INT64_MAX, 0]])
    s = np.concatenate([np.arange(-100, 100), [INT64_MIN, INT64_MAX, INT64_MIN + 1, INT64_MAX - 1, 0, 0, INT64_MIN]])
    cases["extreme_keys"] = (tuples(r), tuples(s))
    # ... and the same without the sentinel-looking keys on the build side
    cases["extreme_probe_only"] = (tuples(np.arange(-50, 50)), tuples(s))
    # (iii) all probe tuples equal: one partition holds 100 % of S
    cases["all_probe_equal_hit"] = (sequential(3000), tuples(np.full(20000, 1234)))
    cases["all_probe_equal_miss"] = (sequential(3000), tuples(np.full(20000, 99999)))
    # (iv) awkward sizes: primes, |R| < P, |R| = 1, empty S
    cases["prime_sizes"] = (tuples(splitmix64(1009, 3).astype(np.int64)),
                            tuples(np.concatenate([splitmix64(1009, 3).astype(np.int64)[::3],
                                                   splitmix64(4999, 4).astype(np.int64)])))
    cases["tiny_build"] = (tuples([42]), tuples([42, 43, 42, -42, 0, 42]))
    cases["build_smaller_than_fanout"] = (sequential(17), tuples(np.arange(0, 40)))
    cases["empty_probe"] = (sequential(100), tuples(np.empty(0, dtype=np.int64)))
    # random 64-bit keys on both sides with ~30 % overlap
    pool = splitmix64(20000, 21).astype(np.int64)
    cases["random64"] = (tuples(pool[:6000]), tuples(np.concatenate([pool[4000:12000], pool[:1000]])))
    # every build key identical (a single table entry, heavily contended insert)
    cases["all_build_equal"] = (tuples(np.full(5000, -7)), tuples(np.concatenate([np.full(100, -7), np.arange(50)])))
    return cases


GENERATOR_CASES = {
    # name: (|R|, |S|, alpha, base_seed, batches) -- R = Sequential(start=1), S = Zipf over [1, |R|]
    "gen_small_105": (20000, 200000, 1.05, 12345, 8),
    "gen_small_125": (20000, 200000, 1.25, 777, 5),
    "gen_small_uniform": (20000, 200000, 0.01, 99, 3),
    "gen_ragged": (12347, 100003, 0.99, 4242, 7),
}


def joined_reference(R, S):
    """Brute-force definition of the joined table (id, payloadA, payloadB), canonically sorted:
    one row per (probe tuple, equal-key build tuple). numpy only; for the tests of the oracle."""
    order = np.argsort(R["id"], kind="stable")
    keys, pay = R["id"][order], R["payload"][order]
    lo = np.searchsorted(keys, S["id"], side="left")
    hi = np.searchsorted(keys, S["id"], side="right")
    cnt = hi - lo
    rows = np.empty(int(cnt.sum()), dtype=[("id", "<i8"), ("payloadA", "<i8"), ("payloadB", "<i8")])
    probe_idx = np.repeat(np.arange(S.shape[0]), cnt)
    within = np.arange(rows.shape[0]) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    rows["id"] = S["id"][probe_idx]
    rows["payloadA"] = pay[lo[probe_idx] + within]
    rows["payloadB"] = S["payload"][probe_idx]
    return np.sort(rows, order=["id", "payloadB", "payloadA"])


def sorted_rows(rows):
    return np.sort(np.asarray(rows), order=["id", "payloadB", "payloadA"])
