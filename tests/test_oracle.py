"""CPU tests: pin the oracle (oracle/phj_oracle.c) against
  (a) the golden vectors committed under tests/golden/ (made by the unmodified reference), and
  (b) when available, the unmodified reference itself (oracle/_ref/libphj_ref.so), live.
"""
import json
import os
import subprocess
import sys
import sysconfig

import numpy as np
import pytest

import _cases
import _oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
SEED_P, SEED_T = 0x9E3779B97F4A7C15, 1


def all_cases(oracle):
    cases = dict(_cases.adversarial_cases())
    for name, (nr, ns, alpha, seed, batches) in _cases.GENERATOR_CASES.items():
        cases[name] = (oracle.fill_sequential(nr, 1), oracle.fill_zipf(ns, alpha, 1, nr, seed, batches))
    return cases


# SURVEY.md Appendix A.1 (independently recorded before this repo existed)
SURVEY_XXH3 = {
    (0, 0): 0xc77b3abb6f87acd9, (1, 0): 0x2fbc593564db792e, (2, 1): 0x1722dee50e00c79c,
    (3, 0x9E3779B97F4A7C15): 0x7fa7d72697eacbba, (10000000, 0): 0xfec30657b4a53159,
    (123456789, 1): 0x9a50464131fae77d, (-1, 0x9E3779B97F4A7C15): 0x8ae69d4026ba9e1e,
    (2**63 - 1, 0): 0xa233cf376558df46,
}


def test_xxh3_survey_vectors(oracle):
    for (key, seed), want in SURVEY_XXH3.items():
        assert oracle.hash_raw(0, key, seed) == want


def test_xxh3_golden(oracle):
    for rec in GOLDEN["xxh3"]:
        assert oracle.hash_raw(0, rec["key"], rec["seed"]) == rec["hash"], rec


def test_hash_modulo(oracle):
    # IHasher::Hash = hash % cardinality (src/Common/XXHasher.hpp:21); survey: seed-0 column % 2048
    want = [1241, 302, 579, 1661, 345, 725, 1043, 1862]
    keys = [0, 1, 2, 3, 10000000, 123456789, -1, 2**63 - 1]
    got = [int(oracle.lib.phjo_hash(0, k, 0, 2048)) for k in keys]
    assert got == want


def test_lcg_golden(oracle):
    for seed, want in GOLDEN["lcg"].items():
        assert oracle.lcg_doubles(int(seed), len(want)) == want
    # SURVEY.md Appendix A.2
    assert oracle.lcg_doubles(123456789, 2) == [0.21841829699390489, 0.95631757655940841]


def test_zipf_golden(oracle):
    for rec in GOLDEN["zipf"]:
        got = oracle.zipf_samples(rec["alpha"], rec["cardinality"], rec["seed"], len(rec["samples"]))
        assert got == rec["samples"], rec["alpha"]


def test_zipf_rejects_small_alpha(oracle):
    with pytest.raises(ValueError):
        oracle.zipf_samples(0.005, 10, 1, 1)  # src/DataGenerator/Zipf.cpp:18-20


def test_zipf_reference_unit_test(oracle):
    """tests/DataGenerator/ZipfTest.hpp:15-51 restated: in range, frequencies non-increasing."""
    samples = oracle.zipf_samples(0.99, 10, 123456789, 10000)
    counts = np.bincount(samples, minlength=11)
    assert min(samples) >= 1 and max(samples) <= 10
    present = counts[1:][counts[1:] > 0]
    assert (np.diff(present) <= 0).all()


def test_fill_golden(oracle):
    f = GOLDEN["fill_zipf"]
    z = oracle.fill_zipf(f["n"], f["alpha"], f["range"][0], f["range"][1], f["base_seed"], f["batches"])
    assert z["id"].tolist() == f["ids"] and (z["payload"] == np.arange(f["n"])).all()
    s = GOLDEN["fill_sequential"]
    q = oracle.fill_sequential(s["n"], s["start"])
    assert q["id"][:5].tolist() == s["first"] and int(q["id"][-1]) == s["last"]
    assert (q["payload"] == np.arange(s["n"])).all()


@pytest.mark.parametrize("name", sorted(GOLDEN["joins"]))
def test_join_counts_golden(oracle, name):
    R, S = all_cases(oracle)[name]
    want = GOLDEN["joins"][name]
    assert (R.shape[0], S.shape[0]) == (want["n_build"], want["n_probe"])
    assert oracle.count_by_sort(R, S) == want["matches"]
    if R.shape[0] == 0:
        return
    assert oracle.join_npj(R, S, table_kind=0) == want["matches"]
    assert oracle.join_npj(R, S, table_kind=1) == want["matches"]
    for P in (1, 3, 32, 64, 100, 2048):
        for workers in (1, 7):
            assert oracle.join_radix(R, S, P, workers=workers, seed_partition=SEED_P) == want["matches"], (P, workers)
    assert oracle.join_radix(R, S, 32, workers=3, table_kind=1) == want["matches"]
    for hash_id in (1, 2):
        assert oracle.join_radix(R, S, 64, workers=2, hash_id=hash_id) == want["matches"]


@pytest.mark.parametrize("kind", [0, 1])
def test_materialized_join_is_the_getall_join(oracle, kind):
    """phjo_join_materialize (the joined table the reference declares, filled through GetAll) against
    the brute-force definition, on every adversarial case and a generator case; its distinct-probe
    projection is the reference's count."""
    cases = dict(_cases.adversarial_cases())
    R = oracle.fill_sequential(3000, 1)
    cases["gen"] = (R, oracle.fill_zipf(20000, 1.05, 1, 6000, 4242, 3))
    for name, (R, S) in cases.items():
        got = oracle.join_materialize(R, S, table_kind=kind)
        want = _cases.joined_reference(R, S)
        assert got.shape[0] == want.shape[0], name
        assert (_cases.sorted_rows(got) == want).all(), name
        assert np.unique(got["payloadB"]).shape[0] == oracle.count_by_sort(R, S), name


def test_npj_empty_build_is_an_error(oracle):
    # LinearProbingHashTable's constructor throws for 0 objects (src/HashTables/LinearProbing.hpp:106-110)
    with pytest.raises(ValueError):
        oracle.join_npj(_cases.tuples([]), _cases.tuples([1, 2]))


def test_tables_golden(oracle):
    one = _cases.tuples([123456789])
    dup = _cases.tuples(np.full(10, 123456789))
    many = _cases.sequential(1000, 0)
    builds = {"InsertGetAndExists": one, "Iterator": dup, "MultiThreadedInsert": many}
    for rec in GOLDEN["tables"]:
        buckets, ex, gi, ga = oracle.table_probe(rec["kind"], rec["ratio"], 0, rec["seed"], rec["objects"],
                                                 builds[rec["test"]], rec["probe"])
        assert buckets == rec["buckets"], rec
        assert ex == rec["exists"] and ga == rec["getall"], rec
        if rec["get_index"] is not None:  # deterministic only for a single-threaded build
            assert gi == rec["get_index"], rec


def test_reference_table_unit_tests_restated(oracle):
    """tests/NoPartitioningHashJoin/HashTableTest.hpp:84-164 on the oracle tables."""
    t = _cases.tuples([123456789])
    t["payload"] = 987654321
    for kind, ratio in ((1, 0.3), (0, 1 / 0.75)):
        _, ex, gi, _ = oracle.table_probe(kind, ratio, 0, 99, 10, t, [123456789])
        assert ex == [1] and gi == [0]  # Get returns the very pointer that was inserted
        _, _, _, ga = oracle.table_probe(kind, ratio, 0, 99, 10, _cases.tuples(np.full(10, 123456789)), [123456789])
        assert ga == [10]  # duplicates are retained
    _, ex, _, _ = oracle.table_probe(1, 0.1, 0, 99, 1000, _cases.sequential(1000, 0), list(range(1000)))
    assert all(ex)


def test_partitioning_configuration(oracle):
    # src/RadixCluster/HashJoin.hpp:149-188, including the stale-batch quirk
    assert oracle.partitioning_configuration(10_000_000, 200_000_000, 7) == (7, 1428571, 28571428)
    assert oracle.partitioning_configuration(20000, 200000, 3) == (2, 10000, 66666)
    assert oracle.partitioning_configuration(3000, 20000, 3) == (2, 10000, 10000)


@pytest.mark.parametrize("P", [1, 2, 3, 32, 100, 256])
def test_radix_partition_invariants(oracle, P):
    """What partitionTable guarantees (src/RadixCluster/HashJoin.hpp:394-412): a permutation,
    partition p holds exactly the keys with Hash(id, P) == p, and input order is kept."""
    rel = _cases.tuples(_cases.splitmix64(5000, 8).astype(np.int64) % 700)
    for workers in (1, 4):
        out, bounds = oracle.radix_partition(rel, P, 0, SEED_P, workers=workers)
        assert bounds[0, 0] == 0 and bounds[-1, 1] == rel.shape[0]
        assert (bounds[1:, 0] == bounds[:-1, 1]).all()
        part = (oracle.hash_batch(0, SEED_P, out["id"]) % np.uint64(P)).astype(np.int64)
        for p in range(P):
            a, b = int(bounds[p, 0]), int(bounds[p, 1])
            assert (part[a:b] == p).all()
            assert (np.diff(out["payload"][a:b]) > 0).all()  # payload = input index: stable
        assert sorted(out["payload"].tolist()) == list(range(rel.shape[0]))


# ---- extension hashers: pinned to the canonical sources shipped in this image ----------------------
def _site():
    return sysconfig.get_paths()["purelib"]


def test_murmur3_against_canonical_source(oracle, tmp_path):
    src = os.path.join(_site(), "sklearn", "utils", "src", "MurmurHash3.cpp")
    if not os.path.exists(src):
        pytest.skip("scikit-learn's copy of Appleby's MurmurHash3.cpp is not in this image")
    so = tmp_path / "mm3.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-I", os.path.dirname(src), "-o", str(so), src], check=True)
    import ctypes as C
    lib = C.CDLL(str(so))
    fn = None
    for name in ("MurmurHash3_x64_128", "_Z19MurmurHash3_x64_128PKvijPv"):
        if hasattr(lib, name):
            fn = getattr(lib, name)
    assert fn is not None
    fn.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
    keys = _cases.splitmix64(2000, 31).astype(np.int64)
    for seed in (0, 1, 0xDEADBEEF, SEED_P):
        for k in keys[:500]:
            kk = C.c_int64(int(k))
            out = (C.c_uint64 * 2)()
            fn(C.byref(kk), 8, seed & 0xFFFFFFFF, out)
            assert oracle.hash_raw(1, int(k), seed) == out[0]


def test_city_against_libcudacxx(oracle, tmp_path):
    inc = os.path.join(_site(), "nvidia", "cuda_cccl", "include")
    if not os.path.exists(os.path.join(inc, "cuda", "std", "__functional", "hash.h")):
        pytest.skip("libcudacxx headers are not in this image")
    # The struct sits behind `#ifndef __cuda_std__`, so <cuda/std/functional> does not expose it to
    # host code; compile its text (murmur2 + CityHash64, verbatim from the header) on its own.
    text = open(os.path.join(inc, "cuda", "std", "__functional", "hash.h")).read()
    begin = text.index("template <class _Size>\n_LIBCUDACXX_HIDE_FROM_ABI _Size __loadword")
    end = text.index("template <class _Tp, size_t = sizeof(_Tp) / sizeof(size_t)>\nstruct __scalar_hash;")
    src = tmp_path / "city.cpp"
    src.write_text(
        "#include <cstring>\n#include <cstdint>\n#include <cstddef>\n#include <utility>\n"
        "#define _LIBCUDACXX_HIDE_FROM_ABI inline\n#define _LIBCUDACXX_DISABLE_UBSAN_UNSIGNED_INTEGER_CHECK\n"
        "#define _CCCL_FALLTHROUGH() [[fallthrough]]\n#define _CUDA_VSTD std\n"
        "namespace extracted { using std::pair; using std::size_t;\n" + text[begin:end] + "}\n"
        'extern "C" uint64_t city64(int64_t key) {\n'
        "  return extracted::__murmur2_or_cityhash<size_t, 64>()(&key, sizeof(key)); }\n")
    so = tmp_path / "city.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-o", str(so), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    import ctypes as C
    lib = C.CDLL(str(so))
    lib.city64.restype, lib.city64.argtypes = C.c_uint64, [C.c_int64]
    k2, mul, M = 0x9ae16a3b2f90404f, 0x9ddfea08eb382d69, 2**64 - 1

    def hash_len_16(u, v):
        a = ((u ^ v) * mul) & M
        a ^= a >> 47
        b = ((v ^ a) * mul) & M
        b ^= b >> 47
        return (b * mul) & M

    for k in _cases.splitmix64(500, 77).astype(np.int64).tolist() + [0, 1, -1, 2**63 - 1, -(2**63)]:
        plain = int(lib.city64(k))  # CityHash64(&key, 8), libc++ flavour
        for seed in (0, 5, SEED_P):
            # CityHash64WithSeed = HashLen16(CityHash64 - k2, seed)
            assert oracle.hash_raw(2, k, seed) == hash_len_16((plain - k2) & M, seed)


# ---- live cross-checks against the unmodified reference (authoring container / prebuilt .so) -------
def test_live_hash_and_generators(oracle, reference):
    keys = _cases.splitmix64(3000, 123).astype(np.int64)
    seeds = _cases.splitmix64(3000, 456)
    for k, s in zip(keys.tolist(), seeds.tolist()):
        assert oracle.hash_raw(0, k, s) == reference.xxh3(k, s)
    for alpha in (0.01, 0.5, 0.995, 1.0, 1.05, 1.25, 3.0):
        a = oracle.fill_zipf(30011, alpha, 5, 5000, 31337, 6)
        b = reference.fill_zipf(30011, alpha, 5, 5000, 31337, 6)
        assert (a["id"] == b["id"]).all() and (a["payload"] == b["payload"]).all(), alpha
    assert (oracle.fill_sequential(54321, -7)["id"] == reference.fill_sequential(54321, -7)["id"]).all()


def test_live_joins(oracle, reference):
    R = oracle.fill_sequential(50000, 1)
    S = oracle.fill_zipf(400000, 1.05, 1, 80000, 2024, 16)  # ~40 % of the key range misses R
    want = oracle.count_by_sort(R, S)
    assert 0 < want < S.shape[0]
    assert reference.join(R, S, 0, threads=3)["matches"] == want
    for P in (32, 1000, 2048):
        assert reference.join(R, S, 1, partitions=P, threads=3)["matches"] == want
        assert oracle.join_radix(R, S, P, workers=3) == want
    assert oracle.join_npj(R, S) == want
