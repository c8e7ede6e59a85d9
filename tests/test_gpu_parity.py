"""GPU parity tests (run on the B200 with -m gpu): everything goes through the C ABI
(libphj_b200.so) and is compared with the oracle (oracle/phj_oracle.c) on the same seeded inputs,
with the golden vectors made by the unmodified reference, and -- at BASELINE.json's full sizes --
through size-independent properties.
"""
import json
import os

import numpy as np
import pytest

import _cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
SEED_P = 0x9E3779B97F4A7C15
HASHES = ["xxh3", "murmur3", "city"]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(phj):
    assert phj.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    info = phj.device_info(0)
    print("device:", info)
    assert info["cc_major"] >= 10


def cases_for(phj, name):
    adv = _cases.adversarial_cases()
    if name in adv:
        return adv[name]
    nr, ns, alpha, seed, batches = _cases.GENERATOR_CASES[name]
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, alpha, 1, nr, seed, batches)
    return R, S


class pass2_histogram:
    """How pass 2 of a two-pass plan gets its histogram: 'full' = from the one 12-bit histogram read
    (radix_histogram_full, the default), 'fused' = counted by the pass-1 scatter (PHJ_FLAG_NO_HIST12),
    'separate' = its own histogram read (PHJ_FLAG_NO_FUSE_HIST2). Yields the engine flags."""
    MODES = ("full", "fused", "separate")

    def __init__(self, phj, mode):
        self.mode = mode
        self.flags = {"full": 0, "fused": phj.FLAG_NO_HIST12, "separate": phj.FLAG_NO_FUSE_HIST2}[mode]

    def __enter__(self):
        return self.flags

    def __exit__(self, *exc):
        pass


def run(phj, R, S, algo, **kw):
    with phj.Engine(algo, **kw) as e:
        e.upload(R, S)
        return e.join()


# ---- K8: device hashers are bit-exact ------------------------------------------------------------
@pytest.mark.parametrize("hash_id", [0, 1, 2])
def test_device_hashers_bit_exact(phj, oracle, hash_id):
    keys = np.concatenate([_cases.splitmix64(200000, 5).astype(np.int64),
                           np.array([0, 1, 2, 3, -1, 2**63 - 1, -(2**63), 10_000_000, 123456789], dtype=np.int64)])
    for seed in (0, 1, SEED_P, 2**64 - 1):
        assert (phj.hash_batch(hash_id, seed, keys) == oracle.hash_batch(hash_id, seed, keys)).all()


def test_device_xxh3_golden(phj):
    keys = np.array([r["key"] for r in GOLDEN["xxh3"]], dtype=np.int64)
    for seed in sorted({r["seed"] for r in GOLDEN["xxh3"]}):
        sel = [i for i, r in enumerate(GOLDEN["xxh3"]) if r["seed"] == seed]
        got = phj.hash_batch(0, seed, keys[sel])
        assert got.tolist() == [GOLDEN["xxh3"][i]["hash"] for i in sel]


# ---- counts: golden (reference) + oracle ----------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GOLDEN["joins"]))
def test_counts_match_reference_golden(phj, oracle, name):
    R, S = cases_for(phj, name)
    want = GOLDEN["joins"][name]["matches"]
    assert oracle.count_by_sort(R, S) == want
    assert run(phj, R, S, "no-partitioning")["matches"] == want
    assert run(phj, R, S, "no-partitioning", flags=phj.FLAG_COOP_PROBE)["matches"] == want  # four lanes per probe
    for P in (0, 1, 2, 3, 32, 64, 100, 1000, 2048, 4096, 65536):
        res = run(phj, R, S, "radix-partitioning", partitions=P)
        assert res["matches"] == want, (name, P, res)
    for bits in ((6, 6), (8, 4), (3, 8), (8, 8), (5, 0), (1, 1)):
        res = run(phj, R, S, "radix-partitioning", partitions=1 << sum(bits), radix_bits=bits)
        assert res["matches"] == want, (name, bits, res)
    for flags in (phj.FLAG_NO_TMA_STORE,):
        res = run(phj, R, S, "radix-partitioning", partitions=4096, flags=flags)
        assert res["matches"] == want, (name, flags, res)
    # per-partition tables in global memory (L2-resident) instead of shared memory, any fan-out
    for P in (1, 2, 32, 64, 100, 256, 4096):
        res = run(phj, R, S, "radix-partitioning", partitions=P, flags=phj.FLAG_L2_TABLES)
        assert res["matches"] == want, (name, P, "l2 tables", res)


@pytest.mark.parametrize("hash", HASHES)
@pytest.mark.parametrize("name", ["dup_build_keys", "extreme_keys", "random64", "gen_small_105"])
def test_counts_every_hasher(phj, name, hash):
    R, S = cases_for(phj, name)
    want = GOLDEN["joins"][name]["matches"]
    for seed in (0, 1, SEED_P):
        assert run(phj, R, S, "no-partitioning", hash=hash, hash_seed=seed)["matches"] == want
        assert run(phj, R, S, "radix-partitioning", partitions=256, hash=hash, hash_seed=seed)["matches"] == want
        assert run(phj, R, S, "radix-partitioning", partitions=77, hash=hash, hash_seed=seed)["matches"] == want


def test_empty_relations(phj):
    R, S = _cases.sequential(100), _cases.tuples(np.arange(50, 150))
    empty = _cases.tuples(np.empty(0, dtype=np.int64))
    for algo, kw in (("no-partitioning", {}), ("radix-partitioning", {"partitions": 64}),
                     ("radix-partitioning", {"partitions": 4096})):
        assert run(phj, R, empty, algo, **kw)["matches"] == 0
        assert run(phj, empty, S, algo, **kw)["matches"] == 0  # deviation: the reference NPJ throws here
        assert run(phj, empty, empty, algo, **kw)["matches"] == 0
        assert run(phj, R, S, algo, **kw)["matches"] == 51  # keys 50..100


def test_reuse_handle_and_reupload(phj, oracle):
    """One handle, several relations of different sizes, repeated joins."""
    with phj.Engine("radix-partitioning", partitions=512) as e:
        for n_r, n_s, seed in ((5000, 40000, 1), (70000, 10000, 2), (123, 456789, 3), (5000, 40000, 4)):
            R = _cases.tuples(_cases.splitmix64(n_r, seed).astype(np.int64) % 9000)
            S = _cases.tuples(_cases.splitmix64(n_s, seed + 100).astype(np.int64) % 12000)
            want = oracle.count_by_sort(R, S)
            e.upload(R, S)
            assert e.join()["matches"] == want
            assert e.join()["matches"] == want
            assert e.join_host(R, S)["matches"] == want


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("algo,kw", [("radix-partitioning", {"partitions": 512}), ("radix-partitioning", {"partitions": 0}),
                                     ("radix-partitioning", {"partitions": 100}), ("no-partitioning", {}),
                                     ("no-partitioning", {"flags": 0x10})])
def test_join_host_streamed(phj, oracle, algo, kw, pinned):
    """phj_join_host with the probe relation uploaded in chunks, each joined as it lands: the sum of
    the chunk counts is the count of the whole join, from pageable and from pinned memory, for every
    chunk count (also more chunks than tuples, ragged last chunks, an empty probe relation), and the
    relations are resident afterwards (a plain phj_join gives the same count)."""
    R = _cases.tuples(_cases.splitmix64(30011, 5).astype(np.int64) % 50000)
    S = _cases.tuples(_cases.splitmix64(400009, 6).astype(np.int64) % 70000)
    want = oracle.count_by_sort(R, S)
    assert 0 < want < S.shape[0]
    keep = []
    if pinned:
        keep = [phj.PinnedTuples(R.shape[0]), phj.PinnedTuples(S.shape[0])]
        keep[0].array[:] = R
        keep[1].array[:] = S
        R, S = keep[0].array, keep[1].array
    for chunks in (1, 2, 3, 7, 32):
        with phj.Engine(algo, upload_chunks=chunks, **kw) as e:
            for _ in range(2):  # the second call reuses the child handle, streams and events
                res = e.join_host(R, S)
                assert res["matches"] == want, (chunks, res)
                assert res["upload_chunks"] == chunks and res["h2d_bytes"] == 16 * (R.shape[0] + S.shape[0])
                assert res["e2e_ns"] >= res["h2d_ns"] > 0
            assert e.join()["matches"] == want          # both relations are resident
            # other sizes on the same handle: fewer tuples than chunks, nothing to probe, a larger relation
            assert e.join_host(R, S[:5])["matches"] == oracle.count_by_sort(R, S[:5])
            assert e.join_host(R, S[:0])["matches"] == 0
            S2 = np.concatenate([S, S[:1000]])
            assert e.join_host(R[:777], S2)["matches"] == oracle.count_by_sort(R[:777], S2)
    # automatic chunking leaves a probe relation of a few MB alone
    with phj.Engine(algo, **kw) as e:
        res = e.join_host(R, S)
        assert res["matches"] == want and res["upload_chunks"] == 1


def test_join_host_streamed_generator_data(phj):
    """Streamed host join at 2 M x 40 M generator tuples (Zipf 1.05: every probe key has a match), the
    automatic chunk count, pinned memory: what bench.py's e2e leg does, at a fifth of its size."""
    Rp, Sp = phj.PinnedTuples(2_000_000), phj.PinnedTuples(40_000_000)
    phj.fill_sequential(Rp.array, 1)
    phj.fill_zipf(Sp.array, 1.05, 1, 2_000_000, 777, 16)
    with phj.Engine("radix-partitioning", partitions=1024) as e:
        res = e.join_host(Rp.array, Sp.array)
        assert res["matches"] == 40_000_000 and res["upload_chunks"] == 3, res
        plain = e.join()
        assert plain["matches"] == 40_000_000
    with phj.Engine("no-partitioning") as e:
        res = e.join_host(Rp.array, Sp.array)
        assert res["matches"] == 40_000_000 and res["upload_chunks"] == 3, res
    # the joiner interface with the same option
    j = phj.RadixClusteringHashJoiner(phj.RadixClusteringConfiguration(NumberOfPartitions=1024), stream_upload=True)
    timer = phj.HashJoinTimer()
    assert j.Run(Rp.array, Sp.array, timer).shape[0] == 0
    assert j.last_result["matches"] == 40_000_000 and j.last_result["upload_chunks"] == 3
    assert timer.GetResult().partitioning_ns > 0
    j.close()


# ---- intermediate state: the partitioned relations equal the reference algorithm's ---------------
def assert_same_partitioning(got, gb, want, wb, n, exact_order):
    """Boundaries identical; partition contents identical -- in input order when `exact_order`,
    else as multisets (the payload is the input index, so sorting by it inside each partition
    restores input order)."""
    assert gb[:-1].tolist() == wb[:, 0].tolist() and int(gb[-1]) == n
    if not exact_order:
        part_of = np.repeat(np.arange(wb.shape[0]), (wb[:, 1] - wb[:, 0]).astype(np.int64))
        order = np.lexsort((got["payload"], part_of))
        got = got[order]
    assert (got["id"] == want["id"]).all() and (got["payload"] == want["payload"]).all()


@pytest.mark.parametrize("hash_id,hash", list(enumerate(HASHES)))
@pytest.mark.parametrize("P,bits", [(64, (0, 0)), (4096, (6, 6)), (2048, (0, 0)), (256, (8, 0)), (256, (4, 4)),
                                    (100, (0, 0)), (1000, (0, 0)), (65536, (8, 8)), (1, (0, 0)), (8192, (7, 6))])
def test_partition_layout_equals_oracle(phj, oracle, P, bits, hash_id, hash):
    """partitionedTable + PartitionsInfo (src/RadixCluster/HashJoin.hpp:16-33,195-198,394-412):
    partition p = Hash(id, P) at [bounds[p], bounds[p+1]), input order kept: the partitioned
    relations are bit-identical to the reference algorithm's output."""
    R = _cases.tuples(_cases.splitmix64(30011, 9).astype(np.int64) % 5003)
    S = _cases.tuples(_cases.splitmix64(250007, 10).astype(np.int64) % 7001)
    want_count = oracle.count_by_sort(R, S)
    for mode, extra in (("full", 0), ("full", phj.FLAG_NO_TMA_STORE), ("fused", 0), ("separate", 0)):
        with pass2_histogram(phj, mode) as flags, \
                phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, hash=hash, hash_seed=SEED_P,
                           flags=flags | extra) as e:
            e.upload(R, S)
            assert e.join()["matches"] == want_count
            for which, rel in ((0, R), (1, S)):
                got, gb = e.read_partitions(which, P)
                want, wb = oracle.radix_partition(rel, P, hash_id, SEED_P, workers=1)
                assert_same_partitioning(got, gb, want, wb, rel.shape[0], True)


@pytest.mark.parametrize("P,bits", [(256, (4, 4)), (4096, (6, 6)), (4096, (5, 7)), (1 << 14, (8, 6)), (64, (6, 0))])
def test_prepartitioned_bind_equals_plain_join(phj, oracle, P, bits):
    """phj_bind_device_partitioned: a SHARD_SPLIT by the plan's pass-1 digit (what the multi-GPU
    shuffle delivers) followed by a join that starts at pass 2 gives the same count AND bit-identical
    final partitions as the ordinary two-pass join -- i.e. as the reference's partitionTable."""
    R = _cases.tuples(_cases.splitmix64(120_000, 31).astype(np.int64) % 90_001)
    S = _cases.tuples(np.where(_cases.splitmix64(1_500_000, 32) % np.uint64(10) < 2, 77,
                               _cases.splitmix64(1_500_000, 33) % np.uint64(130_003)).astype(np.int64))
    b1, b2 = bits
    with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, hash_seed=SEED_P) as plain, \
            phj.Engine("shard-split", partitions=1 << b1, shard_shift=b2, hash_seed=SEED_P) as split, \
            phj.Engine("radix-partitioning", partitions=P, radix_bits=bits, hash_seed=SEED_P) as pre:
        plain.upload(R, S)
        want = plain.join()
        assert want["matches"] == oracle.count_by_sort(R, S)
        split.upload(R, S)
        split.join()
        ptr, n, bnd = [], [], []
        for which in (0, 1):
            p_, _, n_ = split.device_partitions(which)
            ptr.append(p_), n.append(n_), bnd.append(split.read_bounds(which, 1 << b1))
        pre.bind_device_partitioned(ptr[0], n[0], ptr[1], n[1], bnd[0], bnd[1])
        for _ in range(2):  # reusable
            got = pre.join()
            assert got["matches"] == want["matches"]
            assert got["passes"] == want["passes"] - 1
        for which in (0, 1):
            a, ab = plain.read_partitions(which, P)
            b, bb = pre.read_partitions(which, P)
            assert (ab == bb).all()
            assert (a["id"] == b["id"]).all() and (a["payload"] == b["payload"]).all()
        # a plan whose pass 1 has a different fan-out is rejected
        with pytest.raises(phj.PhjError):
            pre.bind_device_partitioned(ptr[0], n[0], ptr[1], n[1], bnd[0][:-1], bnd[1][:-1])


# ---- bucket-chained table ---------------------------------------------------
@pytest.mark.parametrize("hash", HASHES)
def test_chained_table_counts(phj, oracle, hash):
    """No-partitioning join through the bucket-chained table: the oracle's separate-chaining count
    (= the linear-probing count = the independent sort count) on every case, incl. 5000 equal build
    keys (one long chain built under contention) and the int64 extremes (no reserved key value)."""
    cases = dict(_cases.adversarial_cases())
    for name in ("gen_small_105", "gen_ragged"):
        cases[name] = cases_for(phj, name)
    big_r = _cases.tuples(_cases.splitmix64(300_000, 51).astype(np.int64) % 200_003)
    big_s = _cases.tuples(_cases.splitmix64(2_000_000, 52).astype(np.int64) % 300_007)
    cases["big_random"] = (big_r, big_s)
    for name, (R, S) in cases.items():
        want = oracle.count_by_sort(R, S)
        if R.shape[0]:
            assert oracle.join_npj(R, S, table_kind=1) == want
        got = run(phj, R, S, "no-partitioning", hash=hash, flags=phj.FLAG_CHAINED_TABLE)
        assert got["matches"] == want, (name, got["matches"], want)
    j = phj.NoPartitioningHashJoiner(table="separate-chaining")
    j.Run(*cases["big_random"])
    assert j.last_result["matches"] == oracle.count_by_sort(*cases["big_random"])
    j.close()
    with pytest.raises(phj.PhjError):
        phj.Engine("radix-partitioning", flags=phj.FLAG_CHAINED_TABLE)


# ---- device-side generators ------------------------------------------------------
@pytest.mark.parametrize("alpha", [0.01, 1.05, 1.25])
def test_device_generators(phj, oracle, alpha):
    """Same algorithm, seeds and batch split as the host generators; only pow() differs (CUDA vs
    glibc), so nearly every sample is identical and all of them lie in the key range."""
    nr, ns, batches = 50_000, 1_000_000, 500
    dR = phj.DeviceTuples(nr).fill_sequential(1)
    dS = phj.DeviceTuples(ns).fill_zipf(alpha, 1, nr, 4242, batches)
    R, S = dR.download(), dS.download()
    hostR = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    hostS = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(hostR, 1)
    phj.fill_zipf(hostS, alpha, 1, nr, 4242, batches)
    assert (R["id"] == hostR["id"]).all() and (R["payload"] == hostR["payload"]).all()
    assert (S["payload"] == np.arange(ns)).all()
    assert S["id"].min() >= 1 and S["id"].max() <= nr
    same = float((S["id"] == hostS["id"]).mean())
    assert same > 0.999, same
    with phj.Engine("radix-partitioning", partitions=256) as e:  # joined in place, no upload
        e.bind_device(dR.ptr, nr, dS.ptr, ns, keepalive=(dR, dS))
        assert e.join()["matches"] == ns == oracle.count_by_sort(R, S)
    with pytest.raises(phj.PhjError):
        phj.DeviceTuples(10).fill_zipf(0.001, 1, 10, 1, 1)
    dR.close(), dS.close()


# ---- the joined table -----------------------------------------------------------
@pytest.mark.parametrize("name", sorted(_cases.adversarial_cases()) + ["gen_small_105", "gen_ragged"])
def test_materialized_join_equals_oracle(phj, oracle, name):
    """phj_join_materialize: every (probe tuple, equal-key build tuple) row, as a multiset, equals
    the oracle's GetAll join; matches stays the count-only figure."""
    R, S = cases_for(phj, name)
    want = _cases.sorted_rows(oracle.join_materialize(R, S))
    for P, bits in ((0, (0, 0)), (64, (0, 0)), (1024, (5, 5)), (4, (2, 0))):
        with phj.Engine("radix-partitioning", partitions=P, radix_bits=bits) as e:
            e.upload(R, S)
            for _ in range(2):
                res = e.join_materialize()
                assert res["joined_tuples"] == want.shape[0], (name, P)
                assert res["matches"] == oracle.count_by_sort(R, S)
                got = _cases.sorted_rows(e.read_joined())
                assert (got == want).all(), (name, P)
            if want.shape[0] > 10:
                assert (_cases.sorted_rows(e.read_joined(3, 5)).shape[0]) == 5


def test_materialized_join_large_partitions_and_duplicates(phj, oracle):
    """Build partitions far beyond one shared-memory table chunk (block nested loop over chunks)
    and a build key repeated 30 000 times (a 30 000-row fan-out per matching probe tuple)."""
    r = np.concatenate([_cases.splitmix64(60_000, 41).astype(np.int64) % 50_000, np.full(30_000, 123)])
    s = np.concatenate([_cases.splitmix64(200_000, 42).astype(np.int64) % 80_000, np.full(7, 123)])
    R, S = _cases.tuples(r), _cases.tuples(s)
    want = _cases.joined_reference(R, S)
    with phj.Engine("radix-partitioning", partitions=2) as e:
        e.upload(R, S)
        res = e.join_materialize()
        assert res["joined_tuples"] == want.shape[0]
        assert (_cases.sorted_rows(e.read_joined()) == want).all()


def test_joiner_run_returns_the_joined_table(phj, oracle):
    """RadixClusteringHashJoiner(materialize=True).Run returns the filled Table<JoinedTuple>."""
    R, S = cases_for(phj, "gen_small_125")
    j = phj.RadixClusteringHashJoiner(phj.RadixClusteringConfiguration(NumberOfPartitions=256), materialize=True)
    table = j.Run(R, S)
    assert table.dtype == phj.JOINED_DTYPE and table.shape[0] == S.shape[0]
    assert (_cases.sorted_rows(table) == _cases.sorted_rows(oracle.join_materialize(R, S))).all()
    j.close()
    with phj.Engine("no-partitioning") as e:
        e.upload(R, S)
        with pytest.raises(phj.PhjError):
            e.join_materialize()


def test_partition_layout_skewed_generator_data(phj, oracle):
    nr, ns = 100000, 1500000
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, 1.25, 1, nr, 4711, 16)
    want, wb = oracle.radix_partition(S, 4096, 0, SEED_P, workers=1)
    for mode, extra in (("full", 0), ("full", phj.FLAG_NO_TMA_STORE), ("fused", 0), ("separate", 0)):
        with pass2_histogram(phj, mode) as flags, \
                phj.Engine("radix-partitioning", partitions=4096, hash_seed=SEED_P, flags=flags | extra) as e:
            e.upload(R, S)
            assert e.join()["matches"] == ns
            got, gb = e.read_partitions(1, 4096)
            assert_same_partitioning(got, gb, want, wb, ns, True)
            assert e.join()["matches"] == ns   # deterministic: a second run gives the very same bytes
            again, _ = e.read_partitions(1, 4096)
            assert (again["id"] == got["id"]).all() and (again["payload"] == got["payload"]).all()


def test_oversize_partitions_use_the_global_table(phj, oracle):
    """|R|/P far above the shared-memory table: those partitions go through gt_build/gt_probe."""
    R = _cases.tuples(_cases.splitmix64(400000, 1).astype(np.int64) % 1000003)
    S = _cases.tuples(_cases.splitmix64(900000, 2).astype(np.int64) % 1000003)
    want = oracle.count_by_sort(R, S)
    res = run(phj, R, S, "radix-partitioning", partitions=8)
    assert res["matches"] == want and res["fallback_partitions"] == 8
    # one giant build partition next to small ones: 50 000 copies of one key share a partition
    R2 = _cases.tuples(np.concatenate([np.arange(1, 20001), np.full(50000, 7)]))
    S2 = _cases.tuples(_cases.splitmix64(600000, 5).astype(np.int64) % 30000)
    res = run(phj, R2, S2, "radix-partitioning", partitions=64)
    assert res["matches"] == oracle.count_by_sort(R2, S2) and res["fallback_partitions"] == 1


# ---- full size (BASELINE.json configs): size-independent properties ------------------------------
FULL = [(10_000_000, 200_000_000)]


@pytest.fixture(scope="module")
def full_relations(phj):
    nr, ns = FULL[0]
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, 1.05, 1, nr, 12345, 64)
    return R, S


def test_full_size_counts(phj, full_relations):
    """10 M x 200 M (configs[0]/[1]): generator data joins to |S| (every probe key lies in
    [1, |R|]); NPJ == radix for every fan-out/hasher; shrinking R to its first
    half must give exactly the number of probe keys <= |R|/2 (counted independently with numpy)."""
    R, S = full_relations
    ns = S.shape[0]
    half = R.shape[0] // 2
    want_half = int((S["id"] <= half).sum())
    with phj.Engine("no-partitioning") as e:
        e.upload(R, S)
        assert e.join()["matches"] == ns
        e.upload(R[:half], S)
        assert e.join()["matches"] == want_half
    for kw in ({"partitions": 0}, {"partitions": 2048}, {"partitions": 4096, "hash": "murmur3"},
               {"partitions": 4096, "hash": "city"}, {"partitions": 1024}, {"partitions": 32},
               {"partitions": 8192, "radix_bits": (7, 6)}):
        with phj.Engine("radix-partitioning", **kw) as e:
            e.upload(R, S)
            res = e.join()
            assert res["matches"] == ns, (kw, res)
            e.upload(R[:half], S)
            assert e.join()["matches"] == want_half, kw


def test_full_size_partition_properties(phj, oracle, full_relations):
    """At full size the partitioned probe relation is a permutation of S (checksums), every
    partition holds only its own hash class, and order inside partitions is input order."""
    R, S = full_relations
    P = 4096
    with phj.Engine("radix-partitioning", partitions=P, hash_seed=SEED_P) as e:
        e.upload(R, S)
        assert e.join()["matches"] == S.shape[0]
        got, bounds = e.read_partitions(1, P)
    check_full_size_partitioning(oracle, S, got, bounds, P, stable=True)


@pytest.mark.parametrize("alpha", [0.01, 1.25])
def test_full_size_counts_other_skews(phj, alpha):
    """BASELINE.json configs[1] ("uniform" = the reference generator's minimum skew 0.01) and configs[3] (Zipf 1.25,
    key 1 = 22 % of S) at the full 10 M x 200 M: |S| on the whole build side, and on half of it exactly the number
    of probe keys <= |R| / 2 (numpy), through the two-pass plan, the one-pass plan with L2-resident tables and the
    no-partitioning join."""
    nr, ns = FULL[0]
    R = np.empty(nr, dtype=phj.TUPLE_DTYPE)
    S = np.empty(ns, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, alpha, 1, nr, 12345, 64)
    half = nr // 2
    want_half = int((S["id"] <= half).sum())
    assert 0 < want_half < ns
    for algo, kw in (("radix-partitioning", {"partitions": 4096}),
                     ("radix-partitioning", {"partitions": 64, "flags": phj.FLAG_L2_TABLES}),
                     ("radix-partitioning", {"partitions": 8, "flags": phj.FLAG_L2_TABLES}),
                     ("no-partitioning", {})):
        with phj.Engine(algo, **kw) as e:
            e.upload(R, S)
            assert e.join()["matches"] == ns, (algo, kw)
            e.upload(R[:half], S)
            assert e.join()["matches"] == want_half, (algo, kw)


def check_full_size_partitioning(oracle, S, got, bounds, P, stable):
    assert int(bounds[0]) == 0 and int(bounds[-1]) == S.shape[0] and (np.diff(bounds.astype(np.int64)) >= 0).all()
    # permutation: payload is the input index, so the sorted payloads are 0..n-1 and ids follow
    assert int(got["payload"].sum(dtype=np.uint64)) == int(S["payload"].sum(dtype=np.uint64))
    assert int((got["id"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).sum(dtype=np.uint64)) == \
        int((S["id"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).sum(dtype=np.uint64))
    assert (S["id"][got["payload"]] == got["id"]).all()
    # hash class + stability on a sample of partitions (the hot one included)
    sizes = np.diff(bounds.astype(np.int64))
    for p in [0, 1, P // 2, P - 1, int(sizes.argmax()), int(sizes.argmin())]:
        a, b = int(bounds[p]), int(bounds[p + 1])
        part = oracle.hash_batch(0, SEED_P, got["id"][a:b]) % np.uint64(P)
        assert (part == p).all()
        if stable:
            assert (np.diff(got["payload"][a:b]) > 0).all()
