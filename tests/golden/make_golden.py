"""Regenerates tests/golden/reference_vectors.json by running the UNMODIFIED reference
(oracle/_ref/libphj_ref.so, built from /root/reference by oracle/Makefile).

Run from the repo root in the authoring container:  python tests/golden/make_golden.py
The GPU box has no /root/reference; it only reads the committed JSON.

Zipf samples go through glibc pow(): the JSON is valid for this image (x86-64 glibc 2.39).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _cases  # noqa: E402
import _oracle  # noqa: E402

SEED_P, SEED_T = 0x9E3779B97F4A7C15, 1


def main():
    ref = _oracle.Reference()
    g = {"generated_by": "tests/golden/make_golden.py against oracle/_ref (unmodified reference)"}

    # --- XXH3 as XXHasher calls it (src/Common/XXHasher.hpp:19-22) ---
    keys = [0, 1, 2, 3, 10_000_000, 123456789, -1, 2**63 - 1, -(2**63)]
    seeds = [0, 1, SEED_P, 2**64 - 1]
    extra = _cases.splitmix64(64, 5).astype(np.int64).tolist()
    g["xxh3"] = [{"key": k, "seed": s, "hash": ref.xxh3(k, s)} for k in keys + extra for s in seeds]

    # --- LCG (src/Common/Random.cpp:9-30) ---
    g["lcg"] = {str(s): ref.lcg_doubles(s, 8) for s in (123456789, 1, 42, 2147483646)}

    # --- Zipf::generate (src/DataGenerator/Zipf.cpp:14-56) ---
    g["zipf"] = []
    for alpha, card, seed, n in [(1.05, 10_000_000, 123456789, 32), (1.25, 10_000_000, 123456789, 32),
                                 (0.01, 10_000_000, 123456789, 32), (0.99, 10, 123456789, 200),
                                 (1.0, 1000, 7, 64), (1.005, 1000, 7, 64), (0.995, 1000, 7, 64), (2.0, 100, 9, 64)]:
        g["zipf"].append({"alpha": alpha, "cardinality": card, "seed": seed,
                          "samples": ref.zipf_samples(alpha, card, seed, n)})

    # --- seeded fills ---
    z = ref.fill_zipf(1000, 1.05, 1, 100, 12345, 4)
    g["fill_zipf"] = {"n": 1000, "alpha": 1.05, "range": [1, 100], "base_seed": 12345, "batches": 4,
                      "ids": z["id"].tolist(), "payload_is_index": bool((z["payload"] == np.arange(1000)).all())}
    q = ref.fill_sequential(25000, 1)
    g["fill_sequential"] = {"n": 25000, "start": 1, "first": q["id"][:5].tolist(), "last": int(q["id"][-1]),
                            "payload_is_index": bool((q["payload"] == np.arange(25000)).all())}

    # --- join counts from the reference joiners ---
    joins = {}
    cases = dict(_cases.adversarial_cases())
    for name, (nr, ns, alpha, seed, batches) in _cases.GENERATOR_CASES.items():
        cases[name] = (ref.fill_sequential(nr, 1), ref.fill_zipf(ns, alpha, 1, nr, seed, batches))
    for name, (R, S) in cases.items():
        rec = {"n_build": int(R.shape[0]), "n_probe": int(S.shape[0])}
        if R.shape[0] == 0 or S.shape[0] == 0:
            # The reference is undefined here: zero tuples give zero tasks and its thread pool
            # terminates the process (observed: "terminate called without an active exception").
            joins[name] = dict(rec, matches=0, reference_undefined=True)
            continue
        # Radix worker split (src/RadixCluster/HashJoin.hpp:149-188, quirk SURVEY A.4.5): when one
        # relation is re-batched to MinBatchSize the worker count applies to both and a worker can
        # start beyond the smaller relation's end (out-of-bounds reads). One pool thread keeps
        # every adversarial case inside the reference's defined behaviour.
        small = min(R.shape[0], S.shape[0]) < 30000
        t = 1 if small else 3
        rec["npj_lp"] = ref.join(R, S, 0, table_kind=0, threads=3)["matches"]
        rec["npj_sc"] = ref.join(R, S, 0, table_kind=1, threads=3)["matches"]
        for P in (1, 3, 32, 64, 100, 2048):
            rec[f"radix_{P}"] = ref.join(R, S, 1, partitions=P, table_kind=0, threads=t,
                                         seed_partition=SEED_P, seed_table=SEED_T)["matches"]
        rec["radix_32_sc"] = ref.join(R, S, 1, partitions=32, table_kind=1, threads=t)["matches"]
        rec["radix_64_stock_hasher"] = ref.join(R, S, 1, partitions=64, threads=t, seeded=False)["matches"]
        counts = {v for k, v in rec.items() if k not in ("n_build", "n_probe")}
        assert len(counts) == 1, (name, rec)  # the count cannot depend on algorithm / fan-out / table
        joins[name] = {"n_build": rec["n_build"], "n_probe": rec["n_probe"], "matches": counts.pop()}
    g["joins"] = joins

    # --- hash tables: the reference's own test assertions (tests/NoPartitioningHashJoin/HashTableTest.hpp) ---
    tables = []
    one = _cases.tuples([123456789])
    one["payload"] = 987654321
    dup = _cases.tuples(np.full(10, 123456789))
    many = _cases.sequential(1000, 0)
    for kind, ratio in ((0, 1 / 0.75), (1, 0.3), (0, 0.0), (1, 0.0)):
        for label, build, nobj, probe, threads in (("InsertGetAndExists", one, 10, [123456789, 5], 1),
                                                   ("Iterator", dup, 10, [123456789, 1], 1),
                                                   ("MultiThreadedInsert", many, 1000, list(range(0, 1000, 37)) + [1000, -1], 4)):
            if kind == 1 and ratio == 0.3 and label == "MultiThreadedInsert":
                ratio_used = 0.1  # HashTableTest.hpp:143
            else:
                ratio_used = ratio
            buckets, ex, gi, ga = ref.table_probe(kind, ratio_used, 5, nobj, build, probe, insert_threads=threads)
            tables.append({"kind": kind, "ratio": ratio_used, "test": label, "seed": 5, "objects": nobj,
                           "probe": probe, "buckets": buckets, "exists": ex, "getall": ga,
                           "get_index": gi if threads == 1 else None})
    g["tables"] = tables

    # --- JSON rendering (src/Common/Results.hpp:262-279) ---
    params = {"PrimaryRelationSize": "10000000", "SecondaryRelationSize": "200000000", "Skew": "1.050000",
              "Type": "RadixParitioning", "NumberOfPartitions": "32"}
    g["json"] = {"parameters": params, "partition_ns": 666_000_000, "build_ns": 120_400_000, "probe_ns": 572_999_999,
                 "ms": ref.format_json(params, 666_000_000, 120_400_000, 572_999_999, "ms"),
                 "us": ref.format_json(params, 666_000_000, 120_400_000, 572_999_999, "us")}

    out = os.path.join(HERE, "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
