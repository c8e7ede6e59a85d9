"""GPU check of the opt-in experiments (not collected by pytest; run it first thing before measuring them):

    python tools/check_experiments.py

PHJ_L2JOIN=1 (L2-resident partition tables) and PHJ_SCAT_PREFETCH=k (L2 prefetch in the scatter) must
give the same match count as the default path on every golden / adversarial case, and the same
partitioned relations byte for byte (the prefetch does not touch results; the L2 join only replaces
the build + probe kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases  # noqa: E402
import partitionedhashjoin_b200 as phj  # noqa: E402


def run(R, S, env, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        with phj.Engine("radix-partitioning", **kw) as e:
            e.upload(R, S)
            res = e.join()
            parts = e.read_partitions(1, kw["partitions"]) if kw.get("partitions") else None
            return res, parts
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def main():
    cases = dict(_cases.adversarial_cases())
    big_r = _cases.tuples(_cases.splitmix64(300_000, 51).astype(np.int64) % 200_003)
    big_s = _cases.tuples(_cases.splitmix64(2_000_000, 52).astype(np.int64) % 300_007)
    cases["big_random"] = (big_r, big_s)
    R = np.empty(2_000_000, dtype=phj.TUPLE_DTYPE)
    S = np.empty(30_000_000, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, 1.25, 1, 4_000_000, 99, 16)  # half of the key range misses
    cases["gen_2m_30m_zipf125"] = (R, S)
    bad = 0
    for name, (R, S) in cases.items():
        for kw in ({"partitions": 64}, {"partitions": 256, "radix_bits": (8, 0)}, {"partitions": 4096},
                   {"partitions": 1000}, {"partitions": 0}):
            want, wparts = run(R, S, {}, **kw)
            for env in ({"PHJ_L2JOIN": "1"}, {"PHJ_SCAT_PREFETCH": "1"}, {"PHJ_SCAT_PREFETCH": "2", "PHJ_L2JOIN": "1"}):
                got, gparts = run(R, S, env, **kw)
                ok = got["matches"] == want["matches"]
                if wparts is not None:
                    ok = ok and (gparts[0] == wparts[0]).all() and (gparts[1] == wparts[1]).all()
                if not ok:
                    bad += 1
                    print(f"MISMATCH {name} {kw} {env}: {got['matches']} != {want['matches']}")
    print("experiments agree with the default path" if not bad else f"{bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
