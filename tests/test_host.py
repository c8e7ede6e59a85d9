"""CPU tests of the product's host side: the C-ABI library loads and exports everything
include/phj.h declares, the host-compiled hashers and generators agree with the oracle, and
configuration errors behave like the reference's. No compute entry point is called without a GPU.
"""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
SEED_P = 0x9E3779B97F4A7C15


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "phj.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phj_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(phj):
    from partitionedhashjoin_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 18
    raw = C.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} is declared in include/phj.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert phj._lib.lib.phj_abi_version() == 5


def test_flag_values_match_the_header(phj):
    """Every PHJ_FLAG_* of include/phj.h has a ctypes-side constant of the same value, and no two flags share a bit."""
    from partitionedhashjoin_b200 import _lib
    text = open(os.path.join(ROOT, "include", "phj.h")).read()
    flags = {name: int(value, 16) for name, value in re.findall(r"#define\s+PHJ_(FLAG_[A-Z0-9_]+)\s+0x([0-9a-fA-F]+)u", text)}
    assert len(flags) >= 12
    for name, value in flags.items():
        assert getattr(_lib, name) == value, name
    assert len(set(flags.values())) == len(flags) and all(v & (v - 1) == 0 for v in flags.values())


def test_struct_layouts_match_the_header(phj, tmp_path):
    """sizeof/offsetof as the C compiler sees include/phj.h == the ctypes mirrors."""
    from partitionedhashjoin_b200 import _lib
    src = tmp_path / "layout.c"
    fields = {"phj_config": [f for f, _ in _lib.PhjConfig._fields_],
              "phj_result": [f for f, _ in _lib.PhjResult._fields_],
              "phj_device_info": [f for f, _ in _lib.PhjDeviceInfo._fields_],
              "phj_dist_layout": [f for f, _ in _lib.PhjDistLayout._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/phj.h"', "int main(void){"]
    for st, fs in fields.items():
        body.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            body.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    body.append('printf("phj_tuple %zu\\n", sizeof(phj_tuple)); return 0;}')
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirrors = {"phj_config": _lib.PhjConfig, "phj_result": _lib.PhjResult, "phj_device_info": _lib.PhjDeviceInfo,
               "phj_dist_layout": _lib.PhjDistLayout}
    for st, cls in mirrors.items():
        assert int(out[st]) == C.sizeof(cls)
        for f, _ in cls._fields_:
            assert int(out[f"{st}.{f}"]) == getattr(cls, f).offset, (st, f)
    assert int(out["phj_tuple"]) == 16 == phj.TUPLE_DTYPE.itemsize


@pytest.mark.parametrize("hash_id", [0, 1, 2])
def test_host_hashers_equal_oracle(phj, oracle, hash_id):
    keys = np.concatenate([_cases.splitmix64(4000, 17).astype(np.int64),
                           np.array([0, 1, -1, 2**63 - 1, -(2**63), 123456789], dtype=np.int64)])
    for seed in (0, 1, SEED_P, 2**64 - 1, 0xDEADBEEF):
        want = oracle.hash_batch(hash_id, seed, keys)
        got = np.array([phj.hash_host(hash_id, seed, int(k)) for k in keys], dtype=np.uint64)
        assert (got == want).all()


def test_host_xxh3_golden(phj):
    for rec in GOLDEN["xxh3"]:
        assert phj.hash_host(0, rec["seed"], rec["key"]) == rec["hash"]


def test_generators_equal_oracle_and_golden(phj, oracle):
    f = GOLDEN["fill_zipf"]
    z = np.empty(f["n"], dtype=phj.TUPLE_DTYPE)
    phj.fill_zipf(z, f["alpha"], f["range"][0], f["range"][1], f["base_seed"], f["batches"], threads=3)
    assert z["id"].tolist() == f["ids"] and (z["payload"] == np.arange(f["n"])).all()
    for alpha, n, lo, hi, seed, batches in [(1.05, 100003, 1, 20000, 12345, 64), (1.25, 50000, 1, 777, 9, 1),
                                            (0.01, 30000, -5, 5000, 3, 7), (0.995, 20000, 10, 20, 8, 3),
                                            (1.0, 20000, 1, 10**7, 21, 5)]:
        a = np.empty(n, dtype=phj.TUPLE_DTYPE)
        phj.fill_zipf(a, alpha, lo, hi, seed, batches, threads=4)
        b = oracle.fill_zipf(n, alpha, lo, hi, seed, batches)
        assert (a["id"] == b["id"]).all() and (a["payload"] == b["payload"]).all(), alpha
    q = np.empty(70001, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(q, start=-3, threads=5)
    w = oracle.fill_sequential(70001, -3)
    assert (q["id"] == w["id"]).all() and (q["payload"] == w["payload"]).all()


def test_generator_argument_errors(phj):
    a = np.empty(10, dtype=phj.TUPLE_DTYPE)
    with pytest.raises(ValueError):
        phj.fill_zipf(a, 0.001, 1, 10, 5)        # src/DataGenerator/Zipf.cpp:18-20
    with pytest.raises(ValueError):
        phj.fill_zipf(a, 1.05, 10, 10, 5)        # src/DataGenerator/Zipf.cpp:61-67
    with pytest.raises(ValueError):
        phj.fill_zipf(a, 1.05, 1, 10, 2147483647)  # LCG fixed point


def test_configuration_errors(phj):
    from partitionedhashjoin_b200 import _lib
    cfg, h = _lib.PhjConfig(), C.c_void_p()
    cfg.algo = 7
    assert _lib.lib.phj_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    assert b"Unrecognized join algorithm" in _lib.lib.phj_last_error()
    cfg.algo, cfg.hash = 1, 9
    assert _lib.lib.phj_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    # --partitions is only legal with the radix join (src/Arguments.hpp:12-17)
    cfg.algo, cfg.hash, cfg.partitions = 0, 0, 32
    assert _lib.lib.phj_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    assert b"only for RadixParitioning" in _lib.lib.phj_last_error()
    cfg.algo, cfg.partitions, cfg.upload_chunks = 1, 64, 33
    assert _lib.lib.phj_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    assert b"upload_chunks" in _lib.lib.phj_last_error()
    with pytest.raises(ValueError, match="Unrecognized join algorithm type: hash-join."):
        phj.Engine("hash-join")  # src/Common/Configuration.cpp:4-12
    assert _lib.lib.phj_join(None, None) == _lib.ERR_INVALID


def test_no_cpu_fallback(phj):
    """Without a CUDA device the compute entry points fail loudly instead of computing on the CPU."""
    if phj.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(phj.PhjError, match="no CPU fallback") as e:
        phj.Engine("radix-partitioning")
    assert e.value.status == 2
    with pytest.raises(phj.PhjError, match="no CPU fallback"):
        phj.hash_batch(0, 0, np.arange(4))
