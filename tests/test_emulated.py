"""The product's engine and kernel SOURCE, executed on the CPU (tests/emu) and held to the oracle -- so that the
CPU test run also covers what otherwise only the `-m gpu` tests on the B200 see: every join plan, the partition
layout bit for bit, and the sharded join at 2 / 3 / 4 / 6 / 8 ranks (which the driver's single-GPU box skips).

tests/emu compiles the UNMODIFIED phj_engine.cu against a host stand-in for the CUDA runtime and NCCL, and the
kernel headers with g++ through a device-language shim whose launcher runs every CUDA thread as a fiber. It is test
infrastructure: the product never loads it (tests/test_host.py::test_no_cpu_fallback still holds), no number comes
from it, and it says nothing about timing or ptxas -- the `-m gpu` tests on the B200 remain the parity tests proper.
"""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
BUILD = os.path.join(EMU, "_build")
WORLDS = [2, 3, 4, 6, 8]
CHECKS = {"single": ["single"], "group": ["group"], "cli": ["cli"], "failure": ["failure"],
          **{f"ranks{w}": ["ranks", str(w)] for w in WORLDS}}


@pytest.fixture(scope="module")
def emulated_runs():
    """Builds tests/emu/_build (incremental) and starts every check at once, each in its own process."""
    if not shutil.which("nvcc") or not os.path.exists("/usr/bin/g++") or not shutil.which("make"):
        pytest.skip("nvcc, make and /usr/bin/g++ are needed to compile the emulated build")
    r = subprocess.run(["make", "-C", EMU], capture_output=True, text=True)
    assert r.returncode == 0, f"tests/emu does not build:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    env = dict(os.environ)
    env["PHJ_LIB"] = os.path.join(BUILD, "libphj_emu_engine.so")
    env["LD_LIBRARY_PATH"] = BUILD + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env.pop("PHJ_KERNEL_TIMES", None)
    # every order in which the runnable threads of a CTA take their turn is a legal execution: the checks are spread
    # over the emulator's three (tests/emu/README.md)
    order = {"group": "random", "ranks3": "reverse", "ranks8": "random", "ranks4": "reverse"}
    sms = {"ranks8": "148"}  # the B200's SM count (its segment and grid shapes); 2 elsewhere: small grids
    procs = {name: subprocess.Popen([sys.executable, os.path.join(EMU, "checks.py"), *argv],
                                    env={**env, "PHJ_EMU_SCHED": order.get(name, "forward"),
                                         "PHJ_EMU_SMS": sms.get(name, "2")},
                                    stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for name, argv in CHECKS.items()}
    yield procs
    for p in procs.values():
        if p.poll() is None:
            p.kill()


def finished(procs, name):
    out, err = procs[name].communicate(timeout=1500)
    assert procs[name].returncode == 0, f"emulated check {name} failed:\n{out[-1500:]}\n{err[-3000:]}"
    line = json.loads(out.strip().splitlines()[-1])
    assert line["ok"], line
    return line


def test_emulated_build_is_not_the_product():
    """The emulator lives under tests/ only: no product source mentions it, and the product's loader default is the
    CUDA library."""
    from partitionedhashjoin_b200 import _lib
    if "PHJ_LIB" not in os.environ:
        assert os.path.basename(_lib.LIB_PATH) == "libphj_b200.so"
    pkg = os.path.join(ROOT, "partitionedhashjoin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "cuda_emu" not in text and "libphj_emu" not in text and "PHJ_EMULATE" not in text, f


def test_single_gpu_engine_on_emulated_kernels(emulated_runs):
    """Counts of every plan on the adversarial cases, partitioned relations bit-identical to the oracle's stable
    partitioning (src/RadixCluster/HashJoin.hpp:394-412), the joined table, the streamed host join, the fallback."""
    line = finished(emulated_runs, "single")
    assert any("bit-identical" in d for d in line["done"]) and any("counts == oracle" in d for d in line["done"])


def test_one_process_sharded_join_on_emulated_gpus(emulated_runs):
    """phj_config.num_gpus = 2 / 3 / 4 / 6 / 8: the same calls as a single GPU, the global count against the oracle
    (3 and 6 GPUs: the split digit is hash % digits, not a bit field of the hash)."""
    line = finished(emulated_runs, "group")
    assert {f"{w} GPUs ok" for w in WORLDS} <= set(line["done"])


def test_cli_on_the_emulated_engine(emulated_runs):
    """host/main.cpp (the reference's CLI + the C++ joiner classes) linked against the emulated engine: all joiners,
    --stream-upload, --materialize, --gpus N for the radix and the no-partitioning join (N = 2, 3, 5, 6)."""
    finished(emulated_runs, "cli")


def test_failed_allocation_on_one_gpu_is_an_error_everywhere(emulated_runs):
    """The sizing pass of the sharded join: a window that one GPU cannot allocate makes every rank return an error
    (they agree on it before the next collective) instead of leaving the others waiting; the handle recovers."""
    finished(emulated_runs, "failure")


@pytest.mark.parametrize("world", WORLDS)
def test_sharded_join_ranks_on_emulated_gpus(emulated_runs, world):
    """phj_dist_* with one rank per thread: tests/_dist_gpu_worker.py::check_library_join, the very function the
    `-m gpu` multi-rank tests run on real GPUs (window contents tuple by tuple, both count modes, a collective
    re-size, heavy-hitter digits kept local against the plain exchange)."""
    finished(emulated_runs, f"ranks{world}")


def test_bench_refuses_the_emulated_build(emulated_runs):
    """No number may come from the emulator: bench.py stops when PHJ_LIB points at it."""
    env = dict(os.environ)
    env["PHJ_LIB"] = os.path.join(BUILD, "libphj_emu_engine.so")
    env["LD_LIBRARY_PATH"] = BUILD + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--steps", "1"], env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "emulated test build" in r.stderr + r.stdout, (r.returncode, r.stderr[-500:])
    assert not any(line.startswith("{") for line in r.stdout.splitlines())
