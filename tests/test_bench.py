"""CPU tests of bench.py's host side: the helpers behind the roofline / clocks keys of its JSON line, the
reference arm's rank handling and the loud failure of the GPU arm on a box without a GPU. The measurements
themselves need the B200 (the driver runs bench.py there)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("phj_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_scatter_traffic_comes_from_the_newest_profile_summary(bench):
    """roofline.traffic is read from profiles/*_ncu_summary.md (never a pasted constant): per launch it must be
    within a few per cent of the algorithmic 32 B x 210 M tuples -- no wasted re-reads."""
    traffic, source = bench.scatter_dram_traffic()
    assert source and os.path.exists(os.path.join(ROOT, source)) and source.endswith("_ncu_summary.md")
    algorithmic = 32.0 * (bench.N_BUILD + bench.N_PROBE)
    assert 0.95 * algorithmic < traffic < 1.05 * algorithmic, (traffic, algorithmic, source)


def test_hbm_peak_is_the_measured_one(bench):
    peak, source = bench.measured_hbm_peak()
    want = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    assert peak == want and source.startswith("measured")


def test_clock_sampler_summary_parses_nvidia_smi_lines(bench):
    s = bench.ClockSampler(0)
    s.lines = ["1965, 1965, Not Active, Not Active, Not Active, Not Active\n",
               "1950, 1965, Not Active, Not Active, Not Active, Active\n",
               "1920, 1965, Not Active, Not Active, Not Active, Active\n",
               "garbage\n", "N/A, N/A, a, b, c, d\n"]
    out = s.summary()
    assert out == {"sm_mhz": 1950.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    s.lines = []
    assert s.summary()["samples"] == 0 and s.summary()["sm_mhz"] is None


def test_workload_config_names_the_baseline_configuration(bench):
    import argparse
    args = argparse.Namespace(partitions=4096, hash="xxh3", skew=0.01)
    cfg = bench.workload_config(args)
    assert cfg["primary"] == 10_000_000 and cfg["secondary"] == 200_000_000 and cfg["tuple_bytes"] == 16
    assert "2-pass" in cfg["workload"] and "(uniform)" in cfg["workload"] and "model" not in cfg
    assert "(uniform)" not in bench.workload_config(argparse.Namespace(partitions=4096, hash="city", skew=1.25))["workload"]


def run_bench(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          env=e, timeout=300)


def test_reference_arm_other_ranks_exit_without_work():
    """Under torchrun only rank 0 runs the CPU reference; the other ranks print nothing and exit 0."""
    r = run_bench("--impl", "reference", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-300:], r.stderr[-300:])


def test_gpu_arm_fails_loudly_without_a_gpu(phj):
    """No CPU fallback anywhere on the product path: without a CUDA device bench.py stops with a message."""
    if phj.device_count() > 0:
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--quick")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout), (r.returncode, r.stderr[-300:])
    assert not any(line.startswith("{") for line in r.stdout.splitlines())


def test_shuffle_traffic_comes_from_the_split_scatter_capture():
    """roofline.traffic of the 8-GPU line: DRAM read + write of the NVLink scatter from profiles/r02r_ncu_split_summary.md
    (the shard is read once, 1 / 8 of it is written locally); no capture for other GPU counts, so none is claimed."""
    from partitionedhashjoin_b200 import multigpu
    traffic, source = multigpu.shuffle_scatter_dram_traffic(8)
    shard = 16.0 * 210_000_000
    assert source and os.path.exists(os.path.join(ROOT, source)) and 1.0 * shard < traffic < 1.25 * shard
    assert multigpu.shuffle_scatter_dram_traffic(4) == (None, None)
