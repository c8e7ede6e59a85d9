"""The C++ host mirror (partitionedhashjoin_b200/host): CLI surface and JSON rendering match the
reference's (src/main.cpp:141-208, src/Arguments.hpp:7-19, src/Common/Results.hpp:262-279)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "partitionedhashjoin_b200", "host")
PHJOIN = os.path.join(HOST, "phjoin")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))


@pytest.fixture(scope="module")
def phjoin():
    subprocess.run(["make", "-C", os.path.join(ROOT, "partitionedhashjoin_b200", "csrc")], check=True, capture_output=True)
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return PHJOIN


def run(phjoin, *args, cwd=None):
    return subprocess.run([phjoin, *args], capture_output=True, text=True, cwd=cwd)


def test_help_lists_reference_flags(phjoin):
    r = run(phjoin, "--help")
    assert r.returncode == 0
    for flag in ("--primary", "--secondary", "--skew", "--log", "--join", "--format", "--unit", "--output",
                 "--filename", "--partitions", "=10000000", "=200000000", "=1.05", "=hashjoin.txt"):
        assert flag in r.stdout
    assert run(phjoin, "-h").returncode == 0


@pytest.mark.parametrize("args,message", [
    ((), "the option '--join' is required but missing"),
    (("--join", "hash-join"), "Unrecognized join algorithm type: hash-join."),
    (("--join", "no-partitioning", "-p", "32"), "number of partitions can be specified only for RadixParitioning."),
    (("--join", "radix-partitioning", "--unit", "minutes"), "Unrecognized time unit: minutes"),
    (("--join", "radix-partitioning", "--format", "xml"), "Unrecognized results format: xml."),
    (("--join", "radix-partitioning", "--output", "stdout"), "Unrecognized output type: stdout."),
    (("--join", "radix-partitioning", "--filename", ""), "empty configuration filename specified."),
    (("--join", "radix-partitioning", "--log", "verbose"), "Unrecognized logger level: verbose."),
    (("--join", "radix-partitioning", "--primary", "ten"), "option '--primary' is invalid"),
    (("--join", "radix-partitioning", "--bogus", "1"), "unrecognised option '--bogus'"),
    (("--join", "radix-partitioning", "--hash", "sha1"), "Unrecognized hash function: sha1."),
    (("--join", "no-partitioning", "--table", "cuckoo"), "Unrecognized hash table type: cuckoo."),
    (("--join", "no-partitioning", "--materialize"), "the joined table is produced by the RadixParitioning joiner."),
    (("--join", "radix-partitioning", "--stream-upload", "--repeat", "2"), "--stream-upload joins once and count-only"),
    (("--join", "radix-partitioning", "--stream-upload", "--materialize"), "--stream-upload joins once and count-only"),
])
def test_argument_errors_exit_1_with_option_table(phjoin, args, message):
    """Any parse/validation error: message, option table, exit(1) (reference src/main.cpp:199-205)."""
    r = run(phjoin, *args)
    assert r.returncode == 1
    assert message in r.stdout and "Allowed options" in r.stdout


def test_json_rendering_matches_reference(tmp_path):
    """Common::JSONResultsFormatter output == what the reference's formatter (boost ptree) writes."""
    g = GOLDEN["json"]
    src = tmp_path / "fmt.cpp"
    params = "".join(f'  p.SetParameter("{k}", "{v}");\n' for k, v in g["parameters"].items())
    src.write_text(
        f'#include "{HOST}/Common/Results.hpp"\n#include <iostream>\n'
        "int main(int, char** argv) {\n  Common::Parameters p;\n" + params +
        f'  Common::HashJoinTimingResult r(std::chrono::nanoseconds({g["build_ns"]}LL), '
        f'std::chrono::nanoseconds({g["probe_ns"]}LL), std::chrono::nanoseconds({g["partition_ns"]}LL), p);\n'
        "  Common::ResultsFormatConfiguration c; c.TimeUnit = argv[1];\n"
        "  Common::JSONResultsFormatter f(c); f.Format(std::cout, r); return 0; }\n")
    exe = tmp_path / "fmt"
    subprocess.run(["g++", "-std=c++17", "-o", str(exe), str(src)], check=True)
    for unit in ("ms", "us"):
        out = subprocess.run([str(exe), unit], capture_output=True, text=True, check=True).stdout
        assert out == g[unit]


@pytest.mark.gpu
@pytest.mark.parametrize("join,extra,typ", [("no-partitioning", [], "NoPartitioning"),
                                             ("no-partitioning", ["--table", "separate-chaining"], "NoPartitioning"),
                                             ("radix-partitioning", ["-p", "64"], "RadixParitioning"),
                                             ("radix-partitioning", ["--partitions=4096", "--hash", "city", "--repeat", "2"], "RadixParitioning")])
def test_cli_end_to_end(phjoin, tmp_path, join, extra, typ):
    out = tmp_path / "result.txt"
    r = run(phjoin, "--join", join, "--primary", "200000", "--secondary", "3000000", "--skew", "1.25", "-u", "us",
            "-f", str(out), *extra)
    assert r.returncode == 0, r.stderr
    assert "Joined 3000000 tuples." in r.stderr  # generator data: every probe key lies in [1, |R|]
    d = json.load(open(out))
    assert d["id"] == "hashjointimingresult"
    assert d["parameters"]["Type"] == typ and d["parameters"]["PrimaryRelationSize"] == "200000"
    assert d["parameters"]["Skew"] == "1.250000"
    assert (d["parameters"].get("NumberOfPartitions") is not None) == (join == "radix-partitioning")
    assert list(d["results"]) == ["partition", "build", "probe"]
    assert int(d["results"]["probe"]) > 0 and int(d["results"]["build"]) > 0
    assert (int(d["results"]["partition"]) > 0) == (join == "radix-partitioning")


@pytest.mark.gpu
@pytest.mark.parametrize("join,extra", [("no-partitioning", []), ("radix-partitioning", ["-p", "1024"])])
def test_cli_stream_upload(phjoin, tmp_path, join, extra):
    """--stream-upload: Run() through phj_join_host (24 MB probe chunks do not stream by themselves at
    this size, so the count and the end-to-end log line are what is checked; the chunked path itself
    is covered by test_gpu_parity.py::test_join_host_streamed)."""
    out = tmp_path / "result.txt"
    r = run(phjoin, "--join", join, "--primary", "200000", "--secondary", "3000000", "--stream-upload", "--log", "info",
            "-f", str(out), *extra)
    assert r.returncode == 0, r.stderr
    assert "End to end (upload in 1 chunk(s) + join)" in r.stderr
    assert json.load(open(out))["parameters"]["SecondaryRelationSize"] == "3000000"


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_cli_several_gpus(phjoin, tmp_path, gpus):
    """--gpus N: ONE process drives N GPUs through phj_config.num_gpus (the reference's caller is one Run(tableA,
    tableB, timer), src/main.cpp:110-139); -p = GPUs x partitions per GPU. Skipped when the box has fewer GPUs."""
    import partitionedhashjoin_b200 as phj
    if phj.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    out = tmp_path / "result.txt"
    r = run(phjoin, "--join", "radix-partitioning", "--primary", "2000000", "--secondary", "30000000", "--skew", "1.05",
            "--gpus", str(gpus), "-p", "64", "--repeat", "3", "--log", "debug", "-u", "us", "-f", str(out))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Joined 30000000 tuples." in r.stderr
    d = json.load(open(out))
    assert d["parameters"]["Type"] == "RadixParitioning" and d["parameters"]["NumberOfPartitions"] == "64"
    assert int(d["results"]["partition"]) > 0 and int(d["results"]["probe"]) > 0
    # partitions must be GPUs x partitions per GPU
    r = run(phjoin, "--join", "radix-partitioning", "--primary", "1000", "--secondary", "1000", "--gpus", str(gpus), "-p", "1")
    assert r.returncode == 1 and "GPUs x local partitions" in r.stderr
    # the no-partitioning joiner on the same GPUs: every GPU builds the whole table, probes its share of tableB
    r = run(phjoin, "--join", "no-partitioning", "--primary", "1500000", "--secondary", "30000000", "--skew", "1.05",
            "--gpus", str(gpus), "--log", "debug", "-f", str(out))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Joined 30000000 tuples." in r.stderr


def test_cli_several_gpus_argument_errors(phjoin):
    for args, message in ((("--join", "no-partitioning", "--gpus", "2", "-p", "64"), "only for RadixParitioning"),
                          (("--join", "radix-partitioning", "--gpus", "0"), "--gpus must be at least 1"),
                          (("--join", "radix-partitioning", "--gpus", "2", "--materialize"), "--gpus > 1 counts only")):
        r = run(phjoin, *args)
        assert r.returncode == 1 and message in r.stdout


@pytest.mark.gpu
def test_cli_materialize(phjoin, tmp_path):
    """--materialize: Run() returns the filled Table<JoinedTuple>; generator data joins 1:1."""
    out = tmp_path / "result.txt"
    r = run(phjoin, "--join", "radix-partitioning", "--primary", "100000", "--secondary", "1500000", "-p", "256",
            "--materialize", "--log", "info", "-f", str(out))
    assert r.returncode == 0, r.stderr
    assert "Joined table holds 1500000 rows." in r.stderr


@pytest.mark.gpu
def test_partition_sweep_writes_figure_dat(phjoin, tmp_path):
    """tools/sweep.py = scripts/generate.sh:66-80: figure.dat with one column per run."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "-s", "1.05", "--primary", "100000",
                        "--secondary", "1000000", "--outdir", str(tmp_path), "--repeat", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [l.split() for l in open(tmp_path / "figure.dat").read().splitlines()]
    assert rows[0] == ["NumberOfPartitions", "NoPartitioning"] + [f"Radix{p}" for p in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192)]
    assert [r_[0] for r_ in rows[1:]] == ["Partition", "Build", "Probe"]
    assert rows[1][1] == "0" and all(int(v) > 0 for v in rows[1][2:])  # only radix runs partition
