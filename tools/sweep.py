#!/usr/bin/env python
"""Partition sweep of the reference's scripts/generate.sh (lines 66-80) on the B200 engine: one
no-partitioning run, then radix-partitioning with P = 32 .. 8192, each through the phjoin CLI, and
the timing JSONs folded into figure.dat with the reference's layout (one column per run, rows
NumberOfPartitions-header / Partition / Build / Probe) so scripts/figure.plot can draw it.

    python tools/sweep.py --skew 1.05 [--primary N --secondary M] [--outdir DIR] [--unit us]

generate.sh:78 passes --skew 1.05 to every radix run whatever -s says; here
the radix runs use the requested skew unless --reference-skew-bug is given.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHJOIN = os.path.join(ROOT, "partitionedhashjoin_b200", "host", "phjoin")
PARTITIONS = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-p", "--path", default=PHJOIN)
    ap.add_argument("-s", "--skew", type=float, required=True)
    ap.add_argument("--primary", type=int)
    ap.add_argument("--secondary", type=int)
    ap.add_argument("--unit", default="us")
    ap.add_argument("--outdir")
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--reference-skew-bug", action="store_true")
    args = ap.parse_args()
    outdir = args.outdir or str(args.skew)
    os.makedirs(outdir, exist_ok=True)
    extra = ["--unit", args.unit, "--repeat", str(args.repeat)]
    if args.primary:
        extra += ["--primary", str(args.primary)]
    if args.secondary:
        extra += ["--secondary", str(args.secondary)]
    columns = [["NumberOfPartitions", "Partition", "Build", "Probe"]]

    def run(title, join, skew, more):
        name = os.path.join(outdir, f"partitions_{title}.txt")
        cmd = [args.path, "--skew", str(skew), "--join", join, "-o", "file", "--filename", name, *more, *extra]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit(f"{' '.join(cmd)} failed:\n{r.stdout}\n{r.stderr}")
        res = json.load(open(name))["results"]
        columns.append([title if join == "no-partitioning" else f"Radix{title}", *[str(v) for v in res.values()]])

    run("NoPartitioning", "no-partitioning", args.skew, [])
    columns[-1][0] = "NoPartitioning"
    for p in PARTITIONS:
        run(str(p), "radix-partitioning", 1.05 if args.reference_skew_bug else args.skew, ["-p", str(p)])
    with open(os.path.join(outdir, "figure.dat"), "w") as f:
        for row in zip(*columns):
            f.write(" ".join(row) + "\n")
    print(open(os.path.join(outdir, "figure.dat")).read())


if __name__ == "__main__":
    main()
