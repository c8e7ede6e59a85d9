"""Instruction mix (by opcode) of one kernel in an .ncu-rep, weighted by execution count."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
mix = collections.Counter(); tot = 0
for r in data:
    src = r[idx["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = float(r[idx["Instructions Executed"]] or 0)
    mix[op] += n; tot += n
print(f"{kern}: {tot:.0f} warp instructions")
for op, n in mix.most_common(28):
    print(f"  {op:10s} {n:14.0f} {100*n/tot:5.1f}%")
