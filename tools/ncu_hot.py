"""Print the hottest SASS instructions (by warp-stall samples) of one kernel in an .ncu-rep."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
key = "Warp Stall Sampling (All Samples)"
tot = sum(float(r[idx[key]] or 0) for r in data)
inst = sum(float(r[idx["Instructions Executed"]] or 0) for r in data)
print(f"kernel {kern}: {len(data)} SASS lines, {tot:.0f} samples, {inst:.0f} warp instructions")
ranked = sorted(enumerate(data), key=lambda t: -float(t[1][idx[key]] or 0))[:top]
for i, r in ranked:
    print(f"{100 * float(r[idx[key]]) / tot:6.2f}%  #{i:4d} exec={float(r[idx['Instructions Executed']]):12.0f}  {r[idx['Source']].strip()[:100]}")
