"""Turn ncu captures under gpurun_out/ into the tracked summaries under profiles/.

  python tools/make_profile_summary.py r01 gpurun_out/r01_launches.csv gpurun_out/prof_r01_top.ncu-rep ["bench.py flags"]
"""
import csv, io, json, subprocess, sys, collections

tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
flags = sys.argv[4] if len(sys.argv) > 4 else "--steps 2 --warmup 3 --quick --no-cpu-baseline"
out = [f"# ncu summary {tag}", "",
       f"Command: `python bench.py {flags}` on one B200 (gpurun); both ncu passes",
       "ran only after the same command had exited 0 without ncu. Times under ncu are serialised and",
       "cold-cache: compare SHARES, not absolutes (the live CUDA-event numbers are in the bench line).", ""]

# ---- launch list ----
rows = [r for r in csv.reader(l for l in open(launches_csv) if l.startswith('"'))]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void phj::", "")
    name = name.split("<")[0]
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit.startswith("ms") else 1.0)
    per.setdefault(name, []).append(v_us)
total = sum(sum(v) for v in per.values())
out += [f"## Launch list ({sum(len(v) for v in per.values())} launches, `--metrics gpu__time_duration.sum --clock-control none`)", "",
        "| kernel | launches | mean us | total us | share |", "|---|---|---|---|---|"]
for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| {name} | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.0f} | {100*sum(v)/total:.1f} % |")
out.append("")

# ---- full-set capture ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers/thread"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active % of peak"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active % of peak"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % of peak"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "warps stalled on a barrier per issue")]
out += [f"## `--set full` capture of the top kernels (`{rep.split('/')[-1]}`, not tracked: binary)", ""]
seen = set()
for r in data:
    name = r[idx["Kernel Name"]].replace("void phj::", "").split("(phj::")[0]
    if name in seen:
        continue
    seen.add(name)
    out.append(f"### `{name}`")
    out.append("")
    for key, label in want:
        if key in idx and r[idx[key]]:
            out.append(f"* {label}: {r[idx[key]]} {units[idx[key]]}")
    st = [(float(r[idx[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and r[idx[h]]]
    tot = sum(v for v, _ in st) or 1
    out.append("* top stall reasons (pc samples): " + ", ".join(f"{n} {100*v/tot:.0f} %" for v, n in sorted(st, reverse=True)[:5]))
    try:
        rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
        u = units[idx["dram__bytes_read.sum"]]
        out.append(f"* **traffic (read+write) per launch: {rd + wr:.3f} {u}**")
    except Exception:
        pass
    out.append("")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
