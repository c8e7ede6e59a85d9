"""Summarise an `ncu --set full` capture (.ncu-rep) kernel by kernel into markdown (stdout).

  python tools/summarize_rep.py gpurun_out/x.ncu-rep "title / command" >> profiles/rNN_x.md
"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers/thread"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors, global loads"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 requests, global loads"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"), ("lts__t_sectors.sum", "L2 sectors"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]
print(f"## `{rep.split('/')[-1]}` — {title}\n")
seen = set()
for r in data:
    name = r[idx["Kernel Name"]].replace("void phj::", "").split("(phj::")[0]
    if name in seen:
        continue
    seen.add(name)
    print(f"### `{name}`\n")
    for key, label in want:
        if key in idx and r[idx[key]]:
            print(f"* {label}: {r[idx[key]]} {units[idx[key]]}")
    st = [(float(r[idx[h]]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and r[idx[h]]]
    tot = sum(v for v, _ in st) or 1
    print("* top stall reasons (pc samples): " + ", ".join(f"{n} {100 * v / tot:.0f} %" for v, n in sorted(st, reverse=True)[:5]))
    print()
