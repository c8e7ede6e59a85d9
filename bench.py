#!/usr/bin/env python
"""bench.py -- join throughput (|R|+|S|) tuples/s of the radix-partitioned hash join on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one complete join (partition R and S, build + probe every partition, match count back
on the host) of the workload BASELINE.json's metric is quoted on: configs[1], the 10 M x 200 M
radix join with two partitioning passes and XXH3 ("uniform" keys = the reference generator at its
minimum skew 0.01; --skew selects configs[2]/[3]). Prints ONE JSON line:

  value      tuples/s with both relations already resident in HBM (CUDA events, max over ranks)
  e2e        the same join through the C ABI's host entry point (phj_join_host): pinned host
             relations are copied to the device inside the timed region (the probe relation in chunks
             that are joined as they land), the count is read back
  roofline   radix_scatter (the dominant kernel): algorithmic bytes per launch / its mean duration,
             taken from CUDA events recorded around every launch of the timed steps
  cpu_baseline  the UNMODIFIED reference's radix join (oracle/_ref) on this box's host cores

--impl reference times that CPU implementation instead (rank 0 only).
Only the reference / cpu_baseline legs touch oracle/; the GPU arm uses the product library alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BUILD, N_PROBE = 10_000_000, 200_000_000
BASE_SEED, BATCHES = 12345, 64
METRIC = "join throughput (|R|+|S|) tuples/sec"
UNIT = "tuples/s"


def scatter_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one radix_scatter launch at this workload, read from the
    newest profiles/*_ncu_summary.md (written by tools/make_profile_summary.py from an `ncu --set full` capture of
    this command). Returns (bytes per launch or None, the file it came from)."""
    import glob
    import re
    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.md"))):
        in_scatter = False
        for line in open(path):
            if line.startswith("### "):
                in_scatter = "radix_scatter" in line
            m = re.search(r"traffic \(read\+write\) per launch: ([0-9.]+) (\w+)", line)
            if in_scatter and m:
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(m.group(2))
                if scale:
                    best = (float(m.group(1)) * scale, os.path.relpath(path, ROOT))
                in_scatter = False
    return best


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
            time.sleep(0.25)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(phj, n_build, n_probe, skew, rank=0, pinned=True):
    """Reference generator output (Sequential build side, Zipf probe side over [1, |R|])."""
    if pinned:
        Rp, Sp = phj.PinnedTuples(n_build), phj.PinnedTuples(n_probe)
        R, S = Rp.array, Sp.array
    else:
        Rp = Sp = None
        R, S = np.empty(n_build, dtype=phj.TUPLE_DTYPE), np.empty(n_probe, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1)
    phj.fill_zipf(S, skew, 1, n_build, BASE_SEED + 1000 * rank, BATCHES)
    return R, S, (Rp, Sp)


def dist_setup(n_gpus):
    """torch.distributed only for N > 1 (barrier + max over ranks + count all-reduce)."""
    if n_gpus <= 1:
        return None, 0, 1, 0
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    if world != n_gpus:
        raise SystemExit(f"--gpus {n_gpus} needs torchrun with {n_gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, rank, world, local


def reference_arm(args, rank):
    """--impl reference: the UNMODIFIED reference radix join (oracle/_ref) on the host cores, on the FULL
    workload of the GPU arm (10 M x 200 M, same skew, same seeds, the reference's own generators). Every step is
    one RadixClustering::HashJoiner::Run, timed by the reference's own phase timers. The joins run in forked
    children because the reference keeps its partitioned tables alive after Run (oracle/ref_harness.cpp)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    ref = _oracle.Reference()  # nothing of the product is imported on this arm
    hR, hS = ref.generate_tables(N_BUILD, N_PROBE, args.skew, BASE_SEED, BATCHES, threads=0)
    workers = ref.default_workers()
    times, walls, matches = [], [], None
    for i in range(args.steps + args.warmup):
        res = ref.join_tables(hR, hS, 1, partitions=args.ref_partitions, threads=0, seeded=False)
        matches = res["matches"]
        assert matches == N_PROBE, (matches, N_PROBE)
        if i >= args.warmup:
            times.append((res["partition_ns"] + res["build_ns"] + res["probe_ns"]) / 1e9)
            walls.append(res["wall_ns"] / 1e9)
    npj = ref.join_tables(hR, hS, 0, threads=0, seeded=False)
    assert npj["matches"] == N_PROBE
    per_step = sum(times) / len(times)
    value = (N_BUILD + N_PROBE) / per_step
    sample = (f"full workload {N_BUILD} x {N_PROBE}, reference RadixClustering::HashJoiner P={args.ref_partitions}, "
              f"{workers} workers, {len(times)} timed runs: phase sum partition+build+probe (the reference's own timers; "
              f"Run() incl. its untimed 3.36 GB of allocation takes {sum(walls) / len(walls):.2f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference", "sample": sample,
                         "no_partitioning": {"value": (N_BUILD + N_PROBE) / (npj["probe_ns"] / 1e9), "unit": UNIT,
                                             "ms": npj["probe_ns"] / 1e6,
                                             "note": "NoPartitioning::HashJoiner, 1 run; its 'probe' figure is "
                                                     "build + probe (Results.hpp:202), used as the total"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "matches": matches,
    }
    print(json.dumps(line), flush=True)
    ref.free_table(hR)
    ref.free_table(hS)


def workload_config(args, n_probe=N_PROBE, n_build=N_BUILD):
    return {"workload": f"radix-cluster hash join {n_build // 10**6}M x {n_probe // 10**6}M, 2-pass radix "
                        f"partitioning ({args.partitions} partitions), {args.hash}, Zipf skew {args.skew}"
                        f"{' (uniform)' if args.skew <= 0.01 else ''}, count-only",
            "primary": n_build, "secondary": n_probe, "skew": args.skew, "partitions": args.partitions,
            "hash": args.hash, "tuple_bytes": 16, "l2_flush": "inputs (3.36 GB per GPU) are larger than L2"}


def cpu_baseline_leg(args, R, S):
    """The reference's own CPU joins -- radix (P = --ref-partitions) AND no-partitioning -- timed on this box's
    host cores on the very arrays the GPU joined (rank 0, N = 1), each in a forked child (see reference_arm)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle
        ref = _oracle.Reference()
    except Exception as e:  # the checker is optional for the GPU arm
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    workers = ref.default_workers()
    hR, hS = ref.new_table(R.shape[0]), ref.new_table(S.shape[0])
    ref.table_view(hR, R.shape[0])[:] = R
    ref.table_view(hS, S.shape[0])[:] = S
    best = None
    for _ in range(2):
        res = ref.join_tables(hR, hS, 1, partitions=args.ref_partitions, threads=0, seeded=False)
        assert res["matches"] == S.shape[0], res
        if best is None or sum(res[k] for k in ("partition_ns", "build_ns", "probe_ns")) < \
                sum(best[k] for k in ("partition_ns", "build_ns", "probe_ns")):
            best = res
    npj = ref.join_tables(hR, hS, 0, threads=0, seeded=False)
    assert npj["matches"] == S.shape[0], npj
    ref.free_table(hR)
    ref.free_table(hS)
    n = R.shape[0] + S.shape[0]
    phase = (best["partition_ns"] + best["build_ns"] + best["probe_ns"]) / 1e9
    return {"value": n / phase, "unit": UNIT, "cores": workers, "kind": "reference",
            "sample": f"full workload, best of 2 runs of the reference RadixClustering::HashJoiner P={args.ref_partitions} "
                      f"({workers} workers): phases {best['partition_ns'] / 1e6:.0f}/{best['build_ns'] / 1e6:.0f}/"
                      f"{best['probe_ns'] / 1e6:.0f} ms (its own timers), Run() wall {best['wall_ns'] / 1e9:.2f} s",
            "matches": best["matches"],
            "no_partitioning": {"value": n / (npj["probe_ns"] / 1e9), "unit": UNIT, "cores": workers,
                                "build_ms": npj["build_ns"] / 1e6, "total_ms": npj["probe_ns"] / 1e6,
                                "matches": npj["matches"],
                                "sample": "full workload, 1 run of the reference NoPartitioning::HashJoiner; its "
                                          "'probe' figure already contains the build (Results.hpp:202) and is "
                                          "used as the total"}}


def gpu_arm(args):
    dist, rank, world, local = dist_setup(args.gpus)
    os.environ["PHJ_KERNEL_TIMES"] = "1"  # CUDA events around every kernel launch (roofline.achieved)
    import partitionedhashjoin_b200 as phj
    if phj.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    if "emulated" in phj.device_info(0)["name"]:  # PHJ_LIB pointing at tests/emu's build: logic tests only
        raise SystemExit("bench.py measures the B200: the emulated test build (tests/emu) is not a device")

    if world > 1 or args.workload == "scaled":
        from partitionedhashjoin_b200 import multigpu
        return multigpu.bench(args, dist, rank, world, local, make_inputs, workload_config, ClockSampler,
                              METRIC, UNIT, measured_hbm_peak)

    R, S, keep = make_inputs(phj, N_BUILD, N_PROBE, args.skew)
    n_tuples = R.shape[0] + S.shape[0]
    eng = phj.Engine("radix-partitioning", partitions=args.partitions, hash=args.hash, device=local)
    eng.upload(R, S)
    # warm-up: CUDA events around EVERY kernel (the per-kernel table of the line); the timed steps keep them only
    # around radix_scatter, the roofline kernel -- a pair of events costs the stream about a microsecond
    eng.kernel_timing("")
    warm_ns = {}
    for _ in range(args.warmup):
        res = eng.join()
        for name, ns in eng.kernel_times():
            warm_ns.setdefault(name, []).append(ns)
    assert res["matches"] == S.shape[0], res
    eng.kernel_timing("radix_scatter")

    # ---- timed region: K joins, inputs resident in HBM ----
    import ctypes as C
    launches, kernel_ns = 0, {}
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        device_ns = 0
        for _ in range(args.steps):
            res = eng.join()  # synchronous: returns with the count on the host
            device_ns += res["total_ns"]
            launches += res["kernel_launches"]
            for name, ns in eng.kernel_times():
                kernel_ns.setdefault(name, []).append(ns)
        wall = time.perf_counter() - t0
    assert res["matches"] == S.shape[0], res
    ms_per_step = device_ns / args.steps / 1e6
    value = n_tuples / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_hbm_peak()
    scat = [ns for name, v in kernel_ns.items() if name.startswith("radix_scatter") for ns in v]
    scat_ms = sum(scat) / len(scat) / 1e6
    alg_bytes = 32.0 * n_tuples  # every tuple is read once and written once: 16 B + 16 B
    achieved = alg_bytes / (scat_ms / 1e3) / 1e9
    traffic, traffic_src = scatter_dram_traffic()
    # medians: the first warm-up join also pays the kernels' one-time module load
    per_kernel = {name: round(sorted(v)[len(v) // 2] / 1e3, 1) for name, v in {**warm_ns, **kernel_ns}.items()}
    kernel_share = sum(scat) / max(1, device_ns)  # both scatter launches of a step over the step's device time

    # ---- e2e: host relations in, count out, through phj_join_host ----
    # Streamed by default: the probe relation goes up in ~256 MB chunks, each joined as soon as it has
    # landed, so only the last chunk's join is not hidden behind PCIe. The unstreamed call (upload
    # everything, then join) is timed beside it for comparison.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 8))
        eng.join_host(R, S)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r2 = eng.join_host(R, S)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert r2["matches"] == S.shape[0]
        with phj.Engine("radix-partitioning", partitions=args.partitions, hash=args.hash, device=local,
                        upload_chunks=1) as e1:
            e1.join_host(R, S)
            t0 = time.perf_counter()
            for _ in range(3):
                r1 = e1.join_host(R, S)
            e2e_plain_s = (time.perf_counter() - t0) / 3
            assert r1["matches"] == S.shape[0] and r1["upload_chunks"] == 1
        # the same call from PAGEABLE memory -- what a Common::Table (a std::vector, host/Common/Table.hpp) holds
        # and what the `phjoin` CLI therefore delivers; the uploads are issued by the helper thread
        Rq, Sq = np.array(R, copy=True), np.array(S, copy=True)
        eng.join_host(Rq, Sq)
        t0 = time.perf_counter()
        for _ in range(3):
            rq = eng.join_host(Rq, Sq)
        e2e_pageable_s = (time.perf_counter() - t0) / 3
        assert rq["matches"] == S.shape[0]
        del Rq, Sq
        e2e = {"value": n_tuples / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(r2["h2d_bytes"]),
               "host_memory": "pinned (phj_host_alloc)",
               "pageable": {"value": n_tuples / e2e_pageable_s, "unit": UNIT, "ms_per_step": e2e_pageable_s * 1e3,
                            "h2d_ms": rq["h2d_ns"] / 1e6, "steps": 3,
                            "note": "same call from pageable numpy arrays (std::vector-like), as the phjoin CLI"},
               "d2h_bytes_per_step": int(r2["d2h_bytes"]), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "upload_chunks": int(r2["upload_chunks"]), "h2d_ms": r2["h2d_ns"] / 1e6,
               "device_ms": r2["e2e_ns"] / 1e6, "unstreamed_ms_per_step": e2e_plain_s * 1e3,
               "unstreamed_h2d_ms": r1["h2d_ns"] / 1e6}

    # ---- the other configurations of BASELINE.json, a few steps each (informational) ----
    others = {}
    if not args.quick:
        def few(engine, n=3):
            engine.upload(R, S)
            runs = [engine.join() for _ in range(n + 1)][1:]
            for r in runs:  # generator data: every probe key has a build match
                assert r["matches"] == S.shape[0], r
            return round(n_tuples / (min(r["total_ns"] for r in runs) / 1e9) / 1e9, 2)
        for h in ("xxh3", "murmur3", "city"):
            if h != args.hash:
                with phj.Engine("radix-partitioning", partitions=args.partitions, hash=h, device=local) as e2:
                    others[f"radix_{h}_Gtuples_s"] = few(e2)
        with phj.Engine("no-partitioning", device=local) as e2:
            others["no_partitioning_Gtuples_s"] = few(e2)
        with phj.Engine("no-partitioning", device=local, flags=phj.FLAG_CHAINED_TABLE) as e2:
            others["no_partitioning_chained_table_Gtuples_s"] = few(e2)
        with phj.Engine("radix-partitioning", partitions=64, hash=args.hash, device=local) as e2:
            others["radix_1pass_64_partitions_Gtuples_s"] = few(e2)
        # one partitioning pass + per-partition tables probed out of L2 (PHJ_FLAG_L2_TABLES): the plan the sharded
        # join uses locally, and the fastest single-GPU join here (64 B/tuple instead of 96)
        for parts in (64, 8):
            try:
                with phj.Engine("radix-partitioning", partitions=parts, hash=args.hash, device=local,
                                flags=phj.FLAG_L2_TABLES) as e2:
                    others[f"radix_1pass_{parts}_partitions_l2_tables_Gtuples_s"] = few(e2)
            except Exception as exc:  # informational: never takes the headline line down
                others[f"radix_1pass_{parts}_partitions_l2_tables_Gtuples_s"] = f"failed: {exc}"[:200]
        with phj.Engine("radix-partitioning", partitions=args.partitions, hash=args.hash, device=local) as e2:
            e2.upload(R, S)
            e2.join_materialize()
            rm = min((e2.join_materialize() for _ in range(3)), key=lambda r: r["total_ns"])
            assert rm["joined_tuples"] == S.shape[0]
            others["radix_materialized_Gtuples_s"] = round(n_tuples / (rm["total_ns"] / 1e9) / 1e9, 2)
            others["radix_materialized_rows"] = rm["joined_tuples"]
        for skew in (1.05, 1.25):
            if abs(skew - args.skew) > 1e-9:
                phj.fill_zipf(S, skew, 1, N_BUILD, BASE_SEED, BATCHES)
                with phj.Engine("radix-partitioning", partitions=args.partitions, hash=args.hash, device=local) as e2:
                    others[f"radix_zipf{skew}_Gtuples_s"] = few(e2)
                with phj.Engine("no-partitioning", device=local) as e2:
                    others[f"no_partitioning_zipf{skew}_Gtuples_s"] = few(e2)
        phj.fill_zipf(S, args.skew, 1, N_BUILD, BASE_SEED, BATCHES)

    cpu = cpu_baseline_leg(args, R, S) if not args.no_cpu_baseline else None
    eng.close()
    if not args.quick:
        # BASELINE.json configs[4] on ONE GPU (the N = 1 point of its strong scaling): 160 M x 3.2 B through the
        # library's sharded join with a single rank -- 54 GB of input + 60 GB of windows fit one B200's HBM
        import torch

        from partitionedhashjoin_b200 import multigpu
        del keep, R, S
        torch.cuda.set_device(local)
        free, _ = torch.cuda.mem_get_info(local)
        if free > 140e9:
            others["scaled_160Mx3200M"] = multigpu.scaled_other_config(phj, multigpu.SingleRank, torch, 0, 1, local, args)
        else:
            others["scaled_160Mx3200M"] = {"skipped": f"only {free / 1e9:.0f} GB of HBM free"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic", "config": workload_config(args),
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "radix_scatter", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": scat_ms,
                     "launches_per_step": len(scat) // args.steps, "share_of_step": kernel_share,
                     "join_bytes_alg_96B_per_tuple_frac": 96.0 * n_tuples / (ms_per_step / 1e3) / 1e9 / peak,
                     "join_bytes_alg_96B_per_tuple_frac_of_nominal_8000GBs":
                         96.0 * n_tuples / (ms_per_step / 1e3) / 1e9 / 8000.0},
        # the reference's three phases (IHashJoinTimer) of the last timed join, device times: partitioning =
        # histograms + scans + both scatters; build / probe = shares of the CTA with the largest build + probe
        # (src/RadixCluster/HashJoin.hpp:67-87); join = the fused build + probe kernel as a whole
        "phases_ms": {"partition": res["partition_ns"] / 1e6, "build": res["build_ns"] / 1e6,
                      "probe": res["probe_ns"] / 1e6, "join": res["join_ns"] / 1e6},
        "cpu_baseline": cpu, "clocks": clocks.summary(), "kernel_us": per_kernel,
        "wall_ms_per_step": wall / args.steps * 1e3, "matches": res["matches"], "other_configs": others,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skew", type=float, default=0.01)
    ap.add_argument("--hash", default="xxh3", choices=["xxh3", "murmur3", "city"])
    ap.add_argument("--partitions", type=int, default=4096)
    ap.add_argument("--ref-partitions", type=int, default=2048)
    ap.add_argument("--workload", default="default", choices=["default", "scaled"],
                    help="N > 1 only: scaled = BASELINE.json configs[4], 160M x 3.2B in total sharded over the ranks "
                         "(strong scaling), generated on the device; use --partitions 8192")
    ap.add_argument("--chunks", type=int, default=0, help="N > 1: probe chunks of the shuffle (0 = the library's default, 4)")
    ap.add_argument("--dist-partitions", type=int, default=0,
                    help="N > 1: split digits = GPUs x local partitions (a power of two <= 256; 0 = the library's default, 64)")
    ap.add_argument("--shuffle", default="library", choices=["library", "pass1", "fused", "nccl", "npj"],
                    help="N > 1: library = the sharded join inside libphj_b200.so (phj_dist_*: device-side layout, NVLink "
                         "peer stores, probe chunks overlapping the shuffle, L2-table local join); the host-driven paths "
                         "of round 1: pass1 = the split scatter writes (owner : pass-1 digit) pieces into the owners' "
                         "windows and the local join starts at pass 2 (keeps heavy-hitter digits local); fused = same "
                         "stores, split by owner only; nccl = local split + NCCL all-to-all; npj = the no-partitioning "
                         "join: build shards gathered on every rank, probe shards stay (no shuffle)")
    ap.add_argument("--quick", action="store_true", help="skip the informational extra configurations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the end-to-end leg (profiling runs: keeps the ncu launch list to the timed joins)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", 0))
        reference_arm(args, rank)
        return
    gpu_arm(args)


if __name__ == "__main__":
    main()
