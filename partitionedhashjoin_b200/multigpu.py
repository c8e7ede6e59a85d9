"""Multi-GPU radix join: one process per GPU, torch.distributed for the plumbing.

The reference is single-process (no communication backend at all); radix
partitions are independent join units, so the join shards by partition with ONE exchange step:

  1. every rank holds a row-range shard of R and of S (phj_upload);
  2. split   -- PHJ_ALGO_SHARD_SPLIT: histogram / scan / scatter of the shard by owner rank,
                owner = (hash >> SHARD_SHIFT) % world (hash bits disjoint from the local join's);
  3. exchange-- the per-owner pieces travel with one all-to-all per relation (NCCL over NVLink;
                sizes first, a 2*world int64 all-to-all);
  4. local   -- an ordinary PHJ_ALGO_RADIX_PARTITIONING join on what arrived (phj_bind_device on
                the receive buffers, zero-copy);
  5. reduce  -- all-reduce (sum) of the match counts.

Equal keys hash to the same owner, so the sum of the local counts is the global count. The host
logic below is backend-agnostic: `GpuBackend` drives libphj_b200.so; the tests run the same
orchestration over gloo with a CPU stand-in (tests/test_distributed.py).
"""
from __future__ import annotations

import time

import numpy as np

SHARD_SHIFT = 32  # owner digit = hash bits [32, 32 + log2 world): disjoint from the <= 16 partition bits


class _CudaView:
    """Zero-copy torch view of a raw device pointer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, rows: int):
        self.__cuda_array_interface__ = {"shape": (rows, 2), "typestr": "<i8", "data": (ptr, False), "version": 3,
                                         "strides": None}


class GpuBackend:
    """split / local join on the device through the C ABI (no CPU fallback)."""

    def __init__(self, world, device, partitions_local=0, hash="xxh3", hash_seed=0x9E3779B97F4A7C15):
        import torch

        from . import engine
        self.torch, self.world, self.device = torch, world, device
        self.split_engine = engine.Engine("shard-split", partitions=world, hash=hash, hash_seed=hash_seed,
                                          device=device, shard_shift=SHARD_SHIFT)
        self.local_engine = engine.Engine("radix-partitioning", partitions=partitions_local, hash=hash,
                                          hash_seed=hash_seed, device=device)
        self._recv = [None, None]
        self.launches = 0

    def upload(self, R, S):
        self.split_engine.upload(R, S)

    def split(self):
        """-> ([send tensor R, send tensor S] ordered by owner, counts[2][world], device ns)."""
        res = self.split_engine.join()
        self.launches += res["kernel_launches"]
        sends, counts = [], []
        for which in (0, 1):
            ptr, _, n = self.split_engine.device_partitions(which)
            b = self.split_engine.read_bounds(which, self.world).astype(np.int64)
            counts.append(np.diff(b))
            if n == 0:
                sends.append(self.torch.empty((0, 2), dtype=self.torch.int64, device=f"cuda:{self.device}"))
            else:
                sends.append(self.torch.as_tensor(_CudaView(ptr, n), device=f"cuda:{self.device}"))
        return sends, np.stack(counts), res["total_ns"]

    def recv_buffer(self, which, rows):
        buf = self._recv[which]
        if buf is None or buf.shape[0] < rows:
            buf = self.torch.empty((max(rows, 1) + max(rows, 1) // 16, 2), dtype=self.torch.int64,
                                   device=f"cuda:{self.device}")
            self._recv[which] = buf
        return buf[:rows]

    def local_join(self, recv_R, recv_S):
        self.torch.cuda.current_stream().synchronize()  # the all-to-all has landed
        self.local_engine.bind_device(recv_R.data_ptr() if recv_R.shape[0] else 0, recv_R.shape[0],
                                      recv_S.data_ptr() if recv_S.shape[0] else 0, recv_S.shape[0],
                                      keepalive=(recv_R, recv_S))
        res = self.local_engine.join()
        self.launches += res["kernel_launches"]
        return res["matches"], res

    def count_tensor(self, value):
        return self.torch.tensor([value], dtype=self.torch.int64, device=f"cuda:{self.device}")

    def int_tensor(self, array):
        return self.torch.as_tensor(np.ascontiguousarray(array, dtype=np.int64), device=f"cuda:{self.device}")

    def close(self):
        self.split_engine.close()
        self.local_engine.close()


class ShardedRadixJoin:
    """One rank of the sharded join. `dist` is torch.distributed (already initialised) or None for a
    single rank."""

    def __init__(self, dist, rank, world, backend):
        if world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.dist, self.rank, self.world, self.backend = dist, rank, world, backend
        self.last = {}

    def upload(self, R_shard, S_shard):
        self.backend.upload(R_shard, S_shard)

    def join(self) -> dict:
        dist, be, world = self.dist, self.backend, self.world
        t0 = time.perf_counter()
        sends, counts, split_ns = be.split()                      # counts[rel][owner]
        t1 = time.perf_counter()
        if world > 1:
            send_counts = be.int_tensor(counts.T.reshape(-1))     # [owner][rel]
            recv_counts = be.int_tensor(np.zeros(2 * world))
            dist.all_to_all_single(recv_counts, send_counts)
            rc = recv_counts.cpu().numpy().reshape(world, 2)      # [source][rel]
            recvs = []
            for which in (0, 1):
                out_rows = rc[:, which].tolist()
                buf = be.recv_buffer(which, int(sum(out_rows)))
                dist.all_to_all_single(buf, sends[which], output_split_sizes=out_rows,
                                       input_split_sizes=counts[which].tolist())
                recvs.append(buf)
        else:
            recvs = sends
        t2 = time.perf_counter()
        local_matches, res = be.local_join(recvs[0], recvs[1])
        t3 = time.perf_counter()
        total = be.count_tensor(local_matches)
        if world > 1:
            dist.all_reduce(total)
        matches = int(total.item())
        t4 = time.perf_counter()
        self.last = {"matches": matches, "local_matches": int(local_matches), "split_s": t1 - t0,
                     "exchange_s": t2 - t1, "local_s": t3 - t2, "reduce_s": t4 - t3, "total_s": t4 - t0,
                     "recv_rows": [int(r.shape[0]) for r in recvs], "send_bytes_remote":
                     int(16 * (counts.sum() - counts[:, self.rank].sum())) if world > 1 else 0,
                     "local_result": res, "split_device_ns": split_ns}
        return self.last

    def close(self):
        self.backend.close()


def shard_inputs(phj, rank, world, n_build, n_probe, skew, base_seed, batches):
    """Weak-scaling shards: R = keys 1..world*n_build cut by rank, S = Zipf over the whole key range."""
    Rp, Sp = phj.PinnedTuples(n_build), phj.PinnedTuples(n_probe)
    phj.fill_sequential(Rp.array, 1 + rank * n_build)
    Rp.array["payload"] += rank * n_build
    phj.fill_zipf(Sp.array, skew, 1, world * n_build, base_seed + 7919 * rank, batches)
    return Rp, Sp


def bench(args, dist, rank, world, local, make_inputs, workload_config, ClockSampler, metric, unit, measured_hbm_peak):
    """bench.py --gpus N (N > 1): every rank brings a 10 M x 200 M shard (weak scaling)."""
    import json

    import torch

    import partitionedhashjoin_b200 as phj
    n_build, n_probe = 10_000_000, 200_000_000
    Rp, Sp = shard_inputs(phj, rank, world, n_build, n_probe, args.skew, 12345, 64)
    backend = GpuBackend(world, local, partitions_local=args.partitions, hash=args.hash)
    job = ShardedRadixJoin(dist, rank, world, backend)
    job.upload(Rp.array, Sp.array)
    for _ in range(args.warmup):
        res = job.join()
    want = world * n_probe
    assert res["matches"] == want, (res["matches"], want)

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    launches0 = backend.launches
    with ClockSampler(local) as clocks:
        sync()
        t0 = time.perf_counter()
        parts = {"split_s": 0.0, "exchange_s": 0.0, "local_s": 0.0, "reduce_s": 0.0}
        for _ in range(args.steps):
            res = job.join()
            for k in parts:
                parts[k] += res[k]
        sync()
        elapsed = time.perf_counter() - t0
    t = torch.tensor([elapsed], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    assert res["matches"] == want

    # e2e: the host shards are uploaded inside the timed region
    e2e_steps = max(3, min(args.steps, 5))
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.upload(Rp.array, Sp.array)
        r2 = job.join()
    sync()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    assert r2["matches"] == want
    launches = torch.tensor([backend.launches - launches0], dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(launches)

    if rank == 0:
        n_tuples = world * (n_build + n_probe)
        ms = elapsed / args.steps * 1e3
        peak, peak_src = measured_hbm_peak()
        lr = res["local_result"]
        cfg = workload_config(args)
        cfg["workload"] = (f"radix join sharded over {world} B200: {world} x (10M x 200M) row shards = "
                           f"{world * 10}M x {world * 200}M, NCCL all-to-all partition shuffle, then local "
                           f"2-pass radix join ({args.partitions} partitions/GPU), {args.hash}, Zipf skew {args.skew}")
        cfg["parallelism"] = f"partition-sharded x{world}"
        exch_bytes = res["send_bytes_remote"]
        line = {
            "metric": metric, "value": n_tuples / (elapsed / args.steps), "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": cfg,
            "e2e": {"value": n_tuples / (float(te.item()) / e2e_steps), "unit": unit,
                    "h2d_bytes_per_step": 16 * n_tuples, "d2h_bytes_per_step": int(lr["d2h_bytes"] + 8 * (2 * world + 2)) * world,
                    "steps": e2e_steps},
            "gpu_launches": int(launches.item()),
            "roofline": {"bound": "hbm", "kernel": "local radix join (per GPU)", "unit": "GB/s", "peak": peak,
                         "peak_source": peak_src, "achieved": lr["hbm_bytes_alg"] / lr["total_ns"],
                         "frac": lr["hbm_bytes_alg"] / lr["total_ns"] / peak, "traffic": None},
            "phases_ms_rank0": {k[:-2]: v / args.steps * 1e3 for k, v in parts.items()},
            "shuffle": {"bytes_sent_remote_rank0": exch_bytes,
                        "GBps_rank0": exch_bytes / max(parts["exchange_s"] / args.steps, 1e-9) / 1e9,
                        "nvlink_peak_GBps": 770.0},
            "clocks": clocks.summary(), "matches": res["matches"], "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    backend.close()
    dist.barrier()
    dist.destroy_process_group()
