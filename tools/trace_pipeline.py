"""Diagnostic: per-call host times and per-kernel device times of one pipelined multi-GPU join.
torchrun --nproc-per-node N tools/trace_pipeline.py [chunks] [split_ctas]"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PHJ_KERNEL_TIMES"] = "1"
import partitionedhashjoin_b200 as phj  # noqa: E402
from partitionedhashjoin_b200 import multigpu  # noqa: E402

chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 96
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Rp, Sp = multigpu.shard_inputs(phj, rank, world, 10_000_000, 200_000_000, 0.01, 12345, 64)
be = multigpu.PipelinedGpuBackend(world, local, partitions_local=4096, chunks=chunks, split_ctas=ctas,
                                  copy_engines=(len(sys.argv) <= 3 or sys.argv[3] != "sm"))
job = multigpu.PipelinedShardedRadixJoin(dist, rank, world, be)
job.upload(Rp.array, Sp.array)
for _ in range(3):
    job.join()
torch.cuda.synchronize()
dist.barrier()
be.trace = []
t0 = time.perf_counter()
res = job.join()
t1 = time.perf_counter()
if rank == 0:
    print(f"chunks={chunks} ctas={ctas} total {1e3 * (t1 - t0):.3f} ms; split {1e3 * res['split_s']:.3f} sizes {1e3 * res['sizes_s']:.3f} "
          f"pipeline {1e3 * res['pipeline_s']:.3f} reduce {1e3 * res['reduce_s']:.3f}")
    for ev in sorted(be.trace, key=lambda e: e[2]):
        if ev[0] == "scatter":
            _, c, a, b, kt = ev
            print(f"  scatter c={c}: host [{1e3 * (a - t0):7.3f} .. {1e3 * (b - t0):7.3f}] kernels " +
                  " ".join(f"{n}={ns / 1e3:.0f}" for n, ns in kt))
        else:
            _, c, a, m, b, kt = ev
            print(f"  local   c={c}: host [{1e3 * (a - t0):7.3f} bind {1e3 * (m - a):.3f} .. {1e3 * (b - t0):7.3f}] kernels " +
                  " ".join(f"{n}={ns / 1e3:.0f}" for n, ns in kt))
job.close()
dist.barrier()
dist.destroy_process_group()
