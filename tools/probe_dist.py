"""Ad-hoc GPU probe of the library's sharded join (phj_dist_*): per-rank device times at the bench shapes.

    python tools/probe_dist.py                               one rank on GPU 0
    torchrun --nproc-per-node N tools/probe_dist.py          one rank per GPU

Env: NR / NS tuples per rank, ALPHA, CONFIGS (partitions:chunks,...), JOINS, FLAGS (phj_config.flags, e.g. 0x800).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PHJ_KERNEL_TIMES", "1")
import torch  # noqa: E402

import partitionedhashjoin_b200 as phj  # noqa: E402
from partitionedhashjoin_b200 import multigpu  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nr, ns = int(os.environ.get("NR", 10_000_000)), int(os.environ.get("NS", 200_000_000))
    alpha = float(os.environ.get("ALPHA", 0.01))
    joins = int(os.environ.get("JOINS", 6))
    dR = phj.DeviceTuples(nr, local).fill_sequential(1 + rank * nr)
    dS = phj.DeviceTuples(ns, local).fill_zipf(alpha, 1, world * nr, 12345 + 7919 * rank, 1 << 14)
    for cfg in os.environ.get("CONFIGS", "0:4,0:1,0:8,128:4,256:4").split(","):
        # partitions : chunks [: flags [: scatter SMs]]
        f = [int(x, 0) for x in cfg.split(":")] + [0, 0]
        parts, chunks, flags, split_ctas = f[0], f[1], f[2] or int(os.environ.get("FLAGS", "0"), 0), f[3]
        if parts and parts < world:
            continue
        job = multigpu.ShardedJoin(dist, rank, world, local, partitions=parts, chunks=chunks, flags=flags,
                                   split_ctas=split_ctas)
        job.bind_device(dR.ptr, nr, dS.ptr, ns, keepalive=(dR, dS))
        best = None
        for _ in range(joins):
            res = job.join()
            assert res["matches"] == world * ns, res
            if best is None or res["total_ns"] < best["total_ns"]:
                best, kt = res, job.kernel_trace()
        if world > 1:
            t = torch.tensor([best["total_ns"]], dtype=torch.int64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            worst = int(t.item())
        else:
            worst = best["total_ns"]
        if rank == 0:
            lay = job.info()
            print(f"== world {world} partitions {lay['digits']} ({lay['local_partitions']}/GPU) chunks {lay['chunks']} "
                  f"flags {flags:#x} scatter SMs {split_ctas or 'auto'}: "
                  f"total {worst / 1e6:.3f} ms (rank 0: {best['total_ns'] / 1e6:.3f}) count {best['count_ns'] / 1e6:.3f} "
                  f"shuffle {best['shuffle_ns'] / 1e6:.3f} build {best['build_ns'] / 1e6:.3f} probe {best['probe_ns'] / 1e6:.3f} "
                  f"-> {world * (nr + ns) / (worst / 1e9) / 1e9:.2f} G tuples/s; sent {best['shuffle_bytes'] / 1e9:.2f} GB "
                  f"= {best['shuffle_bytes'] / max(best['shuffle_ns'], 1):.0f} GB/s; {json.dumps(lay)}")
            for name, b, e in kt:
                print(f"     {name:28s} {b / 1e3:9.1f} .. {e / 1e3:9.1f} us  ({(e - b) / 1e3:8.1f})")
            sys.stdout.flush()
        job.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
