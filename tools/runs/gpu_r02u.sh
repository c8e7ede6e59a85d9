# round 2, session 2, call 8 (2 GPUs): heavy-hitter digits kept local in the library's sharded join
set -x
timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q -k "library_sharded_join or one_process" > gpurun_out/r02u_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 30 gpurun_out/r02u_disttests.log | cut -c1-250
ALPHA=1.25 CONFIGS=0:4,0:4:0x1000 JOINS=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/probe_dist.py > gpurun_out/r02u_probe_hot2.log 2>&1; echo "probe rc=$?"; grep -v "^\[\|OMP\|\*\*\*\|^$" gpurun_out/r02u_probe_hot2.log | cut -c1-260 | tail -50
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02u_bench2.json 2> gpurun_out/r02u_bench2.err; echo "bench2 rc=$?"; tail -n 5 gpurun_out/r02u_bench2.err; cut -c1-200 gpurun_out/r02u_bench2.json
