# round 2, session 2, call 4 (2 GPUs): multi-rank tests, bench --gpus 2, timeline of the piece-wise count, ncu of the NVLink scatter
set -x
timeout 900 python -m pytest tests/test_distributed.py tests/test_cli.py -m gpu -x -q > gpurun_out/r02q_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 6 gpurun_out/r02q_disttests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02q_bench2.json 2> gpurun_out/r02q_bench2.err; echo "bench2 rc=$?"; tail -n 3 gpurun_out/r02q_bench2.err; cut -c1-600 gpurun_out/r02q_bench2.json
CONFIGS=0:4,0:1,0:6 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/probe_dist.py > gpurun_out/r02q_probe_dist2.log 2>&1; echo "probe rc=$?"; cut -c1-330 gpurun_out/r02q_probe_dist2.log | grep -v "^\[" | head -60
timeout 200 python tools/ncu_split.py 8 64 3 > gpurun_out/r02q_split_plain.log 2>&1; echo "split plain rc=$?"; cut -c1-400 gpurun_out/r02q_split_plain.log
timeout 200 python tools/ncu_split.py 2 64 3 >> gpurun_out/r02q_split_plain.log 2>&1; tail -n 3 gpurun_out/r02q_split_plain.log | cut -c1-300
timeout 500 ncu --set full --clock-control none --import-source on --devices 0 -k regex:radix_scatter -s 2 -c 2 -o gpurun_out/r02q_split python tools/ncu_split.py 8 64 2 > gpurun_out/r02q_split_ncu.log 2>&1; echo "ncu rc=$?"; tail -n 5 gpurun_out/r02q_split_ncu.log
