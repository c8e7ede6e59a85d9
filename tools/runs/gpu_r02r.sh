# round 2, session 2, call 5 (2 GPUs): both count modes, bench --gpus 2, ncu of the NVLink scatter
set -x
timeout 900 python -m pytest tests/test_distributed.py -m gpu -x -q > gpurun_out/r02r_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 4 gpurun_out/r02r_disttests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02r_bench2.json 2> gpurun_out/r02r_bench2.err; echo "bench2 rc=$?"; tail -n 3 gpurun_out/r02r_bench2.err; cut -c1-300 gpurun_out/r02r_bench2.json
CONFIGS=0:4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/probe_dist.py > gpurun_out/r02r_probe_upfront.log 2>&1; echo "probe rc=$?"; grep "^==" gpurun_out/r02r_probe_upfront.log | cut -c1-200
FLAGS=0x800 CONFIGS=0:4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/probe_dist.py > gpurun_out/r02r_probe_piecewise.log 2>&1; echo "probe rc=$?"; grep "^==" gpurun_out/r02r_probe_piecewise.log | cut -c1-200
timeout 200 python tools/ncu_split.py 8 64 3 > gpurun_out/r02r_split_plain.log 2>&1; echo "split plain rc=$?"; cut -c1-300 gpurun_out/r02r_split_plain.log
timeout 200 python tools/ncu_split.py 2 64 3 >> gpurun_out/r02r_split_plain.log 2>&1; tail -n 3 gpurun_out/r02r_split_plain.log | cut -c1-300
timeout 500 ncu --set full --clock-control none --import-source on --devices 0 -k regex:radix_scatter -s 1 -c 1 -o gpurun_out/r02r_split python tools/ncu_split.py 8 64 2 > gpurun_out/r02r_split_ncu.log 2>&1; echo "ncu rc=$?"; tail -n 5 gpurun_out/r02r_split_ncu.log
