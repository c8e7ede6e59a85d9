# round 2, session 2, call 7 (8 GPUs): piece-wise count with the probe released in front of the next scatter
set -x
CONFIGS=0:4:0x400,0:4:0x800,0:6:0x800,0:4:0x800:100,0:4:0x800:116,0:3:0x800 JOINS=5 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/probe_dist.py > gpurun_out/r02t_sweep8.log 2>&1; echo "sweep rc=$?"; grep "^==" gpurun_out/r02t_sweep8.log | cut -c1-230
timeout 400 python -m pytest tests/test_distributed.py -m gpu -x -q -k "library_sharded_join_on_gpus and 8" > gpurun_out/r02t_tests8.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r02t_tests8.log
