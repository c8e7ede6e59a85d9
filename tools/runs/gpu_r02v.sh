# round 2, session 2, call 9 (8 GPUs): heavy-hitter digits kept local at 8 ranks: window-content test + Zipf 1.25 timing
set -x
timeout 300 python -m pytest tests/test_distributed.py -m gpu -x -q -k "library_sharded_join_on_gpus and 8" > gpurun_out/r02v_tests8.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r02v_tests8.log | cut -c1-300
ALPHA=1.25 CONFIGS=0:4,0:4:0x1000 JOINS=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/probe_dist.py > gpurun_out/r02v_probe_hot8.log 2>&1; echo "probe rc=$?"; grep "^==" gpurun_out/r02v_probe_hot8.log | cut -c1-700
