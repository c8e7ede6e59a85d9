# round 2, session 2, call 10 (2 GPUs): heavy-hitter handling as the default of the sharded join -- the multi-rank tests
set -x
timeout 600 python -m pytest tests/test_distributed.py tests/test_cli.py -m gpu -x -q > gpurun_out/r02w_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 12 gpurun_out/r02w_disttests.log | cut -c1-300
