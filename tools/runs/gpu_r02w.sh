# round 2, session 2, call 11 (2 GPUs): heavy-hitter handling as the default of the sharded join -- the multi-rank library tests
set -x
timeout 200 python -m pytest tests/test_distributed.py -m gpu -x -q -k "library_sharded_join_on_gpus or one_process or single_rank" > gpurun_out/r02w_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 12 gpurun_out/r02w_disttests.log | cut -c1-300
