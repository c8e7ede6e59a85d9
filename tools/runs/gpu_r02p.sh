# round 2, session 2, call 3: piece-wise count of the sharded join (1 rank: the whole pipeline incl. NCCL), segment / join-wave variants
set -x
timeout 900 python -m pytest tests/test_distributed.py tests/test_cli.py -m gpu -x -q > gpurun_out/r02p_disttests.log 2>&1; echo "dist tests rc=$?"; tail -n 8 gpurun_out/r02p_disttests.log
timeout 600 python tools/ab_scatter.py base segs24 segs32 joinw8 joinw12 > gpurun_out/r02p_ab.log 2>&1
cat gpurun_out/r02p_ab.log
CONFIGS=0:4,0:1 timeout 300 python tools/probe_dist.py > gpurun_out/r02p_probe_dist1.log 2>&1; echo "probe_dist rc=$?"; cut -c1-400 gpurun_out/r02p_probe_dist1.log | tail -40
