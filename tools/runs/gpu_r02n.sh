# round 2, session 2, call 1: all-warp counter scan + tile shapes of radix_scatter, parity of each build
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python tools/ab_scatter.py check > gpurun_out/r02n_ab.log 2>&1
tail -n 30 gpurun_out/r02n_ab.log
