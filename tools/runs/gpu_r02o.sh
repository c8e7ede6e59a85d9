# round 2, session 2, call 2: fused small kernels (one-launch scan, plan + empty bounds, align + counts, oversize in the
# join kernel, one result copy), leader-free counter store, segments per SM
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02o_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -n 5 gpurun_out/r02o_gputests.log
timeout 600 python tools/ab_scatter.py base allwrite segs16 segs16_allwrite > gpurun_out/r02o_ab.log 2>&1
cat gpurun_out/r02o_ab.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --quick > gpurun_out/r02o_bench_quick.json 2> gpurun_out/r02o_bench_quick.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02o_bench_quick.json
