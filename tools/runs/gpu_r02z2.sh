# round 2, the last seconds of GPU time: the bench line of the final default build (warp_peers via R2P)
timeout 40 python bench.py --steps 20 --warmup 3 --quick --no-cpu-baseline --no-e2e > gpurun_out/r02z_bench_quick.json 2> gpurun_out/r02z_bench_quick.err; echo "rc=$?"; cut -c1-300 gpurun_out/r02z_bench_quick.json
