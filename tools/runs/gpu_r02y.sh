# round 2, last call: smoke() and the bench's other_configs (incl. the one-pass + L2-table plans) on the final build
set -x
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02y_smoke.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02y_bench.json
