# round-2 evidence on ONE B200 (final build): tests, the bench line, ncu launch list + full capture of the top kernels
# (the reference arm and the L2-table / no-partitioning captures of the same round: gpurun_out/r02m_*, tools/runs/gpu_r02*.sh)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -n 4 gpurun_out/r02x_gputests.log
timeout 600 python bench.py > gpurun_out/r02x_bench1.json 2> gpurun_out/r02x_bench1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02x_bench1.json
Q="--steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e"
python bench.py $Q > gpurun_out/r02x_plain.json 2> gpurun_out/r02x_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv python bench.py $Q > gpurun_out/r02x_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|join_partitions|radix_histogram_full" -s 12 -c 4 -o gpurun_out/r02x_top python bench.py $Q > gpurun_out/r02x_ncu2.log 2>&1
echo done
