# round-2 evidence on ONE B200: tests, the bench line, the reference arm, ncu launch list + full captures
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -n 4 gpurun_out/r02m_gputests.log
timeout 600 python bench.py > gpurun_out/r02m_bench1.json 2> gpurun_out/r02m_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02m_ref.json 2> gpurun_out/r02m_ref.err; echo "ref rc=$?"
Q="--steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e"
python bench.py $Q > gpurun_out/r02m_plain.json 2> gpurun_out/r02m_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches.csv python bench.py $Q > gpurun_out/r02m_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|join_partitions|radix_histogram_full" -s 12 -c 4 -o gpurun_out/r02m_top python bench.py $Q > gpurun_out/r02m_ncu2.log 2>&1
python tools/ncu_join.py 64 0x80 > gpurun_out/r02m_l2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"pt_|radix_scatter" -s 3 -c 3 -o gpurun_out/r02m_l2 python tools/ncu_join.py 64 0x80 > gpurun_out/r02m_ncu3.log 2>&1
echo done
