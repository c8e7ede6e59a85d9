#!/bin/bash
# One GPU call: parity subset, bench line, then the ncu launch list + full capture of the same command.
mkdir -p gpurun_out
F="--steps 2 --warmup 3 --quick --no-cpu-baseline --no-e2e"
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partition_layout_skewed or full_size or (partition_layout_equals and 4096)" > gpurun_out/t5.log 2>&1
echo "rc=$?" >> gpurun_out/t5.log
tail -n 3 gpurun_out/t5.log
timeout 80 python bench.py --steps 20 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r01f.json')); print(d['ms_per_step'], d['kernel_us'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
timeout 60 python bench.py $F > gpurun_out/r01f_plain.json 2> gpurun_out/r01f_plain.err || exit 1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py $F > gpurun_out/r01f_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|join_partitions|radix_histogram" -s 8 -c 4 -o gpurun_out/prof_r01f_top -f python bench.py $F > gpurun_out/r01f_ncu2.log 2>&1
echo "ncu2 rc=$?"
python tools/make_profile_summary.py r01f gpurun_out/r01f_launches.csv gpurun_out/prof_r01f_top.ncu-rep "$F" > /dev/null 2> gpurun_out/r01f_summary.err && cp profiles/r01f_ncu_summary.md gpurun_out/
ls -la gpurun_out/prof_r01f_top.ncu-rep
