#!/bin/bash
# Round 2, GPU visit zc: final build (best candidate as a tuple, column loop unrolled by 4, per-lane bounds in registers): full GPU suite, bench line + reference arm,
# launch list of the bench command, full ncu capture of the default kernel (summary + per-source-line view), probe
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2zc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2zc.log
tail -5 gpurun_out/gpu_tests_r2zc.log
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2zc.json 2> gpurun_out/bench_r2zc.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench_r2zc.json; tail -5 gpurun_out/bench_r2zc.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2zc_reference.json 2>> gpurun_out/bench_r2zc.err
tail -c 1500 gpurun_out/bench_r2zc_reference.json
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2zc.txt 2>&1
cat gpurun_out/probe_r2zc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zc_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2zc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2zc \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2zc.log 2>&1
ncu -i /tmp/prof_r2zc.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2zc.csv 2>/dev/null
ncu -i /tmp/prof_r2zc.ncu-rep --page source --csv > gpurun_out/prof_source_r2zc.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2zc.csv gpurun_out/prof_source_r2zc.csv 8388608 > gpurun_out/r2zc_ncu_summary.txt 2>&1
head -70 gpurun_out/r2zc_ncu_summary.txt
for RS in 262144 65536 32768 8192 4096; do
  timeout 120 python tools/run_search.py --launches 2 --rounds $((8388608 / RS)) --round-size $RS | tail -1
done > gpurun_out/round_size_sweep_r2zc.txt 2>&1
cat gpurun_out/round_size_sweep_r2zc.txt
