#!/bin/bash
# Round 2, GPU visit y: final build of the round (rack-field planes, delta bytes, cycle rounds, 896 threads): full GPU suite, bench line + reference arm,
# launch list of the bench command, full ncu capture of the default kernel (summary + per-source-line view), probe
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2y.log
tail -5 gpurun_out/gpu_tests_r2y.log
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench_r2y.json; tail -5 gpurun_out/bench_r2y.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2y_reference.json 2>> gpurun_out/bench_r2y.err
tail -c 1500 gpurun_out/bench_r2y_reference.json
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2y.txt 2>&1
cat gpurun_out/probe_r2y.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2y.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2y \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2y.log 2>&1
ncu -i /tmp/prof_r2y.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2y.csv 2>/dev/null
ncu -i /tmp/prof_r2y.ncu-rep --page source --csv > gpurun_out/prof_source_r2y.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2y.csv gpurun_out/prof_source_r2y.csv 8388608 > gpurun_out/r2y_ncu_summary.txt 2>&1
head -70 gpurun_out/r2y_ncu_summary.txt
for RS in 262144 65536 32768 8192 4096; do
  timeout 120 python tools/run_search.py --launches 2 --rounds $((8388608 / RS)) --round-size $RS | tail -1
done > gpurun_out/round_size_sweep_r2y.txt 2>&1
cat gpurun_out/round_size_sweep_r2y.txt
timeout 400 python tools/solve_configs.py cfg5_p02:32768:12000:0x5EED:delta > gpurun_out/solve_cfg5_delta_r2y.txt 2>&1
tail -c 700 gpurun_out/solve_cfg5_delta_r2y.txt
