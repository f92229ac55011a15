#!/bin/bash
# Round 2, GPU visit n: generator scans the transposed planes (no per-round tables): parity, round-size sweep, probe
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluators.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/gpu_tests_r2n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2n.log
tail -5 gpurun_out/gpu_tests_r2n.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2n.txt 2>&1
cat gpurun_out/probe_r2n.txt
for RS in 262144 65536 32768 8192; do
  timeout 120 python tools/run_search.py --launches 2 --rounds $((8388608 / RS)) --round-size $RS | tail -1
done > gpurun_out/round_size_sweep_r2n.txt 2>&1
cat gpurun_out/round_size_sweep_r2n.txt
python tools/time_solve.py --gpus 1 --calls 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2n \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2n.log 2>&1
ncu -i /tmp/prof_r2n.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2n.csv 2>/dev/null
ncu -i /tmp/prof_r2n.ncu-rep --page source --csv > gpurun_out/prof_source_r2n.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2n.csv gpurun_out/prof_source_r2n.csv 8388608 > gpurun_out/r2n_ncu_summary.txt 2>&1
head -60 gpurun_out/r2n_ncu_summary.txt
