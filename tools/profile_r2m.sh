#!/bin/bash
# Round 2, GPU visit m: generator scans the transposed planes (no per-round tables): parity, round-size sweep, probe
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluators.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/gpu_tests_r2m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2m.log
tail -5 gpurun_out/gpu_tests_r2m.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2m.txt 2>&1
cat gpurun_out/probe_r2m.txt
for RS in 262144 65536 32768 8192; do
  timeout 120 python tools/run_search.py --launches 2 --rounds $((8388608 / RS)) --round-size $RS | tail -1
done > gpurun_out/round_size_sweep_r2m.txt 2>&1
cat gpurun_out/round_size_sweep_r2m.txt
python tools/time_solve.py --gpus 1 --calls 3
