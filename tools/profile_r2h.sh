#!/bin/bash
# Round 2, GPU visit h: column-major evaluator with term planes (sparse objective): parity, probe, full ncu capture
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluators.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/gpu_tests_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2h.log
tail -8 gpurun_out/gpu_tests_r2h.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2h.txt 2>&1
cat gpurun_out/probe_r2h.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2h \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2h.log 2>&1
ncu -i /tmp/prof_r2h.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2h.csv 2>/dev/null
ncu -i /tmp/prof_r2h.ncu-rep --page source --csv > gpurun_out/prof_source_r2h.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2h.csv gpurun_out/prof_source_r2h.csv 8388608 > gpurun_out/r2h_ncu_summary.txt 2>&1
head -64 gpurun_out/r2h_ncu_summary.txt
