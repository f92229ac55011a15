#!/bin/bash
# Round 2, GPU visit i: unrolled column loop (XOR swizzle), popcount-level variants
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_evaluators.py -m gpu -q -x > gpurun_out/gpu_tests_r2i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2i.log
tail -4 gpurun_out/gpu_tests_r2i.log
timeout 400 python tools/run_search.py --probe > gpurun_out/probe_r2i.txt 2>&1
cat gpurun_out/probe_r2i.txt
timeout 300 python tools/run_search.py --config 2 --probe --round-size 65536 > gpurun_out/probe_r2i_cfg2.txt 2>&1
cat gpurun_out/probe_r2i_cfg2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2i \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2i.log 2>&1
ncu -i /tmp/prof_r2i.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2i.csv 2>/dev/null
ncu -i /tmp/prof_r2i.ncu-rep --page source --csv > gpurun_out/prof_source_r2i.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2i.csv gpurun_out/prof_source_r2i.csv 8388608 > gpurun_out/r2i_ncu_summary.txt 2>&1
head -64 gpurun_out/r2i_ncu_summary.txt
