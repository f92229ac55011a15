#!/bin/bash
# Round 2, second GPU visit: parity of the rewritten column-major evaluator on the hardware, timing of every
# built schedule, ncu --set full of the default one.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluators.py tests/test_gpu_parity.py tests/test_gpu_delta.py -m gpu -x -q > gpurun_out/gpu_tests_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2b.log
tail -5 gpurun_out/gpu_tests_r2b.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2b.txt 2>&1
cat gpurun_out/probe_r2b.txt
timeout 300 python tools/run_search.py --config 2 --probe --round-size 65536 > gpurun_out/probe_r2b_cfg2.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2b \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2b.log 2>&1
ncu -i /tmp/prof_r2b.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2b.csv 2>/dev/null
ncu -i /tmp/prof_r2b.ncu-rep --page source --csv > gpurun_out/prof_source_r2b.csv 2>/dev/null
ncu -i /tmp/prof_r2b.ncu-rep --page details > gpurun_out/prof_details_r2b.txt 2>/dev/null
ls -la gpurun_out
