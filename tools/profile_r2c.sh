#!/bin/bash
# Round 2, GPU visit: batched per-lane generation + explicit LOP3 row pass.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_evaluators.py tests/test_gpu_parity.py tests/test_gpu_delta.py tests/test_gpu_configs.py -m gpu -q > gpurun_out/gpu_tests_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2c.log
tail -5 gpurun_out/gpu_tests_r2c.log

timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2c.txt 2>&1
cat gpurun_out/probe_r2c.txt
timeout 300 python tools/run_search.py --config 2 --probe --round-size 65536 > gpurun_out/probe_r2c_cfg2.txt 2>&1
for S in 1,11111,768 1,11133,512; do
TAG=$(echo $S | tr , _)
KAO_SCHEDULE=$S timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_$TAG \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2c_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2c_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_source_r2c_$TAG.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/prof_details_r2c_$TAG.txt 2>/dev/null
done
ls -la gpurun_out | tail -12
