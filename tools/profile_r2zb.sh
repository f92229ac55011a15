#!/bin/bash
# Round 2, GPU visit zb: best candidate of a warp tracked as (violation, cost, index) and packed once per round; column loop
# unrolled by 2 / 4 as schedules: probe, GPU suite, full ncu capture of the default
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2zb.txt 2>&1
cat gpurun_out/probe_r2zb.txt
timeout 300 python tools/run_search.py --probe --config 2 --round-size 65536 > gpurun_out/probe_r2zb_config2.txt 2>&1
cat gpurun_out/probe_r2zb_config2.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2zb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2zb.log
tail -5 gpurun_out/gpu_tests_r2zb.log
