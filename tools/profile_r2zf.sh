#!/bin/bash
# Round 2, GPU visit zf: RF bit masks as kernel parameters (constant-bank operands of the row pass): probe + GPU suite
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2zf.txt 2>&1
cat gpurun_out/probe_r2zf.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r2zf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2zf.log
tail -3 gpurun_out/gpu_tests_r2zf.log
