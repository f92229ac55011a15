#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python tools/run_search.py --probe > gpurun_out/probe_r2e.txt 2>&1
cat gpurun_out/probe_r2e.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2e.log
tail -8 gpurun_out/gpu_tests_r2e.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r2e.json; tail -5 gpurun_out/bench_r2e.err
