#!/bin/bash
# Round 2, GPU visit r: probe (incl. the rolled column loop), full GPU suite, bench line, reference arm
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2r.txt 2>&1
cat gpurun_out/probe_r2r.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2r.log
tail -5 gpurun_out/gpu_tests_r2r.log
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench_r2r.json; tail -5 gpurun_out/bench_r2r.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2r_reference.json 2>> gpurun_out/bench_r2r.err
tail -c 1500 gpurun_out/bench_r2r_reference.json
