#!/bin/bash
# Round 2, GPU visit v: cycle rounds in the candidate stream (MODEL 5): GPU suite, probe, how far the search gets on
# configs 3, 4 (full + delta) and 5' (delta), short bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2v.log
tail -12 gpurun_out/gpu_tests_r2v.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2v.txt 2>&1
cat gpurun_out/probe_r2v.txt
for C in cfg3:16384:600 cfg4:32768:3000 cfg4:32768:3000:7 cfg4:32768:6000:0x5EED:delta cfg5_p02:32768:12000:0x5EED:delta cfg5_p02:32768:40000:7:delta; do
  timeout 300 python tools/solve_configs.py $C 2>&1 | tail -1 | cut -c1-600
done > gpurun_out/solve_r2v.txt
cat gpurun_out/solve_r2v.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r2v.json
