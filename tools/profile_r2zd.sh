#!/bin/bash
# Round 2, GPU visit zd: the bench line of the final build with the clock sampler waiting for NVML
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2zd.json 2> gpurun_out/bench_r2zd.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_r2zd.json; tail -3 gpurun_out/bench_r2zd.err
