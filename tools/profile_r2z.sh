#!/bin/bash
# Round 2, GPU visit z: per-lane bounds in registers; NVML clock sampler in bench.py: probe, GPU suite, bench line
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2z.txt 2>&1
cat gpurun_out/probe_r2z.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2z.log
tail -5 gpurun_out/gpu_tests_r2z.log
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2z.json; tail -5 gpurun_out/bench_r2z.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2z \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2z.log 2>&1
ncu -i /tmp/prof_r2z.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2z.csv 2>/dev/null
ncu -i /tmp/prof_r2z.ncu-rep --page source --csv > gpurun_out/prof_source_r2z.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2z.csv gpurun_out/prof_source_r2z.csv 8388608 > gpurun_out/r2z_ncu_summary.txt 2>&1
head -40 gpurun_out/r2z_ncu_summary.txt
