#!/bin/bash
# Round 2, GPU visit w: rack-field planes in the row pass (8 words per lane instead of 64), restart recipe for config 4:
# GPU suite, probe, full ncu capture, short bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2w.log
tail -12 gpurun_out/gpu_tests_r2w.log
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2w.txt 2>&1
cat gpurun_out/probe_r2w.txt
timeout 300 python tools/run_search.py --probe --config 2 --round-size 65536 > gpurun_out/probe_r2w_config2.txt 2>&1
cat gpurun_out/probe_r2w_config2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2w \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2w.log 2>&1
ncu -i /tmp/prof_r2w.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2w.csv 2>/dev/null
ncu -i /tmp/prof_r2w.ncu-rep --page source --csv > gpurun_out/prof_source_r2w.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2w.csv gpurun_out/prof_source_r2w.csv 8388608 > gpurun_out/r2w_ncu_summary.txt 2>&1
head -64 gpurun_out/r2w_ncu_summary.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config 4 > gpurun_out/bench_r2w_config4.json 2> gpurun_out/bench_r2w.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_r2w_config4.json
