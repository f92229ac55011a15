#!/bin/bash
# Round 2, GPU visit g (after the container was re-created): the evidence of the current default build --
# bench line, launch list of the bench command, one full ncu capture of the headline kernel, GPU tests.
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2g.txt 2>&1
cat gpurun_out/probe_r2g.txt
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench_r2g.json; tail -5 gpurun_out/bench_r2g.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2g \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2g.log 2>&1
ncu -i /tmp/prof_r2g.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2g.csv 2>/dev/null
ncu -i /tmp/prof_r2g.ncu-rep --page source --csv > gpurun_out/prof_source_r2g.csv 2>/dev/null
ncu -i /tmp/prof_r2g.ncu-rep --page details > gpurun_out/prof_details_r2g.txt 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2g.csv gpurun_out/prof_source_r2g.csv 8388608 > gpurun_out/r2g_ncu_summary.txt 2>&1
head -60 gpurun_out/r2g_ncu_summary.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2g.log
tail -8 gpurun_out/gpu_tests_r2g.log
