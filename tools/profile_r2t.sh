#!/bin/bash
# Round 2, GPU visit t: column pass without substitution (base planes summed in full + one delta byte per slot from the
# generating thread): schedule probe (incl. 1024 threads, unrolled loop at 768), parity suites, full ncu capture
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2t.txt 2>&1
cat gpurun_out/probe_r2t.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r2t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2t.log
tail -5 gpurun_out/gpu_tests_r2t.log
timeout 300 python tools/run_search.py --probe --config 2 --round-size 65536 > gpurun_out/probe_r2t_config2.txt 2>&1
cat gpurun_out/probe_r2t_config2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2t \
    python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2t.log 2>&1
ncu -i /tmp/prof_r2t.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2t.csv 2>/dev/null
ncu -i /tmp/prof_r2t.ncu-rep --page source --csv > gpurun_out/prof_source_r2t.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2t.csv gpurun_out/prof_source_r2t.csv 8388608 > gpurun_out/r2t_ncu_summary.txt 2>&1
head -64 gpurun_out/r2t_ncu_summary.txt
