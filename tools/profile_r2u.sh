#!/bin/bash
# Round 2, GPU visit u: column loop with two chunks in flight (sync = 3) against the plain rolled loop, 640..1024 threads;
# full GPU suite; full ncu capture of (2, 0x22, 896) and (3, 0x22, 896)
set -u
mkdir -p gpurun_out
timeout 300 python tools/run_search.py --probe > gpurun_out/probe_r2u.txt 2>&1
cat gpurun_out/probe_r2u.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r2u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2u.log
tail -5 gpurun_out/gpu_tests_r2u.log
timeout 300 python tools/run_search.py --probe --config 2 --round-size 65536 > gpurun_out/probe_r2u_config2.txt 2>&1
cat gpurun_out/probe_r2u_config2.txt
for S in 2,22,896 3,22,896; do
  N=$(echo $S | tr -d ,)
  KAO_SCHEDULE=$S timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2u_$N \
      python tools/run_search.py --launches 2 > gpurun_out/ncu_full_r2u_$N.log 2>&1
  ncu -i /tmp/prof_r2u_$N.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2u_$N.csv 2>/dev/null
  ncu -i /tmp/prof_r2u_$N.ncu-rep --page source --csv > gpurun_out/prof_source_r2u_$N.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_raw_r2u_$N.csv gpurun_out/prof_source_r2u_$N.csv 8388608 > gpurun_out/r2u_ncu_summary_$N.txt 2>&1
  head -40 gpurun_out/r2u_ncu_summary_$N.txt
done
