#!/bin/bash
# Round 2, GPU visit za: where do the ~14 us of fixed cost per round go — full ncu capture of a launch of 2048 rounds x 4096
# candidates (stall samples by source line); smoke(); wide-open rack bound test
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_evaluators.py -m gpu -q 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 1 -c 1 -o /tmp/prof_r2za \
    python tools/run_search.py --launches 2 --rounds 2048 --round-size 4096 > gpurun_out/ncu_full_r2za.log 2>&1
ncu -i /tmp/prof_r2za.ncu-rep --page raw --csv > gpurun_out/prof_raw_r2za.csv 2>/dev/null
ncu -i /tmp/prof_r2za.ncu-rep --page source --csv > gpurun_out/prof_source_r2za.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_raw_r2za.csv gpurun_out/prof_source_r2za.csv 8388608 > gpurun_out/r2za_ncu_summary.txt 2>&1
head -45 gpurun_out/r2za_ncu_summary.txt
