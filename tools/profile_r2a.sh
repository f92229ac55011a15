#!/bin/bash
# Round 2, first capture: ncu --set full of the round-1 headline kernel (column-major, warp-only sync,
# five compressed popcount streams) and of its block-barrier sibling.  CSV exports land in gpurun_out/.
set -u
mkdir -p gpurun_out
for S in 1,2,768,1,0,0 0,2,768,1,0,0; do
  TAG=$(echo $S | tr , _)
  KAO_SCHEDULE=$S timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_persistent -s 2 -c 1 -o /tmp/prof_$TAG \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --evaluator column > gpurun_out/ncu_full_$TAG.log 2>&1
  ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_raw_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_source_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/prof_details_$TAG.txt 2>/dev/null
done
KAO_SCHEDULE=1,2,768,1,0,0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2a.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --evaluator column > gpurun_out/ncu_launches_r2a.log 2>&1
KAO_SCHEDULE=1,2,768,1,0,0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --evaluator column > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
ls -la gpurun_out
