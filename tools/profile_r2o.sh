#!/bin/bash
# Round 2, GPU visit o: how far does the search get on config 5' (4096 x 256 x 16, 2 % re-placed) with more rounds / other seeds
set -u
mkdir -p gpurun_out
for SEED in 0x5EED 7 11; do
  timeout 300 python tools/solve_configs.py cfg5_p02:32768:60000:$SEED:delta 2>&1 | tail -1 | cut -c1-700
done > gpurun_out/solve_cfg5_long_r2o.txt
cat gpurun_out/solve_cfg5_long_r2o.txt
timeout 300 python tools/solve_configs.py cfg5_p02:8192:200000:0x5EED:delta 2>&1 | tail -1 | cut -c1-700 > gpurun_out/solve_cfg5_small_r2o.txt
cat gpurun_out/solve_cfg5_small_r2o.txt
