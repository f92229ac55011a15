#!/bin/bash
# Round 2, GPU visit j: full GPU suite + bench line of the term-plane build; cost of small rounds; config 5 delta solve
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2j.log
tail -6 gpurun_out/gpu_tests_r2j.log
timeout 900 python bench.py --steps 12 --warmup 3 > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench_r2j.json; tail -5 gpurun_out/bench_r2j.err
for RS in 262144 65536 32768 8192; do
  timeout 120 python tools/run_search.py --launches 2 --rounds $((8388608 / RS)) --round-size $RS | tail -1
done > gpurun_out/round_size_sweep_r2j.txt 2>&1
cat gpurun_out/round_size_sweep_r2j.txt
timeout 400 python tools/solve_configs.py cfg5_p02:32768:12000:0x5EED:delta > gpurun_out/solve_cfg5_delta_r2j.txt 2>&1
tail -c 1200 gpurun_out/solve_cfg5_delta_r2j.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2j.log 2>&1
