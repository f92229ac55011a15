#!/bin/bash
# 2 GPUs: multi-GPU parity tests, the N=2 bench line, a delta search on config 5
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q > gpurun_out/gpu_tests_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2f.log
tail -12 gpurun_out/gpu_tests_r2f.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_r2f_n2.json 2> gpurun_out/bench_r2f_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_r2f_n2.json; tail -5 gpurun_out/bench_r2f_n2.err
timeout 300 python tools/solve_configs.py cfg5_p02:32768:6000:0x5EED:delta > gpurun_out/solve_cfg5_delta.txt 2>&1
cat gpurun_out/solve_cfg5_delta.txt | tail -3
