#!/bin/bash
# Round 2, GPU visit k (2 GPUs): multi-GPU parity tests and the N=2 bench line of the term-plane build
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/gpu_tests_r2k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2k.log
tail -6 gpurun_out/gpu_tests_r2k.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_r2k_n2.json 2> gpurun_out/bench_r2k_n2.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench_r2k_n2.json; tail -5 gpurun_out/bench_r2k_n2.err
