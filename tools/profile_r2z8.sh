#!/bin/bash
# Round 2, 8-GPU visit z8: the N = 8 bench line of the final build (both arms as the driver launches them) and kao_solve on 8 GPUs
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 \
    > gpurun_out/bench_r2z8_n8.json 2> gpurun_out/bench_r2z8_n8.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r2z8_n8.json; tail -3 gpurun_out/bench_r2z8_n8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k kao_solve > gpurun_out/gpu_tests_r2z8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2z8.log
tail -4 gpurun_out/gpu_tests_r2z8.log
