#!/bin/bash
# Round 2, 2-GPU visit x2: multi-GPU tests (sharded search, kao_solve n_gpus = 2, restarts side by side) and the N = 2 bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/gpu_tests_r2x2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2x2.log
tail -6 gpurun_out/gpu_tests_r2x2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 \
    > gpurun_out/bench_r2x2_n2.json 2> gpurun_out/bench_r2x2_n2.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench_r2x2_n2.json; tail -3 gpurun_out/bench_r2x2_n2.err
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import kafka_assignment_optimizer_b200 as kao
from kafka_assignment_optimizer_b200 import optimizer as kopt
pb = kao.synthetic_problem(1000, 64, 8, 3, 2)
kw = dict(seed=7, rounds=400, round_size=1 << 12, patience=150, restarts=12)
for n, spread in ((1, False), (2, True), (2, False)):
    kopt.solve(pb, n_gpus=n, spread_restarts=spread, **dict(kw, restarts=2, rounds=20))
    t = time.perf_counter(); r = kopt.solve(pb, n_gpus=n, spread_restarts=spread, **kw); dt = time.perf_counter() - t
    print("config 4, 12 restarts: n_gpus=%d spread=%s -> objective %d moves %d rounds %d, %.1f ms wall (%.1f ms device)" % (n, spread, r.objective, r.moves, r.rounds, dt * 1e3, r.device_ms), flush=True)
PY
