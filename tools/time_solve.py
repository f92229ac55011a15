#!/usr/bin/env python
"""Times kao_solve with host buffers on N GPUs of this process (what bench.py's `e2e` measures):
python tools/time_solve.py [--gpus 2] [--calls 5] [--rounds 32] [--round-size-per-gpu 262144]
KAO_TRACE=1 prints where a multi-GPU solve spends its host time (stderr)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kafka_assignment_optimizer_b200 as kao  # noqa: E402
from kafka_assignment_optimizer_b200 import optimizer as kopt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--calls", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=32)
    ap.add_argument("--round-size-per-gpu", type=int, default=1 << 18)
    a = ap.parse_args()
    pb = kao.synthetic_problem(1000, 64, 8, 3, 0)
    kopt.solve(pb, 1, 2, 1 << 12, 0, n_gpus=a.gpus)              # contexts, modules
    for k in range(a.calls):
        t0 = time.perf_counter()
        r = kopt.solve(pb, 0x5EED + k, a.rounds, a.round_size_per_gpu * a.gpus, 0, n_gpus=a.gpus)
        wall = (time.perf_counter() - t0) * 1e3
        print("call %d: wall %.2f ms, total_ms %.2f, device_ms %.2f, n_gpus %d, candidates/s %.3e" % (
            k, wall, r.total_ms, r.device_ms, r.n_gpus, a.rounds * a.round_size_per_gpu * a.gpus / wall * 1e3), flush=True)


if __name__ == "__main__":
    main()
