#!/usr/bin/env python
"""Runs the persistent search kernel a few times on a BASELINE config (for ncu captures and quick timings):
python tools/run_search.py [--config 3] [--launches 3] [--rounds 32] [--round-size 262144] [--probe]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kafka_assignment_optimizer_b200 as kao  # noqa: E402
from kafka_assignment_optimizer_b200 import tuning  # noqa: E402

CONFIGS = {"2": (256, 32, 4, 3, 0), "3": (1000, 64, 8, 3, 0), "4": (1000, 64, 8, 3, 2), "5": (4096, 256, 16, 3, 0, 0.02, 5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="3")
    ap.add_argument("--launches", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=32)
    ap.add_argument("--round-size", type=int, default=1 << 18)
    ap.add_argument("--probe", action="store_true", help="time every built schedule (tuning.probe)")
    ap.add_argument("--delta", action="store_true")
    a = ap.parse_args()
    pb = kao.synthetic_problem(*CONFIGS[a.config])
    if a.probe:
        tuning.probe(pb, 0, a.rounds, a.round_size)
        return
    sess = kao.Session(pb)
    for i in range(a.launches):
        keys, ms = (sess.search_delta if a.delta else sess.search)(0x5EED, i * a.rounds, a.rounds, a.round_size)
        print(json.dumps({"config": a.config, "launch": i, "ms": ms, "candidates_per_s": a.rounds * a.round_size / ms * 1e3,
                          "last_key": sess.unpack_key(keys[-1])}), flush=True)
    sess.close()


if __name__ == "__main__":
    main()
