"""Runs the BASELINE.json configurations through the engine on one GPU and prints what the search
reached next to the exact optimum where it is known (tests/golden/optima.json)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kafka_assignment_optimizer_b200 as kao  # noqa: E402

CONFIGS = {
    "cfg2_rm2": ((256, 32, 4, 3, 2), 1 << 14, 400),
    "cfg3": ((1000, 64, 8, 3, 0), 1 << 15, 600),
    "cfg4": ((1000, 64, 8, 3, 2), 1 << 15, 3000),
    "cfg5_p02": ((4096, 256, 16, 3, 0, 0.02, 5), 1 << 15, 1500),
}


def main():
    opt = json.load(open(os.path.join(ROOT, "tests", "golden", "optima.json")))
    names = sys.argv[1:] or list(CONFIGS)
    for name in names:
        over = name.split(":")                     # name[:round_size[:rounds[:seed]]]
        name = over[0]
        args, size, rounds = CONFIGS[name]
        if len(over) > 1:
            size = int(over[1])
        if len(over) > 2:
            rounds = int(over[2])
        seed = int(over[3], 0) if len(over) > 3 else 0x5EED
        delta = len(over) > 4 and over[4] == "delta"
        pb = kao.synthetic_problem(*args)
        sess = kao.Session(pb)
        t0 = time.perf_counter()
        done, first_feasible, best_round = 0, None, None
        last = None
        chunk = max(100, rounds // 20)
        dev_ms = 0.0
        while done < rounds:
            keys, ms = (sess.search_delta if delta else sess.search)(seed, done, chunk, size)
            dev_ms += ms
            for i, k in enumerate(keys):
                v, o, _ = sess.unpack_key(k)
                if v == 0 and first_feasible is None:
                    first_feasible = done + i
                if last is None or (v, -o) < last:
                    last, best_round = (v, -o), done + i
            done += chunk
        reps, viol, obj, moves = sess.get_base()
        e = opt.get(name, {})
        print(json.dumps({"config": name, "args": args, "round_size": size, "rounds": rounds, "seed": seed, "mode": "delta" if delta else "full",
                          "candidates": rounds * size, "violation": viol, "objective": obj, "moves": moves,
                          "exact_objective": e.get("objective"), "exact_moves": e.get("moves"),
                          "first_feasible_round": first_feasible, "last_improving_round": best_round,
                          "device_s": round(dev_ms / 1e3, 3), "wall_s": round(time.perf_counter() - t0, 3),
                          "candidates_per_s": round(rounds * size / (dev_ms / 1e3)),
                          "stats": sess.stats()}), flush=True)
        sess.close()


if __name__ == "__main__":
    main()
