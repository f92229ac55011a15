"""Generates the committed fixtures under tests/golden/ from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference snapshot has no code, so nothing here is
produced by importing the reference; the fixtures pin (a) HiGHS optima of the BASELINE.json
configs, (b) per-candidate keys / trajectories of the plain-C restatement."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model as m, ref  # noqa: E402
from problems import SHAPES  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def optima():
    out = {}
    cfgs = {"readme": None, "cfg2": (256, 32, 4, 3, 0), "cfg2_rm2": (256, 32, 4, 3, 2),
            "cfg3": (1000, 64, 8, 3, 0), "cfg4": (1000, 64, 8, 3, 2)}
    for name, args in cfgs.items():
        pb = m.readme_problem() if args is None else m.synthetic_problem(*args)
        s = m.solve_exact(pb)
        out[name] = {"args": args, "status": s.status, "objective": s.objective, "moves": s.moves,
                     "solve_s": round(s.solve_s, 3)}
        print(name, out[name], flush=True)
    # config 5 takes minutes: only with --cfg5
    if "--cfg5" in sys.argv:
        pb = m.synthetic_problem(4096, 256, 16, 3, 0, perturb=0.02, seed=5)
        s = m.solve_exact(pb)
        out["cfg5_p02"] = {"args": (4096, 256, 16, 3, 0, 0.02, 5), "status": s.status,
                           "objective": s.objective, "moves": s.moves, "solve_s": round(s.solve_s, 3)}
    return out


def streams():
    """keys of the first 192 candidates + identity of round 2 (a free round) and of the first 192 candidates of round 3
    (a cycle round, docs/MODEL.md 5), and an 8-round trajectory, per shape
    (packed with the per-problem key layout of docs/MODEL.md 3: obj_bits is stored next to them)"""
    out = {}
    for name in sorted(SHAPES):
        pb = SHAPES[name]()
        r = ref.Ref(pb)
        bits, ld = r.init_base()
        base = r.decode(bits, ld)
        v, o = r.evaluate(bits, ld)
        keys = r.candidate_keys(bits, ld, 0xC0FFEE, 2, 1024, 0, 192)
        last = r.candidate_keys(bits, ld, 0xC0FFEE, 2, 1024, 1023, 1)
        keys3 = r.candidate_keys(bits, ld, 0xC0FFEE, 3, 1024, 0, 192)
        b2, l2 = bits.copy(), ld.copy()
        _, traj = r.search(b2, l2, 0xC0FFEE, 0, 8, 512)
        out[name] = {"W": r.W, "obj_bits": r.obj_bits, "init_base": base.tolist(), "init_eval": [v, o],
                     "keys_round2": [int(k) for k in keys], "identity_key": int(last[0]),
                     "keys_round3": [int(k) for k in keys3],
                     "trajectory": [int(k) for k in traj], "final_base": r.decode(b2, l2).tolist()}
    return out


if __name__ == "__main__":
    if "--streams-only" not in sys.argv:
        old = {}
        path = os.path.join(HERE, "optima.json")
        if os.path.exists(path):
            old = json.load(open(path))
        new = optima()
        old.update(new)
        json.dump(old, open(path, "w"), indent=1)
    json.dump(streams(), open(os.path.join(HERE, "streams.json"), "w"))
    print("golden fixtures written")
