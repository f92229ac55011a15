"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck): a few rounds of the
per-round and persistent kernels, candidate keys and explicit evaluation on two small shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import kafka_assignment_optimizer_b200 as kao  # noqa: E402
from kafka_assignment_optimizer_b200 import optimizer as kopt  # noqa: E402

for args in [(64, 16, 4, 3, 1), (300, 64, 8, 3, 2)]:
    pb = kao.synthetic_problem(*args)
    s = kao.Session(pb)
    s.candidate_keys(1, 0, 256, 0, 256)
    keys, _ = s.search(1, 0, 3, 512)
    s.profile_rounds(1, 3, 2, 512)
    if s.stats()["words_per_row"] <= 2:
        dk = s.candidate_keys_delta(1, 9, 256, 0, 256)
        assert (dk == s.candidate_keys(1, 9, 256, 0, 256)).all()
        s.search_delta(1, 5, 3, 512)
    reps, v, o, mv = s.get_base()
    vv, oo = kopt.evaluate(pb, reps)
    assert (int(vv[0]), int(oo[0])) == (v, o)
    s.close()
print("sanitize driver ok")
