"""SURVEY.md 8(f)3 — delta evaluation: the key of every candidate, computed from the base's totals
and the candidate's patched rows (one thread per candidate), must equal the key of the full
evaluation bit for bit; whole searches must walk the same trajectory."""
import numpy as np
import pytest

import kafka_assignment_optimizer_b200 as kao
from oracle import model as m
from problems import SHAPES

pytestmark = pytest.mark.gpu
NARROW = sorted(SHAPES)      # every layout: rows wider than 64 slots keep their objective table in HBM (kao_plan.hpp)


@pytest.mark.parametrize("name", NARROW)
def test_delta_keys_equal_full_keys(ref_lib, name):
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    sess = kao.Session(kao.Problem.from_fields(pb))
    bits, ld = r.init_base()
    for rnd, size, lo, n in [(0, 2048, 0, 2048), (7, 4096, 4096 - 900, 900)]:
        full = sess.candidate_keys(0xD17A, rnd, size, lo, n)
        delta = sess.candidate_keys_delta(0xD17A, rnd, size, lo, n)
        want = r.candidate_keys(bits, ld, 0xD17A, rnd, size, lo, n)
        bad = np.flatnonzero(delta != full)
        assert bad.size == 0, "idx %d: delta %s full %s" % (lo + bad[0], sess.unpack_key(delta[bad[0]]),
                                                           sess.unpack_key(full[bad[0]]))
        assert (delta == want).all()
    # after some rounds the base is no longer the initial one: compare again
    sess.search(5, 0, 6, 1024)
    full = sess.candidate_keys(9, 3, 1024, 0, 1024)
    assert (sess.candidate_keys_delta(9, 3, 1024, 0, 1024) == full).all()
    sess.close()


@pytest.mark.parametrize("name", ["cfg2_rm2", "cfg3_small", "s32", "dense_unique", "rf_up"])
def test_delta_search_walks_the_same_trajectory(ref_lib, name):
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    _, want = r.search(bits, ld, 0x1234, 0, 12, 2000)
    sess = kao.Session(kao.Problem.from_fields(pb))
    got, _ = sess.search_delta(0x1234, 0, 12, 2000)
    assert (got == want).all()
    reps, v, o, _ = sess.get_base()
    assert (reps == r.decode(bits, ld)).all() and (v, o) == m.evaluate(pb, reps)
    sess.close()


def test_config4_reaches_the_exact_optimum_with_delta_search():
    """BASELINE.json config 4 (1000 x 64, 2 brokers removed): 12 short delta-scored searches (restarts) reach the
    HiGHS optimum (6787, 93 moves); the answer is re-checked by the full evaluator and the model."""
    import json
    import os

    from kafka_assignment_optimizer_b200 import optimizer as kopt

    e = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "optima.json")))["cfg4"]
    pb = m.synthetic_problem(*e["args"])
    res = kopt.solve(kao.Problem.from_fields(pb), seed=0x5EED, rounds=400, round_size=1 << 12, patience=150, restarts=12, delta=True)
    assert res.feasible and m.evaluate(pb, res.replicas) == (0, res.objective)
    assert res.objective == e["objective"] and res.moves == e["moves"]


def test_delta_search_on_wide_rows(ref_lib):
    """W = 4 / 8 words per row (config 5's layout): the delta search walks the restatement's trajectory."""
    for name in ("w4_s16", "w8_s16", "all_slots"):
        pb = SHAPES[name]()
        r = ref_lib.Ref(pb)
        bits, ld = r.init_base()
        _, want = r.search(bits, ld, 0x77, 0, 10, 1500)
        sess = kao.Session(kao.Problem.from_fields(pb))
        got, _ = sess.search_delta(0x77, 0, 10, 1500)
        assert (got == want).all() and (sess.get_base()[0] == r.decode(bits, ld)).all(), name
        sess.close()
