"""GPU parity tests (run with -m gpu on the B200 box): every call goes through the C ABI of
libkao.so; the checker is the oracle (oracle/kao_ref.c restatement, oracle/model.py exact model)
and the committed golden fixtures.  Integer work: the bar is bit-exact."""
import json
import os

import numpy as np
import pytest

import kafka_assignment_optimizer_b200 as kao
from kafka_assignment_optimizer_b200 import optimizer as kopt
from oracle import model as m
from problems import SHAPES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def product(pb):
    return kao.Problem.from_fields(pb)


@pytest.fixture(scope="module")
def golden_streams():
    with open(os.path.join(GOLDEN, "streams.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def golden_optima():
    with open(os.path.join(GOLDEN, "optima.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_candidate_keys_bit_exact_vs_restatement(ref_lib, name):
    """T3: same (seed, round, index) -> same packed (violation, objective, index) key."""
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    sess = kao.Session(product(pb))
    assert sess.stats()["words_per_row"] == r.W
    bits, ld = r.init_base()
    base, v, o, _ = sess.get_base()
    assert (base == r.decode(bits, ld)).all()                  # same initial base (MODEL §4)
    assert (v, o) == r.evaluate(bits, ld)
    for rnd, size, lo, n in [(0, 1024, 0, 1024), (5, 4096, 4096 - 700, 700), (9, 2, 0, 2)]:
        want = r.candidate_keys(bits, ld, 0xC0FFEE, rnd, size, lo, n)
        got = sess.candidate_keys(0xC0FFEE, rnd, size, lo, n)
        bad = np.flatnonzero(want != got)
        assert bad.size == 0, "first mismatch idx %d: want %s got %s (%d bad)" % (
            lo + bad[0], sess.unpack_key(want[bad[0]]), sess.unpack_key(got[bad[0]]), bad.size)
    sess.close()


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_golden_streams(golden_streams, name):
    """Committed fixtures (tests/golden/streams.json): initial base, keys, 8-round trajectory."""
    g = golden_streams[name]
    pb = SHAPES[name]()
    sess = kao.Session(product(pb))
    base, v, o, _ = sess.get_base()
    assert base.tolist() == g["init_base"] and [v, o] == g["init_eval"]
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 2, 1024, 0, 192)] == g["keys_round2"]
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 3, 1024, 0, 192)] == g["keys_round3"]      # a cycle round
    assert int(sess.candidate_keys(0xC0FFEE, 2, 1024, 1023, 1)[0]) == g["identity_key"]
    keys, _ = sess.search(0xC0FFEE, 0, 8, 512)
    assert [int(k) for k in keys] == g["trajectory"]
    assert sess.get_base()[0].tolist() == g["final_base"]
    sess.close()


@pytest.mark.parametrize("name", ["cfg2_rm2", "cfg3_small", "w8_s16", "dense_small", "rf_up"])
def test_search_trajectory_bit_exact(ref_lib, name):
    """Whole rounds: per-round winning keys and the final assignment equal the restatement's."""
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    _, want = r.search(bits, ld, 0xABCDEF12345, 3, 12, 1500)
    sess = kao.Session(product(pb))
    got, _ = sess.search(0xABCDEF12345, 3, 12, 1500)
    assert (want == got).all()
    reps, v, o, moves = sess.get_base()
    assert (reps == r.decode(bits, ld)).all()
    assert (v, o) == m.evaluate(pb, reps) and moves == m.replica_moves(pb, reps)
    sess.close()


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_explicit_evaluation_matches_model(name):
    """kao_eval on arbitrary (mostly infeasible, some malformed) assignments == exact model."""
    pb = SHAPES[name]()
    rng = np.random.RandomState(5)
    cands = []
    for i in range(24):
        reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
        if i % 3 == 1:
            reps[rng.randint(pb.P), -1] = -1                    # short row: C1 violated
        if i % 3 == 2 and pb.RF > 1:
            p = rng.randint(pb.P)
            reps[p, 1] = reps[p, 0]                             # duplicate broker collapses (C5)
        cands.append(reps)
    v, o = kopt.evaluate(product(pb), np.stack(cands))
    for i, reps in enumerate(cands):
        assert (int(v[i]), int(o[i])) == m.evaluate(pb, reps), i


def test_readme_vector_through_the_abi():
    """README.md:83-91 through kao_solve: tie-broken weights make [8,1] the unique optimum."""
    tb = m.with_tiebreak(m.readme_problem())
    exact = m.solve_exact(tb)
    res = kopt.solve(product(tb), seed=1, rounds=40, round_size=4096)
    assert res.feasible and res.objective == exact.objective and res.moves == 1
    assert res.replicas.tolist() == exact.replicas.tolist()
    assert res.replicas.tolist()[1] == [8, 1]
    plain = m.readme_problem()
    res = kopt.solve(product(plain), seed=1, rounds=20, round_size=2048)
    assert res.feasible and res.objective == 58 and res.moves == 1
    doc = kao.problem.reassignment_json(product(plain), res.replicas)
    assert doc["partitions"][1]["replicas"][0] == 8 and doc["partitions"][0]["replicas"] == [7, 18]


@pytest.mark.parametrize("name", ["dense_unique", "dense_unique2", "dense_unique3", "dense_unique4", "readme_tb"])
def test_unique_optimum_bit_exact_winner(name):
    """T2: on instances whose optimum is unique the winner equals the exact solver's, bit for bit."""
    pb = SHAPES[name]()
    exact = m.solve_exact(pb)
    assert exact.status == "optimal" and m.is_unique_optimum(pb, exact)
    res = kopt.solve(product(pb), seed=3, rounds=200, round_size=8192)
    assert res.feasible and res.objective == exact.objective
    assert res.replicas.tolist() == exact.replicas.tolist()


@pytest.mark.parametrize("name,rounds,size", [("cfg2", 4, 4096), ("cfg2_rm2", 400, 8192)])
def test_reaches_exact_optimum_small(golden_optima, name, rounds, size):
    e = golden_optima[name]
    pb = m.synthetic_problem(*e["args"])
    res = kopt.solve(product(pb), seed=0x5EED, rounds=rounds, round_size=size)
    assert res.feasible and res.objective == e["objective"]
    assert m.evaluate(pb, res.replicas) == (0, e["objective"])
    assert res.moves == m.replica_moves(pb, res.replicas)


def test_full_size_properties_config3(golden_optima):
    """BASELINE.json config 3 (1000 x 64 x 8 racks, RF3): size-independent properties —
    feasibility, monotone descent, determinism, identity, objective bounded by the proven optimum."""
    e = golden_optima["cfg3"]
    pb = m.synthetic_problem(*e["args"])
    sess = kao.Session(product(pb))
    keys, ms = sess.search(0x5EED, 0, 150, 1 << 15)
    ks = [sess.unpack_key(k)[:2] for k in keys]
    assert all((a[0], -a[1]) >= (b[0], -b[1]) for a, b in zip(ks, ks[1:]))
    reps, v, o, moves = sess.get_base()
    assert (v, o) == ks[-1] == m.evaluate(pb, reps)
    assert v == 0 and o <= e["objective"]
    assert o == e["objective"], "did not match the exact optimum %d (got %d)" % (e["objective"], o)
    assert moves == m.replica_moves(pb, reps)
    # identity candidate of any later round reproduces the base's own evaluation
    ident = sess.candidate_keys(77, 1000, 4096, 4095, 1)[0]
    assert sess.unpack_key(ident) == (v, o, 4095)
    # determinism: same seed, fresh session -> identical trajectory
    sess2 = kao.Session(product(pb))
    keys2, _ = sess2.search(0x5EED, 0, 150, 1 << 15)
    assert (keys == keys2).all() and (sess2.get_base()[0] == reps).all()
    sess.close()
    sess2.close()


def test_sharded_round_equals_unsharded():
    """MODEL §7: a round split into index ranges and min-reduced gives the same winner, so any
    number of GPUs walks the same trajectory (here: the shards run one after another)."""
    import torch

    pb = SHAPES["cfg3_small"]()
    whole = kao.Session(product(pb))
    want, _ = whole.search(42, 0, 6, 3000)
    parts = kao.Session(product(pb))
    key = torch.empty(1, dtype=torch.int64, device="cuda")
    got = []
    for rnd in range(6):
        key.fill_(kopt.KEY_NONE)
        for lo, hi in [(0, 700), (700, 701), (701, 2048), (2048, 3000)]:
            parts.round_launch(42, rnd, 3000, lo, hi, key.data_ptr())
        parts.round_apply(42, rnd, 3000, key.data_ptr())
        torch.cuda.synchronize()
        got.append(int(key.item()) & 0xFFFFFFFFFFFFFFFF)
    assert got == [int(k) for k in want]
    assert (parts.get_base()[0] == whole.get_base()[0]).all()
    whole.close()
    parts.close()


def test_restarts_keep_the_best_assignment(golden_optima):
    """kao_options.flags: independent restarts; the result is the best of them (never worse than one)."""
    e = golden_optima["cfg2_rm2"]
    pb = product(m.synthetic_problem(*e["args"]))
    one = kopt.solve(pb, seed=11, rounds=60, round_size=2048)
    four = kopt.solve(pb, seed=11, rounds=60, round_size=2048, restarts=4)
    assert four.n_candidates == 4 * one.n_candidates
    assert (four.violation, -four.objective) <= (one.violation, -one.objective)
    assert m.evaluate(m.synthetic_problem(*e["args"]), four.replicas) == (four.violation, four.objective)
    same = kopt.solve(pb, seed=11, rounds=60, round_size=2048, restarts=4, spread_restarts=True)     # one GPU: the flag changes nothing
    assert (same.replicas == four.replicas).all() and (same.key, same.rounds) == (four.key, four.rounds)


def test_patience_stops_early_with_the_same_answer(golden_optima):
    """Early stop: the trajectory up to the stop equals the full one, and the answer on config 3 is
    still the exact optimum, found in a fraction of the rounds."""
    e = golden_optima["cfg3"]
    pb = product(m.synthetic_problem(*e["args"]))
    sess = kao.Session(pb)
    full, _ = sess.search(0x5EED, 0, 400, 1 << 14)
    sess.reset()
    sess.set_patience(60)
    part, _ = sess.search(0x5EED, 0, 400, 1 << 14)
    n = sess.last_rounds()
    assert 60 < n < 400 and (part[:n] == full[:n]).all() and (part[n:] == kopt.KEY_NONE).all()
    assert sess.unpack_key(part[n - 1])[:2] == (0, e["objective"])
    assert sess.get_base()[2] == e["objective"]
    sess.close()
    res = kopt.solve(pb, seed=0x5EED, rounds=400, round_size=1 << 14, patience=60)
    assert res.objective == e["objective"] and res.rounds == n and res.n_candidates == n << 14


def test_abi_rejects_bad_input():
    pb = product(SHAPES["tiny"]())
    with pytest.raises(kao.KaoError):
        kopt.solve(pb, rounds=1, round_size=1)                 # round_size < 2
    bad = kao.Problem.from_fields(pb)
    bad.RF = bad.B                                             # RF must be < B
    with pytest.raises(kao.KaoError):
        kao.Session(bad)


@pytest.mark.parametrize("name", ["cfg2_rm2", "s32", "w4_s16", "ragged", "rf_up"])
def test_candidate_keys_from_an_arbitrary_malformed_base(ref_lib, name):
    """kao_set_base with a damaged assignment (short rows, an empty row, duplicated brokers, rows
    violating every constraint): generator and evaluator still agree with the restatement, in full
    and in delta evaluation."""
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    rng = np.random.RandomState(21)
    reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
    reps[0, :] = -1                                            # no replica at all: leader 0xFF
    if pb.P > 3:
        reps[1, -1] = -1                                       # short row
        reps[2, -1] = reps[2, 0]                               # duplicate collapses
        reps[3, :] = reps[3, 0]
    bits, ld = r.encode(reps)
    sess = kao.Session(product(pb))
    sess.set_base(reps)
    base, v, o, _ = sess.get_base()
    assert (v, o) == r.evaluate(bits, ld) == m.evaluate(pb, r.decode(bits, ld))
    want = r.candidate_keys(bits, ld, 0xBAD, 4, 2048, 0, 2048)
    got = sess.candidate_keys(0xBAD, 4, 2048, 0, 2048)
    bad = np.flatnonzero(want != got)
    assert bad.size == 0, "idx %d: want %s got %s" % (bad[0], sess.unpack_key(want[bad[0]]), sess.unpack_key(got[bad[0]]))
    assert (sess.candidate_keys_delta(0xBAD, 4, 2048, 0, 2048) == want).all()
    _, traj = r.search(bits, ld, 0xBAD, 0, 10, 1024)
    keys, _ = sess.search(0xBAD, 0, 10, 1024)
    assert (keys == traj).all()
    sess.close()
