"""CPU tests of the oracle itself: the exact model (HiGHS) against the reference's only
known-answer vector, and the plain-C restatement against the exact model's semantics."""
import json
import os

import numpy as np
import pytest

from oracle import model as m
from problems import SHAPES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_readme_known_answer():
    # README.md:83-91: removing broker 19 only requires partition 1 to change, [8,19] -> [8,1]
    pb = m.readme_problem()
    sol = m.solve_exact(pb)
    assert sol.status == "optimal" and sol.objective == 58 and sol.moves == 1
    changed = [p for p in range(10) if set(sol.replicas[p]) != {int(b) for b in pb.cur[p] if b >= 0}]
    assert changed == [1] and sol.replicas[1][0] == 8          # leader 8 kept
    assert m.evaluate(pb, sol.replicas) == (0, 58)
    assert not m.is_unique_optimum(pb, sol)                    # 9 co-optimal followers (SURVEY.md hard part 2)


def test_readme_unique_under_tiebreak():
    tb = m.with_tiebreak(m.readme_problem())
    sol = m.solve_exact(tb)
    assert sol.replicas.tolist()[1] == [8, 1] and sol.moves == 1      # README.md:88 exactly
    assert m.is_unique_optimum(tb, sol)
    want = [[7, 18], [8, 1], [9, 10], [0, 11], [1, 12], [2, 13], [3, 14], [4, 15], [5, 16], [6, 17]]
    assert sol.replicas.tolist() == want


def test_readme_bounds_match_lp_text():
    # README.md:158-166,173-180: <=2 / >=1 replicas, <=1 / >=0 leaders per broker, <=1 per partition per rack
    pb = m.readme_problem()
    assert set(pb.rep_hi) == {2} and set(pb.rep_lo) == {1}
    assert set(pb.ldr_hi) == {1} and set(pb.ldr_lo) == {0}
    assert (pb.ppr_lo, pb.ppr_hi) == (1, 1)          # README shows "<= 1"; RF 2 over 2 AZs also forces >= 1
    assert pb.rack_lo.tolist() == [10, 9] and pb.rack_hi.tolist() == [11, 10]


def test_philox_known_answers(ref_lib):
    # Random123 kat_vectors for philox4x32-10
    R = ref_lib.Ref
    assert R.philox([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert R.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert R.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_c_eval_matches_model_semantics(ref_lib, name):
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    rng = np.random.RandomState(11)
    bits, ld = r.init_base()
    reps = r.decode(bits, ld)
    assert r.evaluate(bits, ld) == m.evaluate(pb, reps)
    b2, l2 = r.encode(reps)
    assert (b2 == bits).all() and (l2 == ld).all()
    for _ in range(20):                                        # random, mostly infeasible assignments
        reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
        b, l = r.encode(reps)
        assert r.evaluate(b, l) == m.evaluate(pb, reps)
    # walk the candidate stream: every materialised candidate still agrees with the model
    for idx in range(0, 64):
        cb, cl = r.gen(bits, ld, 7, 3, idx, 64)
        assert r.evaluate(cb, cl) == m.evaluate(pb, r.decode(cb, cl))
        assert (cb.view(np.uint8).reshape(pb.P, -1) != bits.view(np.uint8).reshape(pb.P, -1)).any(axis=1).sum() <= 3


def test_identity_candidate_and_determinism(ref_lib):
    pb = SHAPES["cfg2_rm2"]()
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    cb, cl = r.gen(bits, ld, 1, 0, 63, 64)                     # last index of a round = identity
    assert (cb == bits).all() and (cl == ld).all()
    k1 = r.candidate_keys(bits, ld, 9, 2, 512, 0, 512, nthreads=1)
    k2 = r.candidate_keys(bits, ld, 9, 2, 512, 0, 512, nthreads=4)
    assert (k1 == k2).all()
    v, o = r.evaluate(bits, ld)
    assert r.unpack_key(k1[511]) == (v, o, 511)


@pytest.mark.parametrize("name", ["readme", "cfg2_rm2", "rf_up"])
def test_c_search_reaches_exact_optimum(ref_lib, name):
    pb = SHAPES[name]()
    sol = m.solve_exact(pb)
    assert sol.status == "optimal"
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    _, keys = r.search(bits, ld, 0x5EED, 0, 300, 4096)          # ~1.2M candidates, seconds on CPU
    v, o, _ = r.unpack_key(keys[-1])
    assert v == 0 and o == sol.objective
    assert m.evaluate(pb, r.decode(bits, ld)) == (0, sol.objective)
    ks = [r.unpack_key(k)[:2] for k in keys]
    assert all((a[0], -a[1]) >= (b[0], -b[1]) for a, b in zip(ks, ks[1:]))   # monotone descent


def test_cycle_rounds_draw_closed_exchanges(ref_lib):
    """docs/MODEL.md 5: every fourth round (round mod 4 = 3) draws one family of candidates — a displaced partition
    returns home, a second partition moves onto the broker it left, a holder of the home broker moves to where the
    second one came from: three patched partitions, every broker's replica total as before (unless a target was
    already in use and REPLACE stepped on).  Free rounds draw such an exchange a few times in a hundred."""
    pb = m.synthetic_problem(512, 64, 8, 3, 0, 0.03, 5)                # tight bounds: 24 replicas on every broker
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    r.search(bits, ld, 1, 0, 200, 2048)                                # a feasible base with displaced partitions left
    assert r.evaluate(bits, ld)[0] == 0
    base_totals = np.bincount(r.decode(bits, ld).reshape(-1), minlength=pb.B)
    for rnd in (2, 3, 6, 7):
        three = closed = 0
        for idx in range(200):
            cb, cl = r.gen(bits, ld, 9, rnd, idx, 4096)
            if int(((cb != bits).any(axis=1) | (cl != ld)).sum()) == 3:
                three += 1
                closed += int((np.bincount(r.decode(cb, cl).reshape(-1), minlength=pb.B) == base_totals).all())
        if rnd % 4 == 3:
            assert three >= 180 and closed >= 150, (rnd, three, closed)
        else:
            assert three <= 100 and closed <= 40, (rnd, three, closed)


def test_golden_optima_file_is_consistent():
    """tests/golden/optima.json holds HiGHS optima of the BASELINE.json configs (generated by
    tests/golden/make_golden.py); spot-check the small ones by re-solving."""
    with open(os.path.join(GOLDEN, "optima.json")) as f:
        g = json.load(f)
    for name in ("readme", "cfg2", "cfg2_rm2"):
        e = g[name]
        pb = m.readme_problem() if name == "readme" else m.synthetic_problem(*e["args"])
        sol = m.solve_exact(pb)
        assert (sol.objective, sol.moves) == (e["objective"], e["moves"])


def test_key_layout_gives_large_problems_room_for_their_violation(ref_lib):
    """with a fixed 15-bit violation field P = 8000, RF = 4, 32 -> 64 brokers saturated every key
    (the identity won every round).  The cost field is now as wide as the problem's largest objective
    needs and the violation field takes the rest of the 63 bits."""
    pb = m.synthetic_problem(8000, 64, 8, 4)
    pb.cur[:, :] = np.stack([np.arange(4) + 4 * (p % 8) for p in range(pb.P)])      # everything sits on 32 of the 64 brokers
    pb.wF, pb.wL = m.default_weights(pb.cur, pb.P, pb.B)
    r = ref_lib.Ref(pb)
    assert r.obj_bits == (8000 * 4 * 4).bit_length() == 17
    bits, ld = r.init_base()
    v, o = r.evaluate(bits, ld)
    assert v > 0x7FFF                                            # would have saturated the old layout
    keys = r.candidate_keys(bits, ld, 1, 0, 256, 0, 256)
    assert r.unpack_key(keys[255]) == (v, o, 255)                # the identity's key holds the exact violation
    assert len({r.unpack_key(k)[0] for k in keys}) > 1           # and neighbours differ in it
    # small problems: README example, 10 x 2 x 4 = 80 -> 7 bits
    assert ref_lib.Ref(m.readme_problem()).obj_bits == 7
    assert r.pack_key(1 << 40, 5, 3) >> (24 + 17) == (1 << 22) - 1   # the field still saturates, at 63 - 24 - 17 bits
