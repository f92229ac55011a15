"""The engine's own device functions (csrc/kao_device.cuh: generator, full evaluator, delta
evaluator) compiled for the host under a warp emulator (tests/emu) and held to the same bit-exact
bar as the GPU parity tests: committed golden streams, the oracle restatement, the exact model.
This is a checker for the CUDA source that runs without a GPU — not a CPU path of the product
(libkao.so has none: tests/test_host.py::test_no_gpu_means_loud_failure)."""
import json
import os

import numpy as np
import pytest

import kafka_assignment_optimizer_b200 as kao
from oracle import model as m
from problems import SHAPES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# whole golden trajectories (8 rounds x 512 candidates) for one shape per evaluator family; the
# other shapes check the first rounds only (a fiber switch per lane per warp collective is slow)
FULL_TRAJECTORY = {"readme", "cfg3_small", "w8_s16", "dense_small", "rf_up", "s32"}
BIG = {"max_rows", "w2_rows6000"}


@pytest.fixture(scope="module")
def emu():
    import emu as emu_mod

    emu_mod.lib()
    return emu_mod


@pytest.fixture(scope="module")
def golden_streams():
    with open(os.path.join(GOLDEN, "streams.json")) as f:
        return json.load(f)


def product(pb):
    return kao.Problem.from_fields(pb)


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_device_code_reproduces_golden_streams(emu, golden_streams, name):
    g = golden_streams[name]
    sess = emu.EmuSession(product(SHAPES[name]()))
    base, v, o, _ = sess.get_base()
    assert base.tolist() == g["init_base"] and [v, o] == g["init_eval"]
    nkeys = 48 if name in BIG else 192
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 2, 1024, 0, nkeys)] == g["keys_round2"][:nkeys]
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 3, 1024, 0, nkeys)] == g["keys_round3"][:nkeys]      # a cycle round
    assert int(sess.candidate_keys(0xC0FFEE, 2, 1024, 1023, 1)[0]) == g["identity_key"]
    if name not in BIG:
        rounds = 8 if name in FULL_TRAJECTORY else 2
        keys = sess.search(0xC0FFEE, 0, rounds, 512)
        assert [int(k) for k in keys] == g["trajectory"][:rounds]
        if rounds == 8:
            assert sess.get_base()[0].tolist() == g["final_base"]
    sess.close()


@pytest.mark.parametrize("name", ["cfg2_rm2", "w4_s16", "s64_r1", "rf_down", "ragged", "all_slots"])
def test_device_code_vs_restatement(emu, ref_lib, name):
    """T3 on the CPU: same (seed, round, index) -> same packed key as oracle/kao_ref.c, also late in
    a round and after the base has moved."""
    pb = SHAPES[name]()
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    sess = emu.EmuSession(product(pb))
    assert sess.config()["W"] == r.W
    for rnd, size, lo, n in [(5, 4096, 4096 - 100, 100), (9, 2, 0, 2)]:
        want = r.candidate_keys(bits, ld, 0xC0FFEE, rnd, size, lo, n)
        assert (want == sess.candidate_keys(0xC0FFEE, rnd, size, lo, n)).all()
    _, want = r.search(bits, ld, 0xABCDEF12345, 3, 3, 300)
    assert (want == sess.search(0xABCDEF12345, 3, 3, 300)).all()
    reps, v, o, moves = sess.get_base()
    assert (reps == r.decode(bits, ld)).all()
    assert (v, o) == m.evaluate(pb, reps) and moves == m.replica_moves(pb, reps)
    want = r.candidate_keys(bits, ld, 7, 6, 256, 0, 64)
    assert (want == sess.candidate_keys(7, 6, 256, 0, 64)).all()
    sess.close()


@pytest.mark.parametrize("name", ["readme", "cfg2_rm2", "cfg3_small", "s32", "rf_up", "dense_small", "rf1", "ragged", "w4_s16", "w8_s16"])
def test_delta_evaluator_gives_the_full_evaluators_keys(emu, name):
    """docs/MODEL.md §8: the per-thread generator + delta evaluator return the key of the full evaluation."""
    sess = emu.EmuSession(product(SHAPES[name]()))
    for _ in range(2):
        full = sess.candidate_keys(0xD317A, 4, 512, 0, 160)
        delta = sess.candidate_keys(0xD317A, 4, 512, 0, 160, delta=True)
        assert (full == delta).all()
        assert sess.candidate_keys(0xD317A, 4, 512, 511, 1, delta=True)[0] == sess.candidate_keys(0xD317A, 4, 512, 511, 1)[0]
        sess.search(0xD317A, 0, 1, 256)             # move the base, then compare again
    sess.close()


@pytest.mark.parametrize("name", ["readme", "cfg2", "cfg3_small", "w4_s16", "w8_s16", "s64_r1", "rf_up", "ragged"])
def test_explicit_evaluation_matches_exact_model(emu, name):
    """eval_batch_kernel's evaluator on arbitrary (mostly infeasible, some malformed) assignments."""
    pb = SHAPES[name]()
    rng = np.random.RandomState(5)
    cands = []
    for i in range(12):
        reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
        if i % 3 == 1:
            reps[rng.randint(pb.P), -1] = -1
        if i % 3 == 2 and pb.RF > 1:
            p = rng.randint(pb.P)
            reps[p, 1] = reps[p, 0]
        cands.append(reps)
    sess = emu.EmuSession(product(pb))
    v, o = sess.evaluate(np.stack(cands))
    for i, reps in enumerate(cands):
        assert (int(v[i]), int(o[i])) == m.evaluate(pb, reps), i
    sess.close()


# ---- column-major full evaluator (csrc/kao_device_t.cuh) --------------------------------------
COLUMN_MAJOR = {
    "cfg2": SHAPES["cfg2"], "cfg2_rm2": SHAPES["cfg2_rm2"], "cfg3_small": SHAPES["cfg3_small"],
    "rf1": SHAPES["rf1"], "rf_down": SHAPES["rf_down"],
    "w1_3800": lambda: m.synthetic_problem(3800, 32, 4, 3, remove=1),      # 120 partition words: four per lane, rotated rows wrap
    "w2_2000": lambda: m.synthetic_problem(2000, 64, 8, 3),                # about the largest two-word shape whose planes fit
    "rf4_w2": lambda: m.synthetic_problem(300, 40, 5, 4, remove=3),
    "r8_b61": lambda: m.synthetic_problem(96, 61, 8, 3),                   # unequal racks, padding slots
    "p1100": lambda: m.synthetic_problem(1100, 64, 8, 3, remove=2),        # 40 partition words: two per lane
    "cfg3": lambda: m.synthetic_problem(1000, 64, 8, 3),                   # the headline shape, 32 words
    "max_rows": SHAPES["max_rows"],                                        # 8160 partitions: 256 words per slot, 64 chunks per column
}


@pytest.mark.parametrize("name", ["cfg2", "cfg2_rm2", "cfg3_small", "rf1", "rf_down"])
def test_column_major_evaluator_reproduces_golden_streams(emu, golden_streams, name):
    g = golden_streams[name]
    sess = emu.EmuSession(product(SHAPES[name]()))
    assert sess.set_evaluator(1)
    base, v, o, _ = sess.get_base()
    assert base.tolist() == g["init_base"] and [v, o] == g["init_eval"]
    nkeys = 64 if name in BIG else 192
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 2, 1024, 0, nkeys)] == g["keys_round2"][:nkeys]
    assert [int(k) for k in sess.candidate_keys(0xC0FFEE, 3, 1024, 0, nkeys)] == g["keys_round3"][:nkeys]      # a cycle round
    assert int(sess.candidate_keys(0xC0FFEE, 2, 1024, 1023, 1)[0]) == g["identity_key"]
    if name not in BIG:
        keys = sess.search(0xC0FFEE, 0, 8, 512)          # winners also patch the transposed planes
        assert [int(k) for k in keys] == g["trajectory"]
        assert sess.get_base()[0].tolist() == g["final_base"]
    sess.close()


@pytest.mark.parametrize("name", sorted(COLUMN_MAJOR))
def test_column_major_evaluator_on_arbitrary_bases(emu, ref_lib, name):
    """Same keys as the restatement and the exact model's evaluation on random, short-row and
    duplicate-broker bases (every row / column / rack term is exercised with non-zero violations)."""
    pb = COLUMN_MAJOR[name]()
    r = ref_lib.Ref(pb)
    sess = emu.EmuSession(product(pb))
    assert sess.set_evaluator(1)
    rng = np.random.RandomState(3)
    n = 12 if pb.P > 2000 else 40
    for it in range(4):
        reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
        if it % 3 == 1:
            for _ in range(5):
                reps[rng.randint(pb.P), -1] = -1
        if it % 3 == 2 and pb.RF > 1:
            for _ in range(5):
                p = rng.randint(pb.P)
                reps[p, 1] = reps[p, 0]
        sess.set_base(reps)
        got_reps, v, o, _ = sess.get_base()
        assert (v, o) == m.evaluate(pb, got_reps)
        bits, ld = r.encode(reps)
        want = r.candidate_keys(bits, ld, 11 + it, it, 256, 0, n)
        assert (want == sess.candidate_keys(11 + it, it, 256, 0, n)).all()
    sess.close()


def test_row_pass_takes_its_exact_path_for_broken_rows_only(emu):
    """The row pass of the column-major evaluator flags a partition (and scores it one by one from its
    row-major row) only when a rack field holds two replicas or the number of racks in use is not RF.  On a
    base whose rows are all well formed that path must never run — results would still be right if it did
    (the one-by-one terms are exact), only slow; a wrong truth table in the flag logic went unnoticed that way."""
    pb = COLUMN_MAJOR["cfg3"]()
    sess = emu.EmuSession(product(pb))
    assert sess.set_evaluator(1)
    emu.rows_charged_one_by_one()
    sess.candidate_keys(0x5EED, 0, 4096, 0, 64)
    assert emu.rows_charged_one_by_one() == 0
    reps = sess.get_base()[0].copy()
    reps[17, 2] = -1                                # a short row: n = 2
    reps[40, 1] = (reps[40, 0] + 8) % 64            # two replicas in one rack (brokers b and b + 8 share rack b mod 8)
    sess.set_base(reps)
    sess.candidate_keys(0x5EED, 0, 4096, 4095, 1)   # the identity candidate: nothing patched
    assert emu.rows_charged_one_by_one() == 2
    sess.close()


def test_column_major_forms_agree(emu):
    """1024 padded partitions: the engine picks the specialisation with 32 words per slot fixed at
    compile time; the run-time-sized form and the other schedules' arithmetic must give the same keys."""
    pb = COLUMN_MAJOR["cfg3"]()
    sess = emu.EmuSession(product(pb))
    assert sess.set_evaluator(1)
    fixed = sess.candidate_keys(0x5EED, 3, 4096, 100, 96)
    for form in (2, 3, 4, 5, 6):         # run-time word count, a POPC per word, deeper carry-save on the totals / on every stream
        assert sess.set_evaluator(form)
        assert (fixed == sess.candidate_keys(0x5EED, 3, 4096, 100, 96)).all()
    assert sess.set_evaluator(0)
    assert (fixed == sess.candidate_keys(0x5EED, 3, 4096, 100, 96)).all()
    sess.close()


@pytest.mark.parametrize("name", ["rf4_w2", "cfg2_rm2"])          # two-word and one-word rows
def test_column_major_forms_on_a_malformed_base(emu, ref_lib, name):
    pb = COLUMN_MAJOR[name]()
    r = ref_lib.Ref(pb)
    rng = np.random.RandomState(9)
    reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
    reps[7, -1] = -1
    reps[11, 1] = reps[11, 0]
    bits, ld = r.encode(reps)
    want = r.candidate_keys(bits, ld, 21, 2, 256, 0, 48)
    sess = emu.EmuSession(product(pb))
    for form in (1, 3, 4, 5, 6):
        assert sess.set_evaluator(form)
        sess.set_base(reps)
        assert (want == sess.candidate_keys(21, 2, 256, 0, 48)).all()
    sess.close()


def test_column_major_evaluator_with_wide_open_rack_bounds(emu, ref_lib):
    """C6 bounds are plain int32 of the caller: "no upper bound" as 10^6 (above 16 bits) and a lower bound no rack
    can reach must be charged exactly as the restatement charges them."""
    import dataclasses

    base = COLUMN_MAJOR["cfg2_rm2"]()
    for lo, hi in [(0, 1_000_000), (70_000, 1_000_000), (3, 5)]:
        pb = dataclasses.replace(base, rack_lo=np.full(base.R, lo, np.int32), rack_hi=np.full(base.R, hi, np.int32))
        r = ref_lib.Ref(pb)
        bits, ld = r.init_base()
        sess = emu.EmuSession(product(pb))
        assert sess.set_evaluator(1)
        assert (r.candidate_keys(bits, ld, 5, 1, 256, 0, 64) == sess.candidate_keys(5, 1, 256, 0, 64)).all(), (lo, hi)
        sess.close()


def test_column_major_evaluator_refuses_other_layouts(emu):
    # general rack bounds, whole-word racks, wide rows, dense weights, C7 lower bound, planes too large for shared memory
    for name in ["readme", "s32", "w8_s16", "dense_small", "rf_up", "w2_rows6000"]:
        sess = emu.EmuSession(product(SHAPES[name]()))
        assert not sess.set_evaluator(1)
        sess.close()


@pytest.mark.parametrize("form", [1, 2, 5, 6])
def test_column_major_long_stream_on_the_headline_shape(emu, ref_lib, form):
    """Config 3 (1000 x 64 x 8, RF 3): 12,000 candidates of four rounds against the restatement, with
    the base moved by winners in between (1-, 2- and 3-row patches in every chunk of the column walk)."""
    pb = COLUMN_MAJOR["cfg3"]()
    r = ref_lib.Ref(pb)
    bits, ld = r.init_base()
    sess = emu.EmuSession(product(pb))
    assert sess.set_evaluator(form)
    for rnd in range(4):
        want = r.candidate_keys(bits, ld, 0xFACE, rnd, 3000, 0, 3000)
        got = sess.candidate_keys(0xFACE, rnd, 3000, 0, 3000)
        bad = np.flatnonzero(want != got)
        assert bad.size == 0, (rnd, int(bad[0]), r.unpack_key(want[bad[0]]), r.unpack_key(got[bad[0]]))
        _, wk = r.search(bits, ld, 0xFACE, rnd, 1, 3000)
        assert (wk == sess.search(0xFACE, rnd, 1, 3000)).all()
    assert (sess.get_base()[0] == r.decode(bits, ld)).all()
    sess.close()


def test_column_major_vs_row_major_on_random_topologies(emu):
    """Seeded fuzz: 2..8 racks of 3..8 brokers, RF 1..4, 1..4 current replicas, 5..700 partitions, 0..2
    brokers removed.  Wherever the column-major evaluator accepts the layout, three of its forms give
    the row-major evaluator's keys, from the initial base and from a damaged one, and the base it
    reaches evaluates like the exact model."""
    from conftest import make_problem

    rng = np.random.RandomState(123)
    tested = 0
    for trial in range(40):
        R = int(rng.randint(2, 9))
        sizes = [int(rng.randint(3, 9)) for _ in range(R)]
        RF = int(rng.randint(1, min(R, 4) + 1))
        RFcur = int(rng.randint(1, 5))
        P = int(rng.choice([5, 31, 32, 33, 64, 100, 129, 257, 700]))
        removed = int(rng.randint(0, 3))
        try:
            pb = make_problem(P, sizes, RF, RFcur=RFcur, seed=trial, removed=removed)
            sess = emu.EmuSession(product(pb))
        except Exception:            # noqa: BLE001 — topologies the model builder or the engine rejects are not the subject
            continue
        if not sess.set_evaluator(1):
            sess.close()
            continue
        tested += 1
        for it in range(2):
            if it == 1:
                reps = np.stack([rng.choice(pb.B, size=pb.RF, replace=False) for _ in range(pb.P)]).astype(np.int32)
                reps[rng.randint(pb.P), -1] = -1
                sess.set_base(reps)
            keys = {}
            for form in (0, 1, 5, 6):
                assert sess.set_evaluator(form)
                keys[form] = sess.candidate_keys(77 + trial, it, 300, 0, 300)
            for form in (1, 5, 6):
                assert (keys[0] == keys[form]).all(), (trial, sizes, RF, RFcur, P, form)
            assert sess.set_evaluator(1)
            sess.search(5 + trial, 0, 2, 128)
            reps, v, o, _ = sess.get_base()
            assert (v, o) == m.evaluate(pb, reps), trial
        sess.close()
    assert tested >= 20
