"""The engine's tuning probe (kafka_assignment_optimizer_b200/tuning.py; bench.py times the variant it
picks) runs every full-evaluation variant in a child process.  The selection logic is host code:
checked here with canned child output (no GPU)."""
import json
import subprocess
import types

import kafka_assignment_optimizer_b200 as kao
from kafka_assignment_optimizer_b200 import tuning

PB = kao.synthetic_problem(16, 8, 2, 2)


class bench:                                    # the selection used to live in bench.py: same call shape
    SCHEDULES, DEFAULT_SCHEDULE = tuning.SCHEDULES, tuning.DEFAULT_SCHEDULE

    @staticmethod
    def choose_evaluator(device):
        return tuning.tune(PB, device=device, rounds=2, round_size=64)


def _fake(lines, rc=0, stderr=""):
    def run(*a, **k):
        return types.SimpleNamespace(stdout="".join("PROBE " + json.dumps(l) + "\n" for l in lines) + "noise\n", stderr=stderr, returncode=rc)
    return run


ROW = {"name": "row_major", "column_major": False, "schedule": None, "ms_per_launch": 31.8, "identical_to_row_major": True}
COL = {"name": "column_major", "column_major": True, "schedule": None, "ms_per_launch": 29.5, "identical_to_row_major": True}
FAST = {"name": "column_major sync=1 compress=0 threads=512 unroll=2 roll=0 fuse=0", "column_major": True, "schedule": [1, 0, 512, 2, 0, 0],
        "ms_per_launch": 25.0, "identical_to_row_major": True}
WRONG = {"name": "column_major sync=2 compress=1 threads=768 unroll=1 roll=0 fuse=0", "column_major": True, "schedule": [2, 1, 768, 1, 0, 0],
         "ms_per_launch": 10.0, "identical_to_row_major": False}


def test_fastest_identical_variant_wins(monkeypatch):
    monkeypatch.setattr(subprocess, "run", _fake([ROW, COL, FAST, WRONG]))
    use_col, sched, rep = bench.choose_evaluator(0)
    assert use_col and sched == (1, 0, 512, 2, 0, 0) and rep["selected"] == FAST["name"] and len(rep["variants"]) == 4


def test_default_column_major_has_no_schedule(monkeypatch):
    monkeypatch.setattr(subprocess, "run", _fake([ROW, COL, WRONG]))
    assert bench.choose_evaluator(0)[:2] == (True, None)


def test_row_major_stays_when_it_is_fastest_or_the_others_differ(monkeypatch):
    monkeypatch.setattr(subprocess, "run", _fake([dict(ROW, ms_per_launch=20.0), COL, WRONG]))
    assert bench.choose_evaluator(0)[:2] == (False, None)
    monkeypatch.setattr(subprocess, "run", _fake([ROW, dict(COL, identical_to_row_major=False), {"name": "x", "error": "schedule not built"}]))
    assert bench.choose_evaluator(0)[:2] == (False, None)


def test_partial_output_of_a_crashed_or_hung_child_counts(monkeypatch):
    monkeypatch.setattr(subprocess, "run", _fake([ROW, COL], rc=3, stderr="watchdog"))
    use_col, sched, rep = bench.choose_evaluator(0)
    assert (use_col, sched) == (True, None) and "probe_error" in rep

    def hang(*a, **k):
        raise subprocess.TimeoutExpired(cmd="probe", timeout=1, output=("PROBE " + json.dumps(ROW) + "\nPROBE " + json.dumps(FAST) + "\n").encode())
    monkeypatch.setattr(subprocess, "run", hang)
    assert bench.choose_evaluator(0)[:2] == (True, (1, 0, 512, 2, 0, 0))

    def boom(*a, **k):
        raise OSError("no python")
    monkeypatch.setattr(subprocess, "run", boom)
    assert bench.choose_evaluator(0)[:2] == (False, None)
    monkeypatch.setattr(subprocess, "run", _fake([COL, FAST]))            # no row-major reference: nothing to compare with
    assert bench.choose_evaluator(0)[:2] == (False, None)


def test_schedule_list_matches_the_engine():
    assert len(bench.SCHEDULES) == 77 and len(set(bench.SCHEDULES)) == 77 and bench.DEFAULT_SCHEDULE in bench.SCHEDULES
    assert {(t, u) for _, _, t, u, _, _ in bench.SCHEDULES} == {(768, 1), (512, 1), (512, 2)}
    # the list must be the one the engine builds (KAO_FOR_TUNE_ALL in csrc/kao_kernels.cuh)
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.abspath(tuning.__file__)), "csrc", "kao_kernels.cuh")).read()
    assert "X(S, 1, T, U, 0, 0) X(S, 0, T, U, 0, 0) X(S, 2, T, U, 0, 0)" in src and "X(S, 1, T, U, 1, 0) X(S, 2, T, U, 1, 0)" in src
    assert "X(S, 1, 512, 1, 0, 1) X(S, 2, 512, 1, 0, 1) X(S, 1, 768, 1, 0, 1) X(S, 2, 768, 1, 0, 1)" in src
    assert "KAO_FOR_TUNE_SYNC_4(X) KAO_FOR_TUNE_PLAIN(X, 4)" in src and "KAO_FOR_TUNE_SYNC_0(X) KAO_FOR_TUNE_PLAIN(X, 0)" in src
    assert re.search(r"KAO_FOR_TUNE_SYNC_1\(X\) KAO_FOR_TUNE_LOOSE\(X, 1\)", src) and re.search(r"KAO_FOR_TUNE_SYNC_3\(X\) KAO_FOR_TUNE_LOOSE\(X, 3\)", src)


def test_probe_child_without_a_gpu_fails_loudly_and_tune_keeps_the_default():
    """No monkeypatching: the real child process starts, finds no CUDA device, and tune() falls back."""
    use_col, sched, rep = tuning.tune(PB, device=0, rounds=1, round_size=16, timeout=120)
    assert (use_col, sched) == (False, None) and rep["selected"] == "row_major"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:            # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        assert "no CUDA device" in rep.get("probe_error", "")


def test_apply_sets_the_environment_for_kao_solve(monkeypatch):
    calls = []

    class FakeSession:
        def set_evaluator(self, on):
            calls.append(("eval", on))
            return True

        def set_schedule(self, *s):
            calls.append(("sched", s))
            return True

    monkeypatch.delenv("KAO_SCHEDULE", raising=False)
    import os
    assert tuning.apply(FakeSession(), True, (1, 2, 512, 1, 0, 1)) and os.environ["KAO_SCHEDULE"] == "1,2,512,1,0,1"
    assert not tuning.apply(FakeSession(), False, None)
    monkeypatch.delenv("KAO_SCHEDULE", raising=False)
