"""Run-time choice among the engine's full-evaluation variants.

The engine has two full evaluators with bit-identical results (row-major: every layout; column-major:
the headline layout class, docs/MODEL.md §3.3) and, for two-word rows with 769..1024 partitions,
builds the column-major one in several *schedules* of the same arithmetic (include/kao.h,
kao_set_schedule).  Which is fastest depends on the GPU and the problem, so it is measured, not
guessed: `tune()` runs a few rounds of the caller's problem through every variant in a CHILD process
(a faulting variant must not poison the caller's CUDA context), requires the round keys and the final
assignment to be identical to the row-major evaluator's, and returns the fastest such variant.
`apply()` selects it for a session and, through the KAO_SCHEDULE environment variable, for later
kao_solve calls of this process.  Tuning never changes a result.
"""
from __future__ import annotations

import json
import os
import pickle
import subprocess
import sys
import tempfile
import threading
import time

# (sync, compress, threads, unroll, roll, fuse) — the variants csrc/kao_kernels.cuh builds (KAO_FOR_TUNE_ALL)
SCHEDULES = ([(sy, c, t, u, 0, 0) for sy in (0, 1, 2, 3, 4) for (t, u) in ((768, 1), (512, 1), (512, 2)) for c in (1, 0, 2)] +
             [(sy, c, t, u, 1, 0) for sy in (1, 3) for (t, u) in ((768, 1), (512, 1), (512, 2)) for c in (1, 2)] +
             [(sy, c, t, 1, 0, 1) for sy in (0, 1, 2, 3, 4) for t in (512, 768) for c in (1, 2)])
DEFAULT_SCHEDULE = (0, 1, 768, 1, 0, 0)
SCHEDULE_FIELDS = ("sync", "compress", "threads", "unroll", "roll", "fuse")


def schedule_name(sched) -> str:
    return "column_major " + " ".join("%s=%d" % kv for kv in zip(SCHEDULE_FIELDS, sched))


def probe(pb, device: int, rounds: int, round_size: int, seed: int) -> int:
    """Child side: one warm, one recorded and two timed launches of `rounds` rounds with each variant
    on the same stream of candidates.  One `PROBE {json}` line per variant, flushed as it completes."""
    from .optimizer import Session

    ref = {}
    progress = [time.monotonic()]

    def watchdog():                                          # a variant that hangs ends the probe, not the caller
        while True:
            time.sleep(1.0)
            if time.monotonic() - progress[0] > 25.0:
                os._exit(3)

    threading.Thread(target=watchdog, daemon=True).start()

    def run(name, col, sched):
        progress[0] = time.monotonic()
        sess = Session(pb, device=device)
        try:
            if col and not sess.set_evaluator(True):
                return {"name": name, "error": "layout not covered"}
            if sched is not None and not sess.set_schedule(*sched):
                return {"name": name, "error": "schedule not built for this layout"}
            sess.search(seed, 50_000, 2, round_size)
            sess.reset()
            keys, _ = sess.search(seed, 60_000, rounds, round_size)
            base = sess.get_base()[0]
            ms = min(sess.search(seed, 70_000 + i * rounds, rounds, round_size)[1] for i in range(2))
            if not ref:
                ref["keys"], ref["base"] = keys.copy(), base.copy()
            same = bool((keys == ref["keys"]).all() and (base == ref["base"]).all())
            return {"name": name, "column_major": col, "schedule": sched, "ms_per_launch": ms, "identical_to_row_major": same}
        finally:
            sess.close()

    print("PROBE " + json.dumps(run("row_major", False, None)), flush=True)
    col = run("column_major", True, None)
    print("PROBE " + json.dumps(col), flush=True)
    if "error" not in col:
        for sched in SCHEDULES:
            if sched != DEFAULT_SCHEDULE:
                print("PROBE " + json.dumps(run(schedule_name(sched), True, sched)), flush=True)
    return 0


def tune(pb, device: int = 0, rounds: int = 32, round_size: int = 1 << 18, seed: int = 0x5EED, timeout: float = 240.0):
    """-> (use_column_major, schedule or None, report).  The fastest variant whose results are identical
    to the row-major evaluator's; whatever the child managed to report before a failure counts; any
    failure keeps the row-major evaluator."""
    rows, err = [], None
    fd, path = tempfile.mkstemp(suffix=".kao-problem")
    try:
        with os.fdopen(fd, "wb") as f:
            pickle.dump(pb, f)
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env.pop("KAO_SCHEDULE", None)
        try:
            r = subprocess.run([sys.executable, "-m", __name__, "--probe", path, "--device", str(device), "--rounds", str(rounds),
                                "--round-size", str(round_size), "--seed", str(seed)],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
            rows = [json.loads(l[6:]) for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            if r.returncode != 0:
                err = (r.stderr or r.stdout)[-300:]
        except Exception as e:                               # noqa: BLE001 — any probe failure keeps what is known to work
            err = repr(e)[:300]
            out = getattr(e, "stdout", None) or ""           # TimeoutExpired carries what the child printed so far
            if isinstance(out, bytes):
                out = out.decode(errors="replace")
            try:
                rows = [json.loads(l[6:]) for l in out.splitlines() if l.startswith("PROBE ")]
            except ValueError:
                rows = []
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass
    ok = [x for x in rows if x.get("identical_to_row_major") and "ms_per_launch" in x]
    report = {"how": "untimed probe in a child process: the same candidates through every full-evaluation variant; "
                     "identical round keys and final assignment required",
              "variants": [{k: v for k, v in x.items() if k != "column_major"} for x in rows]}
    if err:
        report["probe_error"] = err
    if not ok or not any(x["name"] == "row_major" for x in ok):
        report["selected"] = "row_major"
        return False, None, report
    best = min(ok, key=lambda x: x["ms_per_launch"])
    report["selected"] = best["name"]
    sched = tuple(best["schedule"]) if best.get("schedule") else None
    return bool(best["column_major"]), sched, report


def apply(sess, use_column_major: bool, sched) -> bool:
    """Selects the tuned variant for `sess` and (KAO_SCHEDULE) for kao_solve calls made later by this
    process.  Returns whether the column-major evaluator is in use."""
    if not use_column_major or not sess.set_evaluator(True):
        return False
    if sched is not None and sess.set_schedule(*sched):
        os.environ["KAO_SCHEDULE"] = ",".join(str(int(v)) for v in sched)
    return True


def _main(argv) -> int:
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--probe", required=True, help="pickled Problem (written by tune())")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=32)
    ap.add_argument("--round-size", type=int, default=1 << 18)
    ap.add_argument("--seed", type=int, default=0x5EED)
    a = ap.parse_args(argv)
    with open(a.probe, "rb") as f:
        pb = pickle.load(f)
    return probe(pb, a.device, a.rounds, a.round_size, a.seed)


if __name__ == "__main__":
    sys.exit(_main(sys.argv[1:]))
