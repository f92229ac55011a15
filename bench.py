#!/usr/bin/env python
"""bench.py — candidate assignments / second of the assignment-search hot path on B200.

One "step" = one pass of the hot path over one batch of candidates: ROUNDS search rounds of
ROUND_SIZE candidates each on BASELINE.json's headline topology (config 3: 1000 partitions x 64
brokers x 8 racks, RF 3, synthetic round-robin current assignment).  Every candidate is generated
on-chip from (seed, round, index) and evaluated in full (C1..C7 + objective); each round ends with
an argmin and the winner becomes the next base.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one JSON line)
  python bench.py --impl reference ...                            the CPU arm: the plain-C restatement
                                                                  of the same path on all host cores
N > 1: launched by torchrun, one rank per GPU; a round of N*ROUND_SIZE candidates is sharded by
index range, min-reduced with one 8-byte NCCL all-reduce, and applied identically on every rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, B, R, RF = 1000, 64, 8, 3                 # BASELINE.json config 3 (the metric's configuration)
ROUNDS, ROUND_SIZE = 32, 1 << 18             # per step and per GPU: 8,388,608 candidates
SEED = 0x5EED
WORKLOAD = "config3: 1000 partitions x 64 brokers x 8 racks, RF3, round-robin current assignment"
METRIC = "candidate assignments/sec at 1k-partition x 64-broker RF3"
ALGO_BYTES = P * ((B + 31) // 32) * 4 + P + 8        # 9,008 B per candidate


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def probe_evaluators(device):
    """`bench.py --probe-evaluators`: the engine's tuning probe (kafka_assignment_optimizer_b200/tuning.py)
    on the bench workload, in this process — one PROBE line per full-evaluation variant."""
    import kafka_assignment_optimizer_b200 as kao
    from kafka_assignment_optimizer_b200 import tuning

    return tuning.probe(kao.synthetic_problem(P, B, R, RF), device, ROUNDS, ROUND_SIZE, SEED)


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: not what we want here)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_rate(seconds_target=10.0):
    """The oracle restatement (plain C + OpenMP, all host cores) on a bounded sample of the same
    workload: whole rounds of the same candidate stream, every candidate evaluated in full, until
    about `seconds_target` seconds of CPU work have been spent."""
    from oracle import model, ref

    pb = model.synthetic_problem(P, B, R, RF)
    r = ref.Ref(pb)
    bits, ld = r.init_base()
    threads = host_threads()
    n = ROUND_SIZE - 1
    r.candidate_keys(bits, ld, SEED, 0, ROUND_SIZE, 0, 1 << 14, nthreads=threads)          # warm-up
    total, spent, rnd = 0, 0.0, 0
    while spent < seconds_target and rnd < 64:
        t0 = time.perf_counter()
        r.candidate_keys(bits, ld, SEED, rnd, ROUND_SIZE, 0, n, nthreads=threads)
        spent += time.perf_counter() - t0
        total += n
        rnd += 1
    return total / spent, threads, "%d candidates (%d rounds of the same Philox stream, full evaluation), %.1f s" % (
        total, rnd, spent)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import model, ref

    pb = model.synthetic_problem(P, B, R, RF)
    r = ref.Ref(pb)
    bits, ld = r.init_base()
    threads = host_threads()
    sample = 1 << 20                                        # candidates per step (bounded sample, ~2 s)
    for w in range(args.warmup):
        r.candidate_keys(bits, ld, SEED, w, ROUND_SIZE, 0, 1 << 14, nthreads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        r.candidate_keys(bits, ld, SEED, k, ROUND_SIZE, 0, min(sample, ROUND_SIZE - 1), nthreads=threads)
    dt = time.perf_counter() - t0
    n = args.steps * min(sample, ROUND_SIZE - 1)
    val = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "candidates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/u32 bitset",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference snapshot has no code and lp_solve is not installed: "
                       "this arm is the plain-C restatement of the same generate+evaluate+argmin path"},
            "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": threads, "kind": "port",
                             "sample": "%d candidates per step of the same stream" % min(sample, ROUND_SIZE - 1)},
            "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--probe-evaluators", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--evaluator", default="auto", choices=["auto", "row", "column"],
                    help="full evaluator of the search kernel: auto = probe both, keep the faster one with identical results")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: per-round min inside the kernel over NVLink peer memory (p2p) or NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.probe_evaluators:
        return probe_evaluators(args.device)

    import numpy as np
    import torch
    import torch.distributed as dist

    import kafka_assignment_optimizer_b200 as kao
    from kafka_assignment_optimizer_b200 import optimizer as kopt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    pb = kao.synthetic_problem(P, B, R, RF)
    # Both full evaluators (and the schedules of the column-major one) give bit-identical keys, so every
    # rank may choose for its own GPU: the engine's tuning probe runs the bench workload through every
    # variant in a child process, untimed, before the warm-up (kafka_assignment_optimizer_b200/tuning.py).
    from kafka_assignment_optimizer_b200 import tuning

    sched = None
    if args.evaluator == "auto":
        use_col, sched, eval_report = tuning.tune(pb, device=local, rounds=ROUNDS, round_size=ROUND_SIZE, seed=SEED)
    else:
        use_col, eval_report = args.evaluator == "column", {"selected": args.evaluator + " (forced)"}
    if world > 1 and args.collective == "nccl":
        use_col, sched, eval_report = False, None, {"selected": "row_major", "note": "the NCCL variant runs the per-round kernels (row-major evaluator)"}
    sess = kao.Session(pb, device=local)
    if use_col and not tuning.apply(sess, use_col, sched):
        use_col, sched, eval_report = False, None, dict(eval_report, selected="row_major", note="column-major refused by the session")
    gsize = ROUND_SIZE * world                               # weak scaling: per-GPU work fixed
    key = torch.full((1,), kopt.KEY_NONE, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    from kafka_assignment_optimizer_b200 import distributed as kd

    launch_cb, apply_cb = kd.session_callbacks(sess, key, SEED, gsize, stream)
    reduce_cb = (lambda k: dist.all_reduce(k, op=dist.ReduceOp.MIN)) if world > 1 else None

    if world > 1 and args.collective == "p2p":
        sess.p2p_setup_torch(dev)

    def step(k):
        """ROUNDS rounds; inputs (tables + base) are already resident in HBM."""
        if world == 1:
            sess.search(SEED, k * ROUNDS, ROUNDS, gsize)
        elif args.collective == "p2p":   # one persistent kernel per rank, keys traded over NVLink peer memory
            sess.search_sharded(SEED, k * ROUNDS, ROUNDS, gsize)
        else:   # per-round kernels + one 8-byte NCCL min all-reduce of the packed key per round
            kd.run_rounds(launch_cb, apply_cb, key, k * ROUNDS, ROUNDS, gsize, rank, world, reduce_cb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w)
    barrier()
    launches0 = sess.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                                # L2 flush between timed steps (outside the events)
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = sess.stats()["kernel_launches"] - launches0
    n_total = args.steps * ROUNDS * gsize
    value = n_total / (ms * 1e-3)
    reps, viol, obj, moves = sess.get_base()

    line = None
    if rank == 0:
        # dominant kernel = the persistent multi-round search kernel (one cooperative launch per step):
        # timed alone with CUDA events on the launching stream (kao_search brackets it)
        prof_steps = 4
        s_ms = sum(sess.search(SEED, 10_000 + i * ROUNDS, ROUNDS, ROUND_SIZE)[1] for i in range(prof_steps)) / prof_steps
        sess.profile_rounds(SEED, 19_000, 1, ROUND_SIZE)                      # load the per-round kernels (lazy module load)
        pr_ms, ap_ms = sess.profile_rounds(SEED, 20_000, 8, ROUND_SIZE)       # per-round kernels (NCCL path)
        peak, peak_src = measured_peak()
        achieved = ALGO_BYTES * ROUND_SIZE * ROUNDS / (s_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("search_persistent_kernel_column_major_dram_bytes_per_launch" if use_col
                                           else "search_persistent_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "kernel": ("search_persistent_kernel<EvalCfgT<W=2,words=32,sync=%d,compress=%d,threads=%d,unroll=%d,roll=%d,fuse=%d>> (column-major evaluator)"
                               % (sched or tuning.DEFAULT_SCHEDULE) if use_col else
                               "search_persistent_kernel<EvalCfg<W=2,NPH=3,rack=8-slot hi1,planes=3>,768>"),
                    "algorithmic_bytes_per_candidate": ALGO_BYTES, "candidates_per_launch": ROUND_SIZE * ROUNDS,
                    "kernel_ms_per_launch": s_ms,
                    "per_round_kernels_ms": {"search_round_kernel": pr_ms / 8, "apply_winner_kernel": ap_ms / 8},
                    "peak_source": peak_src,
                    "note": "candidates are generated and consumed on-chip (shared memory); measured DRAM "
                            "traffic is far below the algorithmic bytes by design"}
        # SURVEY 8(f)3, reported separately: the same search with delta evaluation (NOT the headline metric)
        dkeys, d_ms = sess.search_delta(SEED, 30_000, ROUNDS, ROUND_SIZE)
        dkeys, d_ms = sess.search_delta(SEED, 31_000, ROUNDS, ROUND_SIZE)
        delta = {"value": ROUNDS * ROUND_SIZE / (d_ms * 1e-3), "unit": "candidates/s", "kernel_ms_per_launch": d_ms,
                 "note": "same candidate stream and keys, scored from base totals + patched rows (one thread per "
                         "candidate); not a full evaluation per candidate, so not comparable with `value`"}
        # end to end through the public C-ABI call with HOST buffers (tables up, winner down, every step)
        e2e_steps = max(3, min(args.steps, 6))
        # the step's inputs live in pinned host memory; kao_solve copies them to the device every call
        import dataclasses
        pinned = {f.name: torch.from_numpy(np.ascontiguousarray(getattr(pb, f.name))).pin_memory()
                  for f in dataclasses.fields(pb) if isinstance(getattr(pb, f.name), np.ndarray)}
        pb_host = dataclasses.replace(pb, **{k: v.numpy() for k, v in pinned.items()})
        kopt.solve(pb_host, SEED, 2, 1 << 12, local, column_major=use_col)        # warm the context / module
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            res = kopt.solve(pb_host, SEED + k, ROUNDS, ROUND_SIZE, local, column_major=use_col)
        e2e_s = time.perf_counter() - t0
        h2d = (pb.rack_of.nbytes + pb.wF.nbytes + pb.wL.nbytes + 4 * 4 * pb.B + 2 * 4 * pb.R + pb.cur.nbytes)
        d2h = pb.P * pb.RF * 4 + ROUNDS * 8 + 16
        e2e = {"value": e2e_steps * ROUNDS * ROUND_SIZE / e2e_s, "unit": "candidates/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "kao_solve (pinned host buffers; create+upload+search+download+destroy per step), 1 GPU",
               "last_result": {"violation": int(res.violation), "objective": int(res.objective), "moves": int(res.moves)}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:         # reported at N=1 only
            v, threads, sample = cpu_port_rate()
            cpu = {"value": v, "unit": "candidates/s", "cores": threads, "kind": "port", "sample": sample,
                   "note": "lp_solve (the reference's solver) is not installed here and cannot be timed; "
                           "this is the plain-C/OpenMP restatement of the same path"}
        line = {"metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitset / int32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "rounds_per_step": ROUNDS, "round_size_per_gpu": ROUND_SIZE,
                           "candidates_per_step": ROUNDS * gsize, "seed": SEED,
                           "evaluator": eval_report,
                           "l2": "flushed between timed steps (256 MiB write); working set is shared-memory resident",
                           "parallelism": ("single GPU" if world == 1 else
                                           "index-range sharding over %d ranks; per-round 8-byte min %s" % (
                                               world, "inside the persistent kernel over NVLink peer mailboxes"
                                               if args.collective == "p2p" else "by NCCL all-reduce"))},
                "search_state": {"violation": int(viol), "objective": int(obj), "moves": int(moves),
                                 "exact_optimum": {"objective": 6962, "moves": 38,
                                                   "source": "tests/golden/optima.json (HiGHS)"}},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "delta_evaluation": delta, "gpu_launches": int(launches),
                "clocks": clocks}
        print(json.dumps(line))
    sess.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
