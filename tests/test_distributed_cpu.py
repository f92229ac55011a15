"""N>1 host-side protocol on CPU: two gloo ranks shard every round by index range, min-all-reduce
one int64 key and apply the winner; the trajectory must equal the single-process one.  The oracle
restatement stands in for the GPU kernels (tests may use it; the product may not)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kafka_assignment_optimizer_b200 import distributed as kd  # noqa: E402


def test_shard_range_partitions_the_round():
    for n in (1, 2, 7, 1024, 3001):
        for world in (1, 2, 3, 8):
            cuts = [kd.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, seed, rounds, size, q):
    from oracle import model as m, ref

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb = m.synthetic_problem(200, 64, 8, 3, remove=2)
    r = ref.Ref(pb)
    bits, ld = r.init_base()
    key = torch.zeros(1, dtype=torch.int64)

    def launch(t, lo, hi):
        k = r.candidate_keys(bits, ld, seed, t, size, lo, hi - lo, nthreads=1)
        key[0] = min(int(key[0]), int(k.min())) if k.size else int(key[0])

    def apply(t):
        nb, nl = r.gen(bits, ld, seed, t, int(key[0]) & 0xFFFFFF, size)
        bits[:], ld[:] = nb, nl

    keys = kd.run_rounds(launch, apply, key, 0, rounds, size, rank, world,
                         all_reduce_min=lambda k: dist.all_reduce(k, op=dist.ReduceOp.MIN), record=True)
    q.put((rank, keys, r.decode(bits, ld).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_walk_the_single_process_trajectory():
    from oracle import model as m, ref

    seed, rounds, size = 99, 6, 1500
    pb = m.synthetic_problem(200, 64, 8, 3, remove=2)
    r = ref.Ref(pb)
    bits, ld = r.init_base()
    _, want = r.search(bits, ld, seed, 0, rounds, size)
    want_final = r.decode(bits, ld).tolist()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, seed, rounds, size, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, keys, final in got:
        assert keys == [int(k) for k in want], rank          # identical winners on every rank
        assert final == want_final


def _restart_worker(rank, world, port, restarts, q):
    from oracle import model as m, ref

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb = m.synthetic_problem(256, 32, 4, 3, remove=2)
    r = ref.Ref(pb)

    def solve_one(seed):                                   # one short search; the restatement stands in for kao_solve
        bits, ld = r.init_base()
        last, _ = r.search(bits, ld, seed, 0, 12, 512, nthreads=1)
        v, o, _ = r.unpack_key(last)
        return v, o, r.decode(bits, ld).tolist()

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(obj, src):
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    q.put((rank,) + kd.spread_restarts(solve_one, restarts, 5, rank, world, all_gather, broadcast))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("restarts", [1, 5])
def test_restarts_side_by_side_return_what_one_rank_returns(restarts):
    """kd.spread_restarts (KAO_FLAG_SPREAD_RESTARTS for one process per GPU): two gloo ranks take the restarts in
    turn; both end with the result a single rank gets from the same restarts in sequence — also with more ranks
    than restarts."""
    from oracle import model as m, ref

    pb = m.synthetic_problem(256, 32, 4, 3, remove=2)
    r = ref.Ref(pb)

    def solve_one(seed):
        bits, ld = r.init_base()
        last, _ = r.search(bits, ld, seed, 0, 12, 512, nthreads=1)
        v, o, _ = r.unpack_key(last)
        return v, o, r.decode(bits, ld).tolist()

    want = kd.spread_restarts(solve_one, restarts, 5, 0, 1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_restart_worker, args=(rk, 2, port, restarts, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for g in got:
        assert tuple(g[1:]) == tuple(want), g[0]
