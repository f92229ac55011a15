"""Multi-GPU parity (needs >= 2 GPUs on one node; skipped otherwise): the sharded search — with the
per-round min exchanged inside the kernel through NVLink peer mailboxes, and with the NCCL
all-reduce driver — walks exactly the single-GPU trajectory."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import kafka_assignment_optimizer_b200 as kao
    from kafka_assignment_optimizer_b200 import distributed as kd, optimizer as kopt

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pb = kao.synthetic_problem(200, 64, 8, 3, remove=2)
    seed, rounds, size = 42, 10, 3001
    # (a) in-kernel exchange over peer memory
    sess = kao.Session(pb, device=rank)
    sess.p2p_setup_torch(torch.device("cuda", rank))
    keys_a, _ = sess.search_sharded(seed, 0, rounds, size)
    keys_a2, _ = sess.search_sharded(seed, rounds, 5, size, delta=True)   # second call: other mailbox bank, delta scoring
    base_a = sess.get_base()[0]
    sess.close()
    # (b) per-round kernels + NCCL min all-reduce
    sess = kao.Session(pb, device=rank)
    key = torch.full((1,), kopt.KEY_NONE, dtype=torch.int64, device="cuda")
    launch, apply = kd.session_callbacks(sess, key, seed, size, torch.cuda.current_stream().cuda_stream)
    keys_b = kd.run_rounds(launch, apply, key, 0, rounds + 5, size, rank, world,
                           lambda k: dist.all_reduce(k, op=dist.ReduceOp.MIN), record=True)
    base_b = sess.get_base()[0]
    sess.close()
    q.put((rank, [int(k) for k in keys_a] + [int(k) for k in keys_a2], base_a.tolist(), keys_b, base_b.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_search_matches_single_gpu():
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    import kafka_assignment_optimizer_b200 as kao

    pb = kao.synthetic_problem(200, 64, 8, 3, remove=2)
    one = kao.Session(pb, device=0)
    want, _ = one.search(42, 0, 15, 3001)
    want_base = one.get_base()[0].tolist()
    one.close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, keys_a, base_a, keys_b, base_b in got:
        assert keys_a == [int(k) for k in want], "peer-mailbox path, rank %d" % rank
        assert keys_b == [int(k) for k in want], "NCCL path, rank %d" % rank
        assert base_a == want_base and base_b == want_base


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_kao_solve_gives_the_same_answer_on_any_number_of_gpus():
    """multi-GPU behind the C ABI.  kao_options.n_gpus shards every round over the GPUs of ONE
    process (a host thread per GPU, peer-memory mailboxes, no torch, no IPC); for the same global round_size
    1 / 2 / 4 / 8 GPUs return the identical assignment, key, objective and number of rounds."""
    sys.path.insert(0, ROOT)
    import kafka_assignment_optimizer_b200 as kao
    from kafka_assignment_optimizer_b200 import optimizer as kopt

    ndev = torch.cuda.device_count()
    for pb, kw in [(kao.synthetic_problem(1000, 64, 8, 3), dict(rounds=24, round_size=30011)),
                   (kao.synthetic_problem(256, 32, 4, 3, remove=2), dict(rounds=40, round_size=4099, delta=True)),
                   (kao.synthetic_problem(1000, 64, 8, 3), dict(rounds=300, round_size=1 << 13, patience=25)),
                   (kao.synthetic_problem(600, 96, 6, 3), dict(rounds=6, round_size=2048))]:     # W = 4: row-major kernels
        one = kopt.solve(pb, seed=77, n_gpus=1, **kw)
        for n in [g for g in (2, 4, 8) if g <= ndev]:
            many = kopt.solve(pb, seed=77, n_gpus=n, **kw)
            assert many.n_gpus == n and one.n_gpus == 1
            assert (many.replicas == one.replicas).all(), n
            assert (many.violation, many.objective, many.moves, many.key, many.rounds) == (
                one.violation, one.objective, one.moves, one.key, one.rounds), n
    # explicit device list, restarts
    pb = kao.synthetic_problem(256, 32, 4, 3, remove=2)
    a = kopt.solve(pb, seed=5, rounds=30, round_size=4096, restarts=3)
    b = kopt.solve(pb, seed=5, rounds=30, round_size=4096, restarts=3, n_gpus=0, device_mask=0b11)
    assert b.n_gpus == 2 and (a.replicas == b.replicas).all() and a.key == b.key
    # KAO_FLAG_SPREAD_RESTARTS: the restarts side by side, one single-GPU search per GPU at a time — same winner
    # (ties go to the lowest restart), same number of rounds in all, also with more GPUs than restarts
    for restarts in (1, 5):
        kw = dict(seed=9, rounds=120, round_size=2048, patience=40, restarts=restarts)
        a = kopt.solve(pb, **kw)
        b = kopt.solve(pb, n_gpus=2, spread_restarts=True, **kw)
        assert b.n_gpus == 2 and (a.replicas == b.replicas).all()
        assert (a.violation, a.objective, a.moves, a.key, a.rounds) == (b.violation, b.objective, b.moves, b.key, b.rounds)
    with pytest.raises(kao.KaoError):
        kopt.solve(pb, rounds=1, round_size=64, n_gpus=ndev + 1)            # more GPUs than the box has
    with pytest.raises(kao.KaoError):
        kopt.solve(pb, rounds=1, round_size=64, n_gpus=3, device_mask=0b11)  # n_gpus contradicts the mask
