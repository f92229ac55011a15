"""Sharded search rounds (docs/MODEL.md §7): one process per GPU, the index range of a round is
split contiguously over the ranks, the only exchange is one 8-byte MIN all-reduce of the packed
key (keys are < 2^63, so the signed int64 MIN of NCCL / gloo is the unsigned min-loc with
lowest-index tie-break), and every rank re-materialises the same winner.

`spread_restarts` is the other way to use several ranks (KAO_FLAG_SPREAD_RESTARTS of the C ABI, for one process
per GPU): restart r of a solve runs on rank r mod world as an ordinary single-GPU search, nothing is exchanged
until the end, where one all-gather of (violation, -objective, restart) picks the winner and its rank broadcasts
the assignment."""
from __future__ import annotations

from typing import Callable, List, Tuple

KEY_NONE = 0x7FFFFFFFFFFFFFFF


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of range(n) for `rank` of `world`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_rounds(launch: Callable[[int, int, int], None], apply: Callable[[int], None], key, first_round: int,
               rounds: int, round_size: int, rank: int, world: int, all_reduce_min=None,
               record: bool = False) -> List[int]:
    """launch(round, lo, hi) min-reduces this rank's slice into the 1-element int64 tensor `key`;
    apply(round) makes the winner in `key` the base.  `all_reduce_min(key)` is the collective
    (None for a single rank).  Returns the per-round keys if `record` (forces a sync per round)."""
    lo, hi = shard_range(round_size, rank, world)
    out = []
    for t in range(first_round, first_round + rounds):
        key.fill_(KEY_NONE)
        launch(t, lo, hi)
        if world > 1 and all_reduce_min is not None:
            all_reduce_min(key)
        apply(t)
        if record:
            out.append(int(key.item()))
    return out


def session_callbacks(sess, key, seed: int, round_size: int, stream: int = 0):
    """launch/apply closures over a kafka_assignment_optimizer_b200.Session and a device tensor."""
    ptr = key.data_ptr()
    return (lambda t, lo, hi: sess.round_launch(seed, t, round_size, lo, hi, ptr, stream),
            lambda t: sess.round_apply(seed, t, round_size, ptr, stream))


RESTART_STRIDE = 0x9E3779B97F4A7C15                  # kao_solve: seed of restart r = seed + r * this (mod 2^64)


def spread_restarts(solve_one: Callable[[int], tuple], restarts: int, seed: int, rank: int, world: int,
                    all_gather=None, broadcast=None):
    """Restarts side by side over `world` ranks.  solve_one(seed) -> (violation, objective, payload) runs ONE
    search (e.g. optimizer.solve(pb, seed=seed, device=local_rank, restarts=1, ...)); `payload` is whatever the caller
    wants back from the winner (the replicas).  all_gather(obj) -> list of every rank's obj, broadcast(obj, src) ->
    src's obj (torch.distributed.all_gather_object / broadcast_object_list wrapped by the caller; None for a
    single rank).  Returns (violation, objective, payload, restart) of the best restart — lowest violation, then
    highest objective, then lowest restart index: what one rank returns for the same restarts in sequence."""
    best = None
    for r in range(rank, restarts, world):
        v, o, payload = solve_one((seed + r * RESTART_STRIDE) & (2**64 - 1))
        cand = (int(v), -int(o), r)
        if best is None or cand < best[0]:
            best = (cand, payload)
    mine = best[0] if best is not None else None
    if world == 1 or all_gather is None:
        if best is None:
            raise ValueError("no restart ran")
        return best[0][0], -best[0][1], best[1], best[0][2]
    every = [(c, rk) for rk, c in enumerate(all_gather(mine)) if c is not None]
    if not every:
        raise ValueError("no restart ran")
    (v, neg_o, r), owner = min(every)
    payload = broadcast(best[1] if owner == rank else None, owner)
    return v, -neg_o, payload, r
