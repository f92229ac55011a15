"""Sharded search rounds (docs/MODEL.md §7): one process per GPU, the index range of a round is
split contiguously over the ranks, the only exchange is one 8-byte MIN all-reduce of the packed
key (keys are < 2^63, so the signed int64 MIN of NCCL / gloo is the unsigned min-loc with
lowest-index tie-break), and every rank re-materialises the same winner."""
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
