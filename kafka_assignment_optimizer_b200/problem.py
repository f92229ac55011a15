"""Host-side model builder: Kafka JSON -> dense integer tables of ``kao_problem``.

Mirrors the reference's L2/L3 layers: input shapes from
/root/reference/README.md:52-63 (current assignment JSON), :48 (target broker list), :27-29
(broker -> rack/AZ); the tables are the coefficients and right-hand sides of the LP at
README.md:144-185.  What the README leaves open (weights, bound formulas) is a
documented default here and can be overridden by the caller.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

LEADER_WEIGHT_BY_POS = (4, 2, 1)    # README.md:146 shows coefficients {1, 2, 4}; :131-133 leader > follower
FOLLOWER_WEIGHT_BY_POS = (2, 2, 1)


@dataclasses.dataclass
class Problem:
    P: int
    B: int
    R: int
    RF: int
    rack_of: np.ndarray     # uint8  [B]
    wF: np.ndarray          # uint16 [P, B]   README.md:145-146
    wL: np.ndarray          # uint16 [P, B]
    rep_lo: np.ndarray      # int32  [B]      C3 README.md:158-161
    rep_hi: np.ndarray
    ldr_lo: np.ndarray      # int32  [B]      C4 README.md:163-166
    ldr_hi: np.ndarray
    rack_lo: np.ndarray     # int32  [R]      C6 README.md:173-176
    rack_hi: np.ndarray
    ppr_lo: int             #                 C7 README.md:178-180
    ppr_hi: int
    cur: np.ndarray         # int32  [P, RFcur] dense indices, -1 = absent
    broker_ids: np.ndarray  # int32  [B]      dense index -> Kafka broker id
    topics: Optional[list] = None   # per-row (topic, partition)

    @classmethod
    def from_fields(cls, other) -> "Problem":
        """Copy any object exposing the same attributes (e.g. the test oracle's Problem)."""
        return cls(**{f.name: getattr(other, f.name) for f in dataclasses.fields(cls)})


def default_weights(cur: np.ndarray, P: int, B: int) -> Tuple[np.ndarray, np.ndarray]:
    """Existing placements get weight so the optimum keeps them; the preferred (first) replica
    gets more (README.md:116-120, :131-133)."""
    wF = np.zeros((P, B), np.uint16)
    wL = np.zeros((P, B), np.uint16)
    for i in range(cur.shape[1]):
        col = cur[:, i]
        rows = np.flatnonzero(col >= 0)
        wF[rows, col[rows]] = FOLLOWER_WEIGHT_BY_POS[i] if i < 3 else 1
        wL[rows, col[rows]] = LEADER_WEIGHT_BY_POS[i] if i < 3 else 1
    return wF, wL


def default_bounds(P: int, B: int, R: int, RF: int, rack_of: np.ndarray):
    """floor/ceil balance for C3/C4 (README.md:158-166 show <=2,>=1 and <=1,>=0 for 20 replicas on
    19 brokers), rack totals proportional to rack size for C6 (:173-176), floor/ceil of RF/R for
    C7 (:178-180)."""
    tot = P * RF
    size = np.bincount(rack_of, minlength=R).astype(np.int64)
    return (np.full(B, tot // B, np.int32), np.full(B, -(-tot // B), np.int32),
            np.full(B, P // B, np.int32), np.full(B, -(-P // B), np.int32),
            ((tot * size) // B).astype(np.int32), (-((-tot * size) // B)).astype(np.int32),
            RF // R, -(-RF // R))


def build_problem(current: Sequence[Sequence[int]], broker_ids: Iterable[int],
                  rack_by_broker: Dict[int, str], rf: int, topics: Optional[list] = None) -> Problem:
    ids = sorted({int(b) for b in broker_ids})          # a repeated id is one broker (as kao-cli's build_model)
    dense = {b: i for i, b in enumerate(ids)}
    racks = sorted({str(rack_by_broker[b]) for b in ids})
    ridx = {r: i for i, r in enumerate(racks)}
    rack_of = np.array([ridx[str(rack_by_broker[b])] for b in ids], np.uint8)
    P, B, R = len(current), len(ids), len(racks)
    rfc = max(1, max(len(c) for c in current))
    cur = np.full((P, rfc), -1, np.int32)
    for p, reps in enumerate(current):
        for i, b in enumerate(reps):
            cur[p, i] = dense.get(int(b), -1)
    wF, wL = default_weights(cur, P, B)
    return Problem(P, B, R, int(rf), rack_of, wF, wL, *default_bounds(P, B, R, int(rf), rack_of), cur,
                   np.array(ids, np.int32), topics)


def synthetic_problem(P: int, B0: int, R: int, RF: int, remove: int = 0, perturb: float = 0.0,
                      seed: int = 0) -> Problem:
    """Benchmark topologies: rack_of[b] = b mod R, round-robin current
    assignment, `remove` highest broker ids dropped, optional seeded re-placement of a fraction."""
    rng = np.random.RandomState(seed)
    current = [[(p + i) % B0 for i in range(RF)] for p in range(P)]
    if perturb > 0:
        for _ in range(int(round(perturb * P * RF))):
            p, i, nb = int(rng.randint(P)), int(rng.randint(RF)), int(rng.randint(B0))
            if nb not in current[p]:
                current[p][i] = nb
    return build_problem(current, range(B0 - remove), {b: "r%02d" % (b % R) for b in range(B0)}, RF)


# ---------------------------------------------------------------------------------- Kafka JSON
def parse_assignment_json(text: str):
    """`kafka-reassign-partitions --generate` "Current partition replica assignment" JSON
    (README.md:52-63).  Returns (rows of replica lists, [(topic, partition)])."""
    doc = json.loads(text) if isinstance(text, str) else text
    parts = sorted(doc["partitions"], key=lambda e: (e["topic"], int(e["partition"])))
    return [list(map(int, e["replicas"])) for e in parts], [(e["topic"], int(e["partition"])) for e in parts]


def parse_broker_list(text: str) -> List[int]:
    """`--broker-list 0,1,...,18` (README.md:48)."""
    return [int(t) for t in str(text).replace(" ", "").split(",") if t != ""]


def parse_rack_map(text: str) -> Dict[int, str]:
    """`id:rack` pairs, comma separated (the wire format is not in the reference snapshot;
    README.md:27-29 only describes the topology), e.g. "0:a,1:b,2:a"."""
    out = {}
    for tok in str(text).replace(" ", "").split(","):
        if tok:
            k, v = tok.split(":", 1)
            out[int(k)] = v
    return out


def reassignment_json(pb: Problem, replicas: np.ndarray) -> dict:
    """`--reassignment-json-file` document (README.md:67-78, :88): same shape as the input,
    leader (preferred replica) first."""
    parts = []
    for p in range(pb.P):
        topic, part = pb.topics[p] if pb.topics else ("t1", p)
        parts.append({"topic": topic, "partition": part,
                      "replicas": [int(pb.broker_ids[b]) for b in replicas[p] if b >= 0]})
    return {"version": 1, "partitions": parts}
