"""Named problem shapes shared by the CPU and GPU tests (name -> oracle Problem factory).
Together they cover every layout class of the engine: W = 1/2/4/8 words per row, slot fields of
8/16/32/64 brokers per rack, unequal racks with padding slots, removed brokers, RF raised and
lowered, C7 lower bound > 0, and dense (tie-broken) weight tables."""
from oracle import model as m

from conftest import make_problem

SHAPES = {
    "readme": lambda: m.readme_problem(),                                       # W1 S16, README.md:27-63
    "readme_tb": lambda: m.with_tiebreak(m.readme_problem()),                   # dense weights
    "cfg2": lambda: m.synthetic_problem(256, 32, 4, 3),                         # W1 S8
    "cfg2_rm2": lambda: m.synthetic_problem(256, 32, 4, 3, remove=2),           # unequal racks
    "cfg3_small": lambda: m.synthetic_problem(200, 64, 8, 3, remove=2),         # W2 S8
    "w4_s16": lambda: make_problem(150, [12, 11, 12, 10, 12, 12, 9, 12], 3, seed=1, removed=3),
    "w8_s16": lambda: make_problem(300, [16] * 16, 3, seed=2, removed=5),       # cfg5 layout
    "s32": lambda: make_problem(120, [20, 19], 2, seed=3, removed=1),           # W2, whole-word racks
    "s64_r1": lambda: make_problem(90, [40], 3, seed=4),                        # one rack of 40: C7 lo = hi = 3
    "rf_up": lambda: make_problem(100, [6, 6, 6], 4, RFcur=2, seed=5),          # RF raised, ppr [1,2]
    "rf_down": lambda: make_problem(100, [8, 8, 8, 8], 2, RFcur=4, seed=6),     # RF lowered
    "dense_small": lambda: make_problem(24, [4, 4, 4], 3, seed=7, tiebreak=True),
    # unique optimum (checked with the exact model in the test): round robin, one broker removed,
    # weights scaled + seeded random per-cell preference
    "dense_unique": lambda: m.with_random_tiebreak(m.synthetic_problem(20, 12, 4, 3, 1), 0),
    "dense_unique2": lambda: m.with_random_tiebreak(m.synthetic_problem(12, 9, 3, 2, 1), 1),
    "dense_unique3": lambda: m.with_random_tiebreak(m.synthetic_problem(12, 9, 3, 2, 1), 0),
    "dense_unique4": lambda: m.with_random_tiebreak(m.synthetic_problem(20, 12, 4, 3, 1), 2),
    "tiny": lambda: make_problem(3, [2, 2], 2, seed=8),
    # edge cases: ragged current assignment (1..4 replicas per partition, some entirely on removed
    # brokers), RF 1, a single partition, the largest row count, every one of the 256 slots in use
    "ragged": lambda: ragged_problem(),
    "rf1": lambda: make_problem(40, [3, 3, 3], 1, seed=9, removed=1),
    "one_partition": lambda: make_problem(1, [2, 2, 2], 3, seed=10),
    "max_rows": lambda: m.synthetic_problem(8160, 16, 4, 2, remove=1),
    "all_slots": lambda: m.synthetic_problem(64, 256, 16, 3),
    # two-word rows too many for mask planes + one-hot plane in shared memory: the engine falls back
    # to packed weight entries (same keys)
    "w2_rows6000": lambda: m.synthetic_problem(6000, 64, 8, 3, remove=1),
}


def ragged_problem():
    import numpy as np

    rng = np.random.RandomState(12)
    B0, removed = 14, 3
    current = []
    for p in range(60):
        k = 1 + p % 4
        current.append(list(map(int, rng.choice(B0, size=k, replace=False))))
    current[5] = [12, 13]                 # every replica on a broker that leaves the cluster
    current[6] = [11]
    racks = {b: "az%d" % (b % 3) for b in range(B0)}
    return m.build_problem(current, list(range(B0 - removed)), racks, 3)
