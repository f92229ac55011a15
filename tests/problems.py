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
    "tiny": lambda: make_problem(3, [2, 2], 2, seed=8),
}
