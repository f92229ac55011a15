import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def make_problem(P, rack_sizes, RF, RFcur=None, seed=0, removed=0, tiebreak=False):
    """Random topology in the oracle's Problem form: racks of the given sizes (interleaved broker
    ids), a random current assignment of RFcur replicas per partition, the `removed` highest
    broker ids dropped from the target list."""
    from oracle import model as m

    rng = np.random.RandomState(seed)
    RFcur = RF if RFcur is None else RFcur
    racks = []
    for r, n in enumerate(rack_sizes):
        racks += [r] * n
    order = rng.permutation(len(racks))
    rack_by_broker = {int(b): "rack%02d" % racks[order[b]] for b in range(len(racks))}
    B0 = len(racks)
    current = [list(map(int, rng.choice(B0, size=RFcur, replace=False))) for _ in range(P)]
    pb = m.build_problem(current, list(range(B0 - removed)), rack_by_broker, RF)
    if tiebreak == "random":
        pb = m.with_random_tiebreak(pb, seed)
    elif tiebreak:
        pb = m.with_tiebreak(pb)
    return pb


@pytest.fixture(scope="session")
def ref_lib():
    from oracle import ref

    ref.build()
    return ref
