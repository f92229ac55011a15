"""The /submit façade on CPU: the HTTP layer and the request->response mapping
are exercised with a stand-in solver (the exact oracle), since solving needs a GPU."""
import json
import threading
import urllib.request

import numpy as np

from kafka_assignment_optimizer_b200 import service
from kafka_assignment_optimizer_b200.optimizer import SolveResult
from oracle import model as m
from test_host import README_CURRENT


SEEN = {}


def oracle_solver(pb, **kw):
    SEEN.update(kw)                                   # what the façade passed on
    sol = m.solve_exact(m.Problem(**{f: getattr(pb, f) for f in m.Problem.__dataclass_fields__}))
    return SolveResult(sol.replicas, sol.objective, 0, sol.moves, True, 0, 0, 0, 0.0, 0.0,
                       objective_bound=sol.objective, optimal=bool(kw.get("tight_bound")))


def test_submit_round_trip_over_http():
    srv = service.make_server(port=0, solver=oracle_solver)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        body = {"assignment": json.loads(README_CURRENT), "brokers": ",".join(map(str, range(19))),
                "racks": ",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20)), "rf": 2}
        req = urllib.request.Request("http://127.0.0.1:%d/submit" % srv.server_address[1],
                                     data=json.dumps(body).encode(), headers={"Content-Type": "application/json"})
        out = json.loads(urllib.request.urlopen(req, timeout=60).read())
        assert out["feasible"] and out["objective"] == 58 and out["moves"] == 1
        parts = out["reassignment"]["partitions"]
        assert parts[0] == {"topic": "x.y.z.t", "partition": 0, "replicas": [7, 18]}
        assert parts[1]["replicas"][0] == 8 and 19 not in parts[1]["replicas"]
        bad = urllib.request.Request("http://127.0.0.1:%d/submit" % srv.server_address[1], data=b"{}")
        try:
            urllib.request.urlopen(bad, timeout=10)
            assert False
        except urllib.error.HTTPError as e:
            assert e.code == 400
    finally:
        srv.shutdown()


def test_submit_refuses_unbounded_requests():
    """`rounds` / `round_size` / `restarts` of a request are checked before they reach the C ABI
    (a negative value would wrap to 4 billion rounds through c_uint32)."""
    import pytest

    base = {"assignment": json.loads(README_CURRENT), "brokers": ",".join(map(str, range(19))),
            "racks": ",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20)), "rf": 2}
    for bad in ({"rounds": -1}, {"rounds": 1 << 21}, {"round_size": 1}, {"round_size": 1 << 25}, {"restarts": 0},
                {"restarts": 256}, {"patience": 1 << 16}, {"rounds": 1 << 20, "round_size": 1 << 20}):
        with pytest.raises(ValueError):
            service.handle_submit(dict(base, **bad), solver=oracle_solver)
    assert service.handle_submit(dict(base, rounds=0), solver=oracle_solver)["feasible"]
    for bad in ({"gpus": 0}, {"gpus": 9}):
        with pytest.raises(ValueError):
            service.handle_submit(dict(base, **bad), solver=oracle_solver)
    out = service.handle_submit(dict(base, gpus=2, spread_restarts=True, restarts=4, certificate=True), solver=oracle_solver)
    assert (SEEN["n_gpus"], SEEN["spread_restarts"], SEEN["restarts"], SEEN["tight_bound"]) == (2, True, 4, True)
    assert out["objective_bound"] == out["objective"] == 58 and out["proven_optimal"]


def test_duplicate_broker_ids_are_one_broker():
    """a repeated id in the broker list must not create a phantom broker (kao-cli dedupes too)."""
    from kafka_assignment_optimizer_b200.problem import build_problem

    racks = {b: ("b" if b % 2 else "a") for b in range(4)}
    cur = [[0, 1], [2, 3]]
    a = build_problem(cur, [0, 1, 2, 3], racks, 2)
    b = build_problem(cur, [0, 1, 1, 2, 3, 3], racks, 2)
    assert a.B == b.B == 4 and (a.rep_hi == b.rep_hi).all() and (a.cur == b.cur).all()
