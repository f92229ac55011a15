"""The compiled host side (kao-cli) end to end on the GPU: README JSON in, reassignment JSON out."""
import json
import os
import subprocess

import pytest

from oracle import model as m
from test_host import README_CURRENT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kafka_assignment_optimizer_b200", "kao-cli")


def test_cli_solves_the_readme_example(tmp_path):
    f = tmp_path / "current.json"
    f.write_text(README_CURRENT)
    racks = ",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20))
    p = subprocess.run([CLI, "--assignment", str(f), "--brokers", ",".join(map(str, range(19))), "--racks", racks,
                        "--rounds", "16", "--round-size", "2048", "--stats"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    doc = json.loads(p.stdout)                                  # README.md:67-78 shape
    assert doc["version"] == 1 and len(doc["partitions"]) == 10
    rows = {e["partition"]: e["replicas"] for e in doc["partitions"]}
    want = {0: [7, 18], 2: [9, 10], 3: [0, 11], 4: [1, 12], 5: [2, 13], 6: [3, 14], 7: [4, 15], 8: [5, 16], 9: [6, 17]}
    assert all(rows[k] == v for k, v in want.items())           # "All the other moves are not required"
    assert rows[1][0] == 8 and rows[1][1] % 2 == 1 and rows[1][1] != 19   # leader kept, follower in the other AZ
    assert "objective 58 (upper bound 58: proven optimal), violation 0, replica moves 1" in p.stderr
    pb = m.readme_problem()
    reps = [[int(b) for b in rows[i]] for i in range(10)]
    assert m.evaluate(pb, __import__("numpy").array(reps)) == (0, 58)


def test_python_operator_surface_json_in_json_out():
    """AssignmentOptimizer mirrors the reference's surface: --generate JSON + broker list + rack map
    in, --reassignment-json-file JSON out (README.md:52-63 -> :67-78)."""
    import kafka_assignment_optimizer_b200 as kao

    racks = ",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20))
    opt = kao.AssignmentOptimizer(rounds=16, round_size=2048)
    doc, res = opt.optimize(README_CURRENT, ",".join(map(str, range(19))), racks)
    assert res.feasible and res.objective == 58 and res.moves == 1
    assert doc["version"] == 1 and [e["partition"] for e in doc["partitions"]] == list(range(10))
    assert doc["partitions"][1]["replicas"][0] == 8 and 19 not in doc["partitions"][1]["replicas"]
    assert all(e["topic"] == "x.y.z.t" for e in doc["partitions"])
