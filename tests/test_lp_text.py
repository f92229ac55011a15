"""The LP-text emitter (`kao-cli --emit-lp`) writes the reference's lp_solve model
(/root/reference/README.md:139-185).  No lp_solve here, so the text is parsed back with a small
LP-format reader and solved with HiGHS: it must be the same program as oracle/model.py."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

from oracle import model as m
from test_host import README_CURRENT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kafka_assignment_optimizer_b200", "kao-cli")


def parse_lp(text):
    text = re.sub(r"//[^\n]*", "", text)
    head, binpart = text.split("\nbin\n")
    stmts = [s.strip() for s in head.split(";") if s.strip()]
    names = [v.strip() for v in binpart.replace(";", "").split(",") if v.strip()]
    idx = {v: i for i, v in enumerate(names)}

    def lin(expr):
        row = np.zeros(len(names))
        for term in expr.split("+"):
            tok = term.split()
            coef, var = (float(tok[0]), tok[1]) if len(tok) == 2 else (1.0, tok[0])
            row[idx[var]] += coef
        return row

    assert stmts[0].startswith("max:")
    c = lin(stmts[0][4:])
    rows, lo, hi = [], [], []
    for s in stmts[1:]:
        mt = re.match(r"(.*?)(<=|>=|=)\s*(-?\d+)$", s, re.S)
        rows.append(lin(mt.group(1)))
        rhs = float(mt.group(3))
        lo.append(rhs if mt.group(2) in (">=", "=") else -np.inf)
        hi.append(rhs if mt.group(2) in ("<=", "=") else np.inf)
    return names, c, sp.csr_matrix(np.array(rows)), np.array(lo), np.array(hi)


def test_emitted_lp_is_the_readme_model(tmp_path):
    import __graft_entry__ as g

    if not os.path.exists(CLI):
        g.build()
    f = tmp_path / "current.json"
    f.write_text(README_CURRENT)
    racks = ",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20))
    text = subprocess.check_output([CLI, "--assignment", str(f), "--brokers", ",".join(map(str, range(19))),
                                    "--racks", racks, "--emit-lp"], text=True)
    # the constraint families and the naming of README.md:144-185
    for needle in ("max: ", "t1b7p0_l", "// Constrain on replication factor for every partition",
                   "// Constraint on having one and only one leader per partition", "\nbin\n"):
        assert needle in text
    assert "t1b19p" not in text                               # broker 19 is not in the target list
    names, c, A, lo, hi = parse_lp(text)
    assert len(names) == 2 * 10 * 19                          # 380 binaries
    res = milp(-c, constraints=LinearConstraint(A, lo, hi), integrality=np.ones(len(names)), bounds=Bounds(0, 1))
    assert res.status == 0 and round(-res.fun) == 58          # the README optimum (oracle/model.py)
    chosen = {names[i] for i in np.flatnonzero(np.rint(res.x))}
    moved = [v for v in chosen if v.startswith("t1b") and not v.endswith("_l")
             and int(re.match(r"t1b(\d+)p(\d+)", v).group(1)) not in
             [[7, 18], [8, 19], [9, 10], [0, 11], [1, 12], [2, 13], [3, 14], [4, 15], [5, 16], [6, 17]][
                 int(re.match(r"t1b(\d+)p(\d+)", v).group(2))]]
    assert len(moved) == 1 and moved[0].endswith("p1")        # only partition 1 moves (README.md:83-91)
    assert m.solve_exact(m.readme_problem()).objective == 58
