"""kao_objective_bound (csrc/kao_bound.hpp): the upper bound that lets kao_solve say "proven optimal".  Host
code, no GPU.  Checked against HiGHS: the bound is the sum of the optima of the two network-flow relaxations
(placement Y, leadership L), solved here as linear programs; it is never below the exact optimum of the 0/1
program, and it does not depend on which feasible assignment the cycle cancelling starts from."""
import dataclasses

import numpy as np
import pytest

import kafka_assignment_optimizer_b200 as kao
from oracle import model as m
from problems import SHAPES


def relaxation_optima(pb):
    """Y* + L* by linear programming (scipy / HiGHS): the same relaxations kao_bound.hpp solves by flows."""
    import scipy.sparse as sp
    from scipy.optimize import linprog

    P, B, R = pb.P, pb.B, pb.R
    Sp = sp.kron(sp.identity(P, format="csr"), np.ones((1, B)), format="csr")
    Sb = sp.kron(np.ones((1, P)), sp.identity(B, format="csr"), format="csr")
    M = sp.csr_matrix((np.ones(B), (pb.rack_of.astype(np.int64), np.arange(B))), shape=(R, B))
    Spr = sp.kron(sp.identity(P, format="csr"), M, format="csr")
    A = sp.vstack([Sp, Sb, M @ Sb, Spr])
    lo = np.concatenate([np.full(P, pb.RF), pb.rep_lo, pb.rack_lo, np.full(P * R, pb.ppr_lo)]).astype(float)
    hi = np.concatenate([np.full(P, pb.RF), pb.rep_hi, pb.rack_hi, np.full(P * R, pb.ppr_hi)]).astype(float)
    y = linprog(-pb.wF.reshape(-1).astype(float), A_ub=sp.vstack([A, -A]), b_ub=np.concatenate([hi, -lo]), bounds=(0, 1), method="highs")
    A2 = sp.vstack([Sp, Sb])
    lo2 = np.concatenate([np.ones(P), pb.ldr_lo]).astype(float)
    hi2 = np.concatenate([np.ones(P), pb.ldr_hi]).astype(float)
    bonus = (pb.wL.astype(float) - pb.wF.astype(float)).reshape(-1)
    l = linprog(-bonus, A_ub=sp.vstack([A2, -A2]), b_ub=np.concatenate([hi2, -lo2]), bounds=(0, 1), method="highs")
    assert y.status == 0 and l.status == 0
    return int(round(-y.fun)) + int(round(-l.fun))


@pytest.mark.parametrize("name", ["readme", "cfg2", "cfg2_rm2", "cfg3_small", "rf_up", "rf_down", "w4_s16", "w8_s16", "dense_small", "tiny"])
def test_flow_bound_equals_the_relaxations_and_bounds_the_optimum(name):
    pb = SHAPES[name]()
    sol = m.solve_exact(pb)
    assert sol.status == "optimal"
    kp = kao.Problem.from_fields(pb)
    cheap = kao.objective_bound(kp)
    tight = kao.objective_bound(kp, sol.replicas)
    assert sol.objective <= tight <= cheap
    assert tight == min(cheap, relaxation_optima(pb))


@pytest.mark.parametrize("name", ["cfg2_rm2", "cfg3_small", "rf_up"])
def test_flow_bound_does_not_depend_on_the_starting_assignment(name):
    """Cycle cancelling from a feasible assignment that is far from optimal (the optimum of the same constraints
    under unrelated weights) reaches the same bound as from the optimum itself."""
    pb = SHAPES[name]()
    rng = np.random.RandomState(4)
    other = dataclasses.replace(pb, wF=rng.randint(0, 3, size=pb.wF.shape).astype(np.uint16),
                                wL=rng.randint(0, 5, size=pb.wL.shape).astype(np.uint16))
    far = m.solve_exact(other).replicas
    assert m.evaluate(pb, far)[0] == 0
    kp = kao.Problem.from_fields(pb)
    best = m.solve_exact(pb)
    assert kao.objective_bound(kp, far) == kao.objective_bound(kp, best.replicas) >= best.objective > m.evaluate(pb, far)[1]


def test_readme_example_is_proven_optimal_by_the_bound():
    """README.md:83-91: objective 58 with one move; both bounds say nothing better exists."""
    pb = m.readme_problem()
    sol = m.solve_exact(pb)
    kp = kao.Problem.from_fields(pb)
    assert kao.objective_bound(kp) == kao.objective_bound(kp, sol.replicas) == sol.objective == 58


def test_an_infeasible_or_malformed_start_falls_back_to_the_cheap_bound():
    pb = SHAPES["cfg2_rm2"]()
    kp = kao.Problem.from_fields(pb)
    reps = m.solve_exact(pb).replicas.copy()
    reps[3, 1] = reps[3, 0]                                   # a duplicated broker: not an assignment
    assert kao.objective_bound(kp, reps) == kao.objective_bound(kp)
