"""CPU statements of the device solver (oracle/kkt.py, oracle/ipm.py): the block-cyclic-tridiagonal factorisation against
scipy's sparse LU, and the interior-point iteration pinned on a converged lap of the reference (open chicane section,
f1_optimal_laptime_catalunya_chicane*.xml via tests/golden)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from tests.helpers import load_golden, oracle_nlp, equality_rows
from oracle import kkt, ipm


def test_pivoted_gauss_jordan_inverse_and_inertia():
    rng = np.random.default_rng(0)
    for trial in range(40):
        if trial % 2 == 0:
            A = rng.standard_normal((28, 28)); A = A + A.T
        else:
            H = rng.standard_normal((15, 15)); H = H + H.T
            J = rng.standard_normal((13, 15)) * 10 ** rng.uniform(-2, 3)
            A = np.block([[H, J.T], [J, -1e-9 * np.eye(13)]])
        Ai, neg, zero = kkt.pivoted_gauss_jordan(A)
        w = np.linalg.eigvalsh(A)
        assert neg == np.sum(w < 0) and zero == 0
        assert np.max(np.abs(Ai @ A - np.eye(28))) <= 1e-12 * np.linalg.cond(A) + 1e-9


@pytest.mark.parametrize("case", ["f1_catalunya_chicane", "kart_vendrell"])
def test_block_factorisation_matches_sparse_lu(case):
    p, ex = load_golden(case)
    rng = np.random.default_rng(0)
    x = ex["x"] * (1 + 1e-3 * rng.standard_normal(p.n))
    r = ipm.oracle_evaluator(oracle_nlp(p))(x, ex["lam"], 1.0, True)
    H, J = r["H"], r["J"]
    lay = kkt.StageLayout(p.n_points, p.n_elements, p.nv, p.ncpe, equality_rows(p), p.closed)
    sig_x = rng.uniform(0.01, 10.0, p.n)
    D = np.full(p.m, 1e-9)
    D[lay.all_ineq] = rng.uniform(0.1, 10.0, len(lay.all_ineq))
    r1, r2 = rng.standard_normal(p.n), rng.standard_normal(p.m)
    Dg, Lo = kkt.assemble_blocks(lay, H, J, sig_x, D)
    F = kkt.BlockCyclicLDL(Dg, Lo, p.closed)
    z = F.solve(kkt.condense_rhs(lay, J, D, r1, r2))
    dx, dy = kkt.expand_solution(lay, J, D, r2, z)
    assert F.inertia == (p.n, lay.n_stages * lay.neq, 0)
    K = sp.bmat([[H + sp.diags(sig_x), J.T], [J, -sp.diags(D)]], format="csc")
    sol = spla.splu(K).solve(np.concatenate([r1, r2]))
    assert np.max(np.abs(np.concatenate([dx, dy]) - sol)) <= 1e-8 * np.max(np.abs(sol))


def test_interior_point_returns_to_the_reference_solution():
    """Open chicane section (144 nodes): from the golden point perturbed by 1 % the CPU iteration converges back to it."""
    from fastest_lap_b200.nlp import problem_bounds
    p, ex = load_golden("f1_catalunya_chicane")
    xl, xu, gl, gu = problem_bounds(p)
    o = oracle_nlp(p)
    rng = np.random.default_rng(0)
    x0 = ex["x"] * (1 + 0.01 * rng.standard_normal(p.n))
    lay = kkt.StageLayout(p.n_points, p.n_elements, p.nv, p.ncpe, equality_rows(p), p.closed)
    res = ipm.solve(ipm.oracle_evaluator(o), x0, xl, xu, gl, gu, layout=lay)
    assert res.status == "success"
    assert np.max(np.abs(res.x - ex["x"])) <= 1e-5
    f_gold = o.eval(ex["x"], derivs=False)["f"]
    assert abs(res.f - f_gold) <= 1e-8 * abs(f_gold)


@pytest.mark.parametrize("case,n_sep", [("f1_catalunya_chicane", 8), ("kart_vendrell", 22)])
def test_partitioned_factorisation_equals_the_sequential_one(case, n_sep):
    """The chain cut at separators (oracle/kkt.py::PartitionedLDL, the statement of csrc/kkt_part_kernels.cuh): same solution,
    same inertia -- also when the Hessian is shifted to be indefinite."""
    p, ex = load_golden(case)
    rng = np.random.default_rng(0)
    x = ex["x"] * (1 + 1e-3 * rng.standard_normal(p.n))
    r = ipm.oracle_evaluator(oracle_nlp(p))(x, ex["lam"], 1.0, True)
    H, J = r["H"], r["J"]
    lay = kkt.StageLayout(p.n_points, p.n_elements, p.nv, p.ncpe, equality_rows(p), p.closed)
    sig_x = rng.uniform(0.01, 10.0, p.n)
    D = np.full(p.m, 1e-9)
    D[lay.all_ineq] = rng.uniform(0.1, 10.0, len(lay.all_ineq))
    b = kkt.condense_rhs(lay, J, D, rng.standard_normal(p.n), rng.standard_normal(p.m))
    for shift in (0.0, -50.0):
        Dg, Lo = kkt.assemble_blocks(lay, H, J, sig_x + shift, D)
        F0 = kkt.BlockCyclicLDL(Dg.copy(), Lo.copy(), p.closed)
        Fp = kkt.PartitionedLDL(Dg.copy(), Lo.copy(), p.closed, n_sep)
        assert Fp.inertia == F0.inertia
        if shift == 0.0:
            assert F0.inertia == (p.n, lay.n_stages * lay.neq, 0)
            z0, zp = F0.solve(b), Fp.solve(b)
            assert np.max(np.abs(z0 - zp)) <= 1e-9 * np.max(np.abs(z0))
