"""Accuracy of the two KKT schedules without iterative refinement (run on a GPU box)."""
import os, sys
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.helpers import load_golden, oracle_nlp, equality_rows
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver
from oracle import ipm as oipm

case = sys.argv[1] if len(sys.argv) > 1 else "f1_catalunya_discrete"
p, ex = load_golden(case)
rng = np.random.default_rng(3)
x = ex["x"] * (1 + 1e-3 * rng.standard_normal(p.n))
lam = ex["lam"]
r = oipm.oracle_evaluator(oracle_nlp(p))(x, lam, 1.0, True)
H, J = r["H"], r["J"]
eq = set(equality_rows(p).tolist())
ineq = np.array([k for k in range(p.ncpe) if k not in eq])
rows = np.arange(p.m).reshape(p.n_elements, p.ncpe)
all_ineq = rows[:, ineq].ravel()
for label, sigx, sigs, dw in (("mild", rng.uniform(0.01, 10.0, p.n), rng.uniform(0.1, 10.0, all_ineq.size), 1e-3),
                              ("wide", 10 ** rng.uniform(-8, 10, p.n), 10 ** rng.uniform(-8, 8, all_ineq.size), 1e-3),
                              ("wide, dw 1", 10 ** rng.uniform(-8, 10, p.n), 10 ** rng.uniform(-8, 8, all_ineq.size), 1.0)):
    dc = 2.0e-9
    D = np.full(p.m, dc); D[all_ineq] = 1.0 / (sigs + dw) + dc
    r1, r2 = rng.standard_normal(p.n), rng.standard_normal(p.m)
    K = sp.bmat([[H + sp.diags(sigx + dw), J.T], [J, -sp.diags(D)]], format="csc")
    rhs = np.concatenate([r1, r2])
    sol = spla.splu(K).solve(rhs)
    sol += spla.splu(K).solve(rhs - K @ sol)
    for warp in ("0", "1"):
        os.environ["FL_KKT_WARP"] = warp
        nlp = CollocationNLP(p); s = LapSolver(nlp)
        for refine in (0, 1):
            dx, dy, neg, zero = s.kkt_factor_solve(x, lam, sigx, sigs, dw, dc, r1, r2, refine_steps=refine)
            got = np.concatenate([dx[0], dy[0]])
            res = K @ got - rhs
            print("%-12s warp %s refine %d: neg %d zero %d  err %.2e  residual %.2e (rhs 1)" % (label, warp, refine, neg[0], zero[0], np.abs(got - sol).max() / np.abs(sol).max(), np.abs(res).max()))
        s.close(); nlp.close()
