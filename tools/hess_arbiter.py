"""GPU Hessian vs the oracle in double and in extended precision at a stress point (N(0,1) multipliers).  GPU box."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.helpers import load_golden, oracle_nlp
from fastest_lap_b200.nlp import CollocationNLP

for case in sys.argv[1:] or ["f1_catalunya_discrete", "kart_vendrell", "f1_catalunya_3d", "f1_melbourne_derivative", "kart_catalunya_direct"]:
    p, ex = load_golden(case)
    rng = np.random.default_rng(11)
    x = ex["x"] * (1.0 + 0.01 * rng.standard_normal(p.n)) + 1.0e-3 * rng.standard_normal(p.n)
    lam, obj = rng.standard_normal(p.m), float(rng.uniform(0.1, 1.0))
    nlp = CollocationNLP(p)
    out = nlp.eval_all(x, lam, obj)
    jr, jc, hr, hc = nlp.structure()
    csr = lambda r, c, v: sp.coo_matrix((v, (r, c)), shape=(p.n, p.n)).tocsr()
    Hg = csr(hr, hc, out["hess"][0])
    Hd = csr(*oracle_nlp(p).eval(x, lam, obj)["hess"])
    Hl = csr(*oracle_nlp(p, extended=True).eval(x, lam, obj)["hess"])
    B = abs(Hl).tocoo()
    S = np.zeros(p.n // p.nv + 1); np.maximum.at(S, B.row // p.nv, B.data)
    def err(A):
        D = (A - Hl).tocoo()
        ref = np.asarray(abs(Hl)[D.row, D.col]).ravel()
        scale = np.maximum(ref, 0.1 * S[D.row // p.nv]); scale[scale == 0] = 1.0
        return D, np.abs(D.data) / scale
    Dg, eg = err(Hg); Dd, ed = err(Hd)
    print("%-28s GPU vs extended oracle: max %.2e (entries > 1e-10: %d of %d) | double oracle vs extended: max %.2e (> 1e-10: %d)" % (
        case, eg.max(), int((eg > 1e-10).sum()), Hl.nnz, ed.max(), int((ed > 1e-10).sum())))
    nlp.close()
