"""Diagnosis: where a converged lap deviates from the reference's stored one (run on a GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.helpers import load_golden
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, lap_time, default_options

name = sys.argv[1] if len(sys.argv) > 1 else "f1_catalunya_discrete"
p, ex = load_golden(name)
rng = np.random.default_rng(0)
x0 = ex["x"] * (1 + 0.01 * rng.standard_normal(p.n))
for relax in (1e-8, 0.0):
    o = default_options()
    o.bound_relax_factor = relax
    nlp = CollocationNLP(p)
    s = LapSolver(nlp, options=o)
    out = s.solve(x0)
    X, R = out["x"][0].reshape(-1, p.nv), ex["x"].reshape(-1, p.nv)
    d = np.abs(X - R) / np.maximum(1.0, np.abs(R))
    print("relax %.0e: status %d iters %d laptime %.9f (stored %.6f)" % (relax, out["status"][0], out["iters"][0], lap_time(p, out["x"][0], out["obj"][0]), float(ex["laptime"])))
    print("  max deviation per variable:", np.array2string(d.max(axis=0), precision=2))
    k = np.unravel_index(np.argmax(d), d.shape)
    print("  worst at node %d variable %d: %.9g vs %.9g" % (k[0], k[1], X[k], R[k]))
    s.close(); nlp.close()
