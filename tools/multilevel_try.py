"""Experiment: mesh continuation for long laps (coarse solve -> interpolated primal-dual warm start on the fine mesh)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, lap_time

def solve(p, x0, warm=None, mu=1e-4, **kw):
    nlp = CollocationNLP(p); s = LapSolver(nlp)
    t0 = time.perf_counter(); r = s.solve(x0, warm=warm, warm_mu=mu, **kw); dt = time.perf_counter() - t0
    s.close(); nlp.close()
    return r, dt

def interp_periodic(s_to, s_from, L, A):
    A = np.asarray(A); out = np.empty((len(s_to), A.shape[1]))
    s_ext = np.concatenate([[s_from[-1] - L], s_from, [s_from[0] + L]])
    for k in range(A.shape[1]):
        out[:, k] = np.interp(s_to, s_ext, np.concatenate([[A[-1, k]], A[:, k], [A[0, k]]]))
    return out

def prolong(pc, pf, r, ntd):
    nv, ncpe = pc.nv, pc.m // pc.n_points
    L = pc.track_length
    hc, hf = L / pc.n_points, L / pf.n_points
    x = interp_periodic(pf.s, pc.s, L, r["x"][0].reshape(-1, nv)).ravel()
    # element e ends at node e+1: rows live at the right node
    sc_rows = (pc.s + hc) % L; order = np.argsort(sc_rows)
    Y = r["y"][0].reshape(-1, ncpe)[order]
    yf = interp_periodic((pf.s + hf) % L, sc_rows[order], L, Y)
    yf[:, ntd:] *= hf / hc
    zl = interp_periodic(pf.s, pc.s, L, r["zl"][0].reshape(-1, nv)) * (hf / hc)
    zu = interp_periodic(pf.s, pc.s, L, r["zu"][0].reshape(-1, nv)) * (hf / hc)
    return x, (yf.ravel(), zl.ravel(), zu.ravel())

levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [500, 8000]
mus = [float(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1e-4]
p0 = workloads.f1_catalunya(levels[0])
x0 = workloads.f1_start_point(p0, 50.0)
solve(p0, x0)
r, dt = solve(p0, x0)
print("level %d: %.3f s, %d iterations, lap %.6f status %d" % (levels[0], dt, r["iters"][0], lap_time(p0, r["x"][0], r["obj"][0]), r["status"][0]))
for mu in mus:
    pc, rc, total = p0, r, dt
    for n in levels[1:]:
        pf = workloads.f1_catalunya(n)
        x, w = prolong(pc, pf, rc, 9)
        rf, dtf = solve(pf, x, warm=w, mu=mu)
        total += dtf
        print("  mu0 %.0e level %d: %.3f s, %d iterations, %d factorisations, lap %.6f status %d" % (mu, n, dtf, rf["iters"][0], rf["stats"]["factor_launches"], lap_time(pf, rf["x"][0], rf["obj"][0]), rf["status"][0]))
        pc, rc = pf, rf
    print("  total %.3f s" % total)
# cold from the interpolated primal only
pf = workloads.f1_catalunya(levels[-1])
x, w = prolong(p0, pf, r, 9)
rf, dtf = solve(pf, x)
print("cold from interpolated x: level %d: %.3f s, %d iterations, lap %.6f status %d" % (levels[-1], dtf, rf["iters"][0], lap_time(pf, rf["x"][0], rf["obj"][0]), rf["status"][0]))
