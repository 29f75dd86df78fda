import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from fastest_lap_b200 import gg, workloads
import fastest_lap_b200.gg as G
params = workloads._vehicles()["roberto_lot_kart_2016"]
speeds = np.linspace(50.0, 120.0, 32) / 3.6
G.gg_diagram(params, speeds[:1], n_lat=8)
log = []
orig = G.solve_batch
def timed(params, spec, x0, device=0, bounds=None, max_iter=300):
    from fastest_lap_b200.nlp import SteadyStateNLP
    from fastest_lap_b200.ipm import LapSolver
    t0 = time.perf_counter()
    nlp = SteadyStateNLP(G._model_of(params), params, spec, device=device)
    solver = LapSolver(nlp, options=G._options(max_iter=max_iter), bounds=bounds)
    t1 = time.perf_counter()
    out = solver.solve(x0)
    t2 = time.perf_counter()
    solver.close(); nlp.close()
    t3 = time.perf_counter()
    log.append((len(spec), t1 - t0, t2 - t1, t3 - t2, int(out["iters"].max()), out["device_ms"], out["stats"]["solver_launches"]))
    return out
G.solve_batch = timed
for rep in range(2):
    log.clear()
    t0 = time.perf_counter()
    r = G.gg_diagram(params, speeds, n_lat=64)
    tot = time.perf_counter() - t0
    print("total %.1f ms; in solve_batch %.1f ms (create %.1f, solve %.1f, destroy %.1f)" % (tot * 1e3, sum(a + b + c for _, a, b, c, _, _, _ in log) * 1e3,
          sum(l[1] for l in log) * 1e3, sum(l[2] for l in log) * 1e3, sum(l[3] for l in log) * 1e3))
    for l in log: print("  batch %5d: create %.2f ms, solve %.2f ms (device %.2f ms, %d iterations, %d launches), destroy %.2f ms" % (l[0], l[1]*1e3, l[2]*1e3, l[5], l[4], l[6], l[3]*1e3))
