"""Solves a batch of F1 laps (perturbed vehicles) on the GPU from the steady-state start.  python tools/ipm_batch.py [laps] [nodes] [verbose]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, lap_time
from fastest_lap_b200.vehicle import MODEL_F1

laps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 500
verbose = len(sys.argv) > 3
p = workloads.f1_catalunya(nodes)
params = workloads.sweep_parameters(MODEL_F1, laps)
params[0] = p.params
nlp = CollocationNLP(p, vehicle_params=params)
from fastest_lap_b200.ipm import default_options
opts = default_options()
opts.max_iter = int(os.environ.get('MAX_ITER', 3000))
solver = LapSolver(nlp, opts)
t0 = time.time()
if os.environ.get('START') == 'golden':          # for ncu: no steady-state solve (its KKT kernels have the same base name) in front of the laps
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'f1_catalunya_discrete.npz'))
    x_start = workloads.resample_point(p, g['s'], g['x'], p.nv) * (1.0 + 0.05 * np.random.default_rng(0).standard_normal(p.n))
else:
    x_start = workloads.f1_start_point(p, 50.0)
out = solver.solve(x_start, verbose=verbose)
dt = time.time() - t0
st, it = out["status"], out["iters"]
times = np.array([lap_time(p, out["x"][b], out["obj"][b]) for b in range(laps)])
print("laps %d nodes %d: wall %.2f s device %.1f ms -> %.1f laps/s; status counts %s; iters min/median/max %d/%d/%d; lap0 %.6f s; lap times %.3f..%.3f" % (
    laps, nodes, dt, out["device_ms"], laps / dt, {int(k): int((st == k).sum()) for k in np.unique(st)}, it.min(), np.median(it), it.max(), times[0],
    times[st == 1].min() if (st == 1).any() else -1, times[st == 1].max() if (st == 1).any() else -1))
print(out["stats"])
