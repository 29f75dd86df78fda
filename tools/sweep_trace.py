"""Per-round trace of the batched lap solve (FL_IPM_TRACE): running laps, schedule and cumulative time of every lock-step round.
    FL_IPM_PROFILE=1 python tools/sweep_trace.py LAPS OUT.csv"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
laps, out = int(sys.argv[1]), sys.argv[2]
if os.path.exists(out): os.remove(out)
os.environ["FL_IPM_TRACE"] = out
from fastest_lap_b200 import workloads
from fastest_lap_b200.ipm import default_options, solve_sweep
from fastest_lap_b200.vehicle import MODEL_F1
p = workloads.f1_catalunya(500)
params = workloads.sweep_parameters(MODEL_F1, laps); params[0] = p.params
opts = default_options(); opts.max_iter = 250
t0 = time.perf_counter()
r = solve_sweep(p, params, options=opts, device=0)
print("laps %d: %.2f s, %.1f laps/s, converged %d, phases %s, stats %s" % (laps, time.perf_counter() - t0, laps / (time.perf_counter() - t0), int((r["status"] >= 1).sum()), r["phases_s"], r["stats"]))
d = np.genfromtxt(out, delimiter=",", names=True, invalid_raise=False)
d = d[~np.isnan(d["round"])]
first = d[: int(np.argmax(np.diff(d["round"]) < 0)) + 1] if np.any(np.diff(d["round"]) < 0) else d
ms = np.diff(first["ms"]); run = first["running"][:-1]
for lo, hi in ((2049, 1 << 30), (513, 2048), (129, 512), (33, 128), (1, 32)):
    k = (run >= lo) & (run <= hi)
    if k.any(): print("rounds with %5d..%-6d running laps: %4d rounds, %8.1f ms (%.1f ms each), partitioned %d" % (lo, min(hi, laps), k.sum(), ms[k].sum(), ms[k].mean(), int(first["partitioned"][:-1][k].sum())))
