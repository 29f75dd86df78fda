import sys, time, os, numpy as np
sys.path.insert(0, "/root/repo")
from fastest_lap_b200 import gg, workloads
import fastest_lap_b200.gg as G
params = workloads._vehicles()["roberto_lot_kart_2016"]
speeds = np.linspace(50.0, 120.0, 32) / 3.6
G.gg_diagram(params, speeds[:1], n_lat=8)
ref = None
for cap, refine in ((300, 1), (30, 1), (30, 0), (300, 0)):
    G.FIRST_PASS_ITER = cap; os.environ['FL_GG_REFINE'] = str(refine)
    G.gg_diagram(params, speeds, n_lat=64)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter(); r = G.gg_diagram(params, speeds, n_lat=64); ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = r
        it = np.sort(r["ax_max_iters"].ravel()); it2 = np.sort(r["ax_min_iters"].ravel())
        print("iteration percentiles max:", [int(np.percentile(it, q)) for q in (50, 90, 95, 99, 100)], "min:", [int(np.percentile(it2, q)) for q in (50, 90, 95, 99, 100)])
    dmax = np.max(np.abs(r["ax_max"] - ref["ax_max"])); dmin = np.max(np.abs(r["ax_min"] - ref["ax_min"]))
    print("refine %d" % refine, "cap %3d: %.1f ms (%d NLPs/s), solved max %.4f min %.4f, differs from cap 300 by %.2e / %.2e" % (cap, min(ts) * 1e3, 4160 / min(ts), r["ax_max_solved"].mean(), r["ax_min_solved"].mean(), dmax, dmin))
