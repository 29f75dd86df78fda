"""A/B of the two KKT schedules inside the interior-point iteration (run on a GPU box): python tools/kkt_ab.py [case] [laps] [max_iter]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.helpers import load_golden
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, default_options

case = sys.argv[1] if len(sys.argv) > 1 else "f1_ovaltrack_closed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 60
p, ex = load_golden(case)
rng = np.random.default_rng(0)
x0 = ex["x"][None, :] * (1 + 0.01 * rng.standard_normal((B, p.n)))
params = np.tile(p.params, (B, 1))
for warp in ("0", "1"):
    os.environ["FL_KKT_WARP"] = warp
    nlp = CollocationNLP(p, vehicle_params=params)
    o = default_options(); o.max_iter = max_iter
    s = LapSolver(nlp, options=o)
    s.set_partitioned(-1)
    out = s.solve(x0, verbose=len(sys.argv) > 4)
    print("FL_KKT_WARP=%s: status %s iters %s obj %s stats %s" % (warp, out["status"], out["iters"], np.array2string(out["obj"], precision=9), out["stats"]))
    s.close(); nlp.close()
