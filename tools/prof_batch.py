"""Small driver for ncu: evaluates a batch of F1 laps (per-lap parameters) a few times.  python tools/prof_batch.py [laps] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP, EVAL_ALL
from fastest_lap_b200.problem import LapProblem
from fastest_lap_b200.vehicle import MODEL_F1

laps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shared = len(sys.argv) > 3 and sys.argv[3] == "shared"
p, ex = LapProblem.load_npz(os.path.join(ROOT, "tests", "golden", "f1_catalunya_discrete.npz"))
params = workloads.sweep_parameters(MODEL_F1, laps)
if shared:
    params = np.tile(p.params, (laps, 1))
nlp = CollocationNLP(p, vehicle_params=params)
x = np.tile(ex["x"], (laps, 1))
lam = np.tile(np.random.default_rng(1).standard_normal(p.m), (laps, 1))
nlp.eval_all(x, lam, 1.0, want=("f",))
for _ in range(reps):
    ms = nlp.time_device(EVAL_ALL, 1.0, 1)
print("laps %d: %.3f ms per round, %.3e node-evals/s" % (laps, ms, laps * p.n_points / ms * 1e3))
