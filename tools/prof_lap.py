"""Small driver for ncu: full callback rounds of the 8000-node lap.  python tools/prof_lap.py [nodes] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP, EVAL_ALL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
p = workloads.f1_catalunya(n)
nlp = CollocationNLP(p)
g = np.load(os.path.join(ROOT, "tests", "golden", "f1_catalunya_discrete.npz"))
x = workloads.resample_point(p, g["s"], g["x"], p.nv)
lam = np.random.default_rng(1).standard_normal(p.m)
nlp.eval_all(x, lam, 1.0, want=("f",))
for _ in range(reps):
    ms = nlp.time_device(EVAL_ALL, 1.0, 1)
print("nodes %d: %.4f ms per round" % (n, ms))
