"""Iteration log of one F1 lap solve: python tools/lap_log.py NODES"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, lap_time
n = int(sys.argv[1])
p = workloads.f1_catalunya(n)
x0 = workloads.f1_start_point(p, 50.0)
nlp = CollocationNLP(p); s = LapSolver(nlp)
t0 = time.perf_counter(); r = s.solve(x0, verbose=True); dt = time.perf_counter() - t0
print("nodes %d: %.3f s, %d iterations, lap %.6f status %d %s" % (n, dt, r["iters"][0], lap_time(p, r["x"][0], r["obj"][0]), r["status"][0], r["stats"]))
