"""Mesh study of the F1 Catalunya lap (BASELINE configs[0] -> configs[2]): lap time against the number of trapezoidal nodes,
and the KKT error of every returned point evaluated by the CPU oracle.  python tools/mesh_study.py [nodes ...] (GPU box)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastest_lap_b200 import workloads
from fastest_lap_b200.nlp import CollocationNLP
from fastest_lap_b200.ipm import LapSolver, lap_time
from tests.helpers import oracle_kkt


def study(nodes, refine=False, verbose=True, sequencing=False):
    """refine: every mesh after the first starts from the previous mesh's solution interpolated to it (mesh refinement);
    sequencing: ipm.solve_lap (meshes of 3000 nodes or more start from the lap on ~500 of their nodes); otherwise every mesh
    starts from the steady state at 50 km/h, the reference's starting point."""
    from fastest_lap_b200.ipm import solve_lap
    rows, prev = [], None
    for n in nodes:
        p = workloads.f1_catalunya(n)
        nlp = CollocationNLP(p)
        s = LapSolver(nlp)
        t0 = time.time()
        if sequencing:
            out, levels = solve_lap(p, workloads.f1_start_point(p, 50.0)[:p.nv])
            out["iters"][0] = sum(l["iterations"] for l in levels)
        else:
            x0 = workloads.resample_point(p, prev[0].s, prev[1], p.nv) if (refine and prev is not None) else workloads.f1_start_point(p, 50.0)
            out = s.solve(x0)
        dt = time.time() - t0
        stat, veq, vb, comp = oracle_kkt(p, out)
        lt = lap_time(p, out["x"][0], out["obj"][0])
        rows.append(dict(nodes=n, lap_time=lt, penalty=out["obj"][0] - lt, status=int(out["status"][0]), iterations=int(out["iters"][0]), seconds=dt,
                         stationarity=stat, equality_violation=veq, bound_violation=vb, complementarity=comp))
        if verbose:
            print("nodes %(nodes)5d: lap time %(lap_time).6f s  penalty %(penalty).2e  status %(status)d  iterations %(iterations)3d  %(seconds).2f s | oracle: stationarity "
                  "%(stationarity).1e  equalities %(equality_violation).1e  bounds %(bound_violation).1e  complementarity %(complementarity).1e" % rows[-1], flush=True)
        prev = (p, out["x"][0].copy())
        s.close(); nlp.close()
    return rows


if __name__ == "__main__":
    nodes = [int(a) for a in sys.argv[1:]] or [500, 1000, 2000, 4000, 8000]
    for mode in ("steady", "refine", "sequencing"):
        print({"steady": "---- every mesh from the steady-state start", "refine": "---- every mesh from the previous mesh's solution",
               "sequencing": "---- ipm.solve_lap: meshes of 3000 nodes or more from the lap on ~500 of their nodes"}[mode])
        rows = study(nodes, mode == "refine", sequencing=mode == "sequencing")
        lts = [r["lap_time"] for r in rows]
        print("differences between consecutive meshes [ms]:", ["%.2f" % (1e3 * (b - a)) for a, b in zip(lts[:-1], lts[1:])])
