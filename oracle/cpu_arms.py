"""TEST INFRASTRUCTURE ONLY -- timed CPU arms of bench.py for the two solver metrics (laps/s, NLPs/s).

The reference solves a lap with Ipopt + MUMPS over CppAD callbacks on one thread (optimal_laptime.hpp:528-556) and a G-G
diagram as a sequence of small Ipopt solves, each warm-started from the previous one (steady_state.hpp:594-880).  None of
that stack is in this image, so the arms below time the CPU statement of the same path: the oracle port's callbacks
(oracle/*.hpp) under the interior-point iteration of oracle/ipm.py with a sparse direct solver (scipy's SuperLU; dense
numpy for the 8-variable steady-state NLPs).  P independent processes with one thread each, as the reference would be
run on P cores.  Only bench.py's cpu legs import this module.
"""
import os
import time
import numpy as np


def _one_thread():
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"


def f1_steady_start(params, v):
    """Node vector of the straight-line steady state at speed v (the reference's starting point, fastestlapc.cpp:1508-1545)."""
    from scipy.optimize import fsolve
    from . import oracle as orc
    x0 = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -0.25, -0.25, -0.25, -0.25, 0.0, 0.0])        # limebeer2014f1.h:57-60
    x = fsolve(lambda z: orc.f1_steady(np.concatenate([z, [0.0, 0.0]]), params, v)["c"][:11], x0, xtol=1e-13)
    ang = 10.0 * np.pi / 180.0
    return np.array([x[0], x[1], x[2], x[3], v * np.cos(x[4] * ang), -v * np.sin(x[4] * ang), 0.0, x[5], x[6], x[7], x[8], 0.0, x[4] * ang, x[9] * ang, x[10]])


def _lap_worker(args):
    """One F1 lap (Catalunya, `n_nodes` nodes, vehicle `params`) from the steady state at 50 km/h to tol 1e-10."""
    _one_thread()
    import sys
    root, n_nodes, params, max_iter = args
    sys.path.insert(0, root)
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import problem_bounds
    from tests.helpers import oracle_nlp
    from oracle import ipm
    p = workloads.f1_catalunya(n_nodes)
    p.params = np.asarray(params, dtype=np.float64).copy()
    xl, xu, gl, gu = problem_bounds(p)
    o = oracle_nlp(p)
    x0 = np.tile(f1_steady_start(p.params, 50.0 / 3.6), p.n_points)
    opts = ipm.Options()
    opts.max_iter = max_iter
    t0 = time.perf_counter()
    r = ipm.solve(ipm.oracle_evaluator(o), x0, xl, xu, gl, gu, opts=opts, layout=None)
    return dict(seconds=time.perf_counter() - t0, status=r.status, iters=int(r.iters), obj=float(r.f))


def laps_per_s(root, processes, n_nodes=500, max_iter=250):
    """`processes` laps of the parameter sweep (workloads.sweep_parameters, lap 0 = the reference vehicle) solved at the same
    time, one per process and core: laps/s = laps solved / wall time of the slowest."""
    import multiprocessing as mp
    import sys
    sys.path.insert(0, root)
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.vehicle import MODEL_F1
    params = workloads.sweep_parameters(MODEL_F1, max(processes, 1))
    params[0] = workloads.f1_catalunya(16).params
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        res = pool.map(_lap_worker, [(root, n_nodes, params[k], max_iter) for k in range(processes)])
    wall = max(r["seconds"] for r in res)          # the slowest worker's own solve time (process start-up excluded)
    ok = [r for r in res if r["status"] in ("success", "acceptable")]
    return dict(value=len(ok) / wall, unit="laps/s", cores=processes, kind="port", seconds=wall, laps=processes, converged=len(ok),
                iterations=[r["iters"] for r in res], seconds_per_lap=[round(r["seconds"], 2) for r in res], lap0_objective=res[0]["obj"],
                sample="%d laps of the sweep (one per process, 1 thread each): oracle callbacks + interior-point iteration of oracle/ipm.py + "
                       "scipy SuperLU, %d nodes, tol 1e-10" % (processes, n_nodes))


def _kart_evaluator(params, spec):
    import scipy.sparse as sp
    from . import oracle as orc
    v, ax, ay, fax, fay, cax, cay = spec

    def evaluate(x, y, obj, derivs):
        r = orc.kart_steady(x, params, v, ax, ay, bool(fax), bool(fay), cax, cay, lam=y, obj_factor=obj)
        out = {"f": r["f"], "g": r["c"]}
        if derivs:
            out["grad"], out["J"], out["H"] = r["grad"], sp.csr_matrix(r["jac"]), sp.csr_matrix(r["hess"])
        return out
    return evaluate


def _gg_worker(args):
    """The reference's order at one speed (steady_state.hpp:594-880): 0 g, maximum lateral acceleration, then the fan of
    lateral accelerations, maximum longitudinal acceleration at each, every solve started from the previous solution."""
    _one_thread()
    import sys
    root, params, v, n_lat = args
    sys.path.insert(0, root)
    from oracle import ipm
    # lot2016kart.h:81-121 (variables) and :194-198 (kappa +-0.11, lambda +-0.09): the bounds the device solver gets (fl_problem_bounds)
    deg = np.pi / 180.0
    xl = np.array([-0.5, 1.0e-4, -10.0 * deg, -10.0 * deg, -10.0 * deg, -10.0 * deg, -20.0, -20.0])
    xu = np.array([0.5, 0.04, 10.0 * deg, 10.0 * deg, 10.0 * deg, 10.0 * deg, 20.0, 20.0])
    gu = np.array([0.0] * 6 + [0.11, 0.11, 0.09, 0.09, 0.09, 0.09])
    gl = -gu
    opts = ipm.Options()
    opts.tol, opts.constr_viol_tol, opts.acceptable_tol, opts.max_iter = 1.0e-8, 1.0e-8, 1.0e-6, 300
    x = np.array([0.0, 0.02, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    solved = total = 0
    t0 = time.perf_counter()

    def run(spec, start):
        nonlocal solved, total
        r = ipm.solve(_kart_evaluator(params, spec), start, xl, xu, gl, gu, opts=opts, layout=None)
        total += 1
        ok = r.status in ("success", "acceptable")
        solved += int(ok)
        return r.x if ok else start

    x = run((v, 0, 0, 0, 0, 0, 0), x)
    x_lat = run((v, 0, 0, 1, 1, 0, -1.0), x)
    ay_max = x_lat[7]
    xc = x.copy()
    for k in range(n_lat - 1):
        xc = run((v, 0, ay_max * k / (n_lat - 1), 1, 0, -1.0, 0), xc)
    return dict(seconds=time.perf_counter() - t0, solved=solved, nlps=total, ay_max=float(ay_max))


def nlps_per_s(root, processes, n_lat=16):
    """`processes` speeds of the kart G-G sweep, one per process: 0 g + maximum lateral + `n_lat` - 1 maximum-longitudinal NLPs."""
    import multiprocessing as mp
    import sys
    sys.path.insert(0, root)
    from fastest_lap_b200 import workloads
    params = workloads._vehicles()["roberto_lot_kart_2016"]
    speeds = np.linspace(50.0, 120.0, max(processes, 2))[:processes] / 3.6
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        res = pool.map(_gg_worker, [(root, params, float(v), n_lat) for v in speeds])
    wall = max(r["seconds"] for r in res)          # the slowest worker's own solve time (process start-up excluded)
    n = sum(r["nlps"] for r in res)
    return dict(value=n / wall, unit="NLPs/s", cores=processes, kind="port", seconds=wall, nlps=n, solved=sum(r["solved"] for r in res),
                sample="%d speeds (one per process, 1 thread each) x (0 g, maximum lateral, %d maximum-longitudinal NLPs in the reference's "
                       "continuation order): oracle steady-state equations (second-order jets) + interior-point iteration of oracle/ipm.py, dense" % (processes, n_lat - 1))
