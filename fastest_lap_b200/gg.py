"""G-G diagram of the kart as batched steady-state NLPs on the GPU.

Mirror of Steady_state<...>::gg_diagram (src/core/applications/steady_state.hpp:594-880) and of the `gg_diagram` entry of
the C API (src/main/c/fastestlapc.cpp): for every speed, (1) the 0 g steady state, (2) the maximum lateral acceleration,
(3) for a fan of lateral accelerations the maximum and the minimum longitudinal acceleration.  The reference solves these
~130 NLPs per speed one after the other, each warm-started from the previous one; here every stage is ONE batch of
independent NLPs over all speeds (and all lateral accelerations), each started from the 0 g state of its speed.
"""
import numpy as np

from .nlp import SteadyStateNLP
from .ipm import LapSolver, default_options
from .vehicle import MODEL_KART, MODEL_F1

X0 = np.array([0.0, 0.02, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])          # lot2016kart::steady_state_initial_guess (lot2016kart.h:83-86)
X0_F1 = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -0.25, -0.25, -0.25, -0.25, 0.0, 0.0, 0.0, 0.0])   # limebeer2014f1.h:57-60 (+ ax, ay)


def _model_of(params):
    return MODEL_F1 if np.shape(params)[-1] == 58 else MODEL_KART


def _options(tol=1.0e-8):
    o = default_options()
    o.tol, o.constr_viol_tol, o.acceptable_tol = tol, tol, 1.0e-6        # steady_state.hpp:84-88
    o.max_iter = 300          # an NLP of 8-13 variables that has not converged by then is retried by continuation instead
    return o


def solve_batch(params, spec, x0, device=0, bounds=None):
    """Solves one batch of steady-state NLPs; returns the LapSolver result dict."""
    nlp = SteadyStateNLP(_model_of(params), params, spec, device=device)
    solver = LapSolver(nlp, options=_options(), bounds=bounds)
    out = solver.solve(x0)
    solver.close(); nlp.close()
    return out


def f1_bounds(brake=False):
    """limebeer2014f1::steady_state_variable_bounds[_brake] (limebeer2014f1.h:62-80) + ax, ay in g, and the constraint bounds."""
    if brake:
        xl = [-1.15, -1.15, -1.30, -1.30, 0.0, -2.0, -2.0, -2.0, -2.0, 0.0, -1.0, -8.0, -8.0]
        xu = [0.25, 0.25, 0.25, 0.25, 1.0, 0.1, 0.1, 0.1, 0.1, 1.0, 0.5, 8.0, 8.0]
    else:
        xl = [-1.15, -1.15, -1.15, -1.15, 0.0, -2.0, -2.0, -2.0, -2.0, 0.0, -1.0, -8.0, -8.0]
        xu = [0.25, 0.25, 1.00, 1.00, 1.5, 0.1, 0.1, 0.1, 0.1, 1.5, 1.0, 8.0, 8.0]
    return np.array(xl), np.array(xu), np.array([0.0] * 11 + [-1.0] * 4), np.array([0.0] * 11 + [1.2] * 4)


def solve_with_continuation(params, spec, x0, group, order, device=0, max_sweeps=8, bounds=None, polish=0, objective=None):
    """Solves a batch; problems that fail are solved again from the converged solution of their nearest predecessor in
    `order` within the same `group` (the reference warm-starts every G-G point from the previous lateral acceleration,
    steady_state.hpp:615-633: this is the same continuation, applied only where the cold start fails)."""
    spec, x0 = np.asarray(spec, dtype=np.float64), np.array(x0, dtype=np.float64)
    out = solve_batch(params, spec, x0, device, bounds)
    x, status, iters = out["x"].copy(), out["status"].copy(), out["iters"].copy()
    for _ in range(max_sweeps):
        bad = np.where(status < 1)[0]
        if bad.size == 0:
            break
        starts, todo = [], []
        for b in bad:
            cand = np.where((group == group[b]) & (order < order[b]) & (status >= 1))[0]
            if cand.size:
                k = cand[np.argmax(order[cand])]
                starts.append(x[k]); todo.append(b)
        if not todo:
            break
        todo = np.array(todo)
        p_b = params if np.ndim(params) == 1 else np.asarray(params)[todo]
        o2 = solve_batch(p_b, spec[todo], np.array(starts), device, bounds)
        good = o2["status"] >= 1
        if not good.any():
            break
        x[todo[good]] = o2["x"][good]; status[todo[good]] = o2["status"][good]; iters[todo[good]] += o2["iters"][good]
    # polish: these NLPs have several local optima; the reference follows one branch by warm-starting every point from
    # the previous lateral acceleration.  Every point is solved again from its predecessor's solution and the better
    # optimum is kept.
    for _ in range(polish):
        idx, starts = [], []
        for b in range(len(spec)):
            cand = np.where((group == group[b]) & (order < order[b]) & (status >= 1))[0]
            if cand.size:
                idx.append(b); starts.append(x[cand[np.argmax(order[cand])]])
        if not idx:
            break
        idx = np.array(idx)
        p_b = params if np.ndim(params) == 1 else np.asarray(params)[idx]
        o2 = solve_batch(p_b, spec[idx], np.array(starts), device, bounds)
        better = (o2["status"] >= 1) & ((status[idx] < 1) | (objective(o2["x"], spec[idx]) < objective(x[idx], spec[idx]) - 1.0e-9))
        x[idx[better]] = o2["x"][better]; status[idx[better]] = o2["status"][better]
        if not better.any():
            break
    return dict(x=x, status=status, iters=iters)


def gg_diagram(params, speeds, n_lat=64, device=0, polish=3):
    """speeds [m/s].  Returns dict(ay [S][n_lat], ax_max, ax_min [S][n_lat], ay_max [S], solved masks); accelerations in the
    model's units (m/s2 for the kart, g for the F1: acceleration_units of lot2016kart.h:79 / limebeer2014f1.h:51)."""
    speeds = np.asarray(speeds, dtype=np.float64)
    S = len(speeds)
    z = np.zeros(S)
    f1 = _model_of(params) == MODEL_F1
    ia, iy = (11, 12) if f1 else (6, 7)
    # (1) 0 g
    out0 = solve_batch(params, np.stack([speeds, z, z, z, z, z, z], axis=1), np.tile(X0_F1 if f1 else X0, (S, 1)), device)
    x_0g = out0["x"].copy()
    # (2) maximum lateral acceleration: ax and ay free, f = -ay
    one = np.ones(S)
    out_lat = solve_batch(params, np.stack([speeds, z, z, one, one, z, -one], axis=1), x_0g, device)
    ay_max, ax_at_max = out_lat["x"][:, iy], out_lat["x"][:, ia]
    # (3) fan of lateral accelerations: max / min longitudinal acceleration (f = -+ax), ay given
    ay = np.linspace(0.0, 1.0, n_lat)[None, :] * ay_max[:, None]
    v_b = np.repeat(speeds, n_lat)
    zb, ob = np.zeros(S * n_lat), np.ones(S * n_lat)
    x0 = np.repeat(x_0g, n_lat, axis=0)
    x0[:, iy] = 0.0
    res = {}
    grp, order = np.repeat(np.arange(S), n_lat), np.tile(np.arange(n_lat), S)
    for name, sign in (("ax_max", -1.0), ("ax_min", 1.0)):
        bnd = f1_bounds(brake=sign > 0) if f1 else None         # the F1 brakes within its own variable bounds (steady_state.hpp:803-806)
        out = solve_with_continuation(params, np.stack([v_b, zb, ay.ravel(), ob, zb, sign * ob, zb], axis=1), x0, grp, order, device, bounds=bnd,
                                      polish=polish, objective=lambda xx, sp: sp[:, 5] * xx[:, ia])
        res[name] = out["x"][:, ia].reshape(S, n_lat)
        res[name + "_solved"] = (out["status"] == 1).reshape(S, n_lat) | (out["status"] == 2).reshape(S, n_lat)
        res[name + "_iters"] = out["iters"].reshape(S, n_lat)
        # the last point is the maximum-lateral-acceleration solution itself (steady_state.hpp:715, 857-870: the loop stops at
        # n_points - 1): there the feasible set of the NLP has shrunk to that single point
        res[name][:, -1] = ax_at_max
        res[name + "_solved"][:, -1] = out_lat["status"] >= 1
    res.update(ay=ay, ay_max=ay_max, ax_at_ay_max=ax_at_max, zero_g=x_0g, zero_g_solved=out0["status"] >= 1, ay_max_solved=out_lat["status"] >= 1)
    return res
