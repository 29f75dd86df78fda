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


import os

# Iterations given to the FIRST attempt at the 2 B maximum / minimum NLPs of a sweep.  A batch advances in lock step and costs its
# slowest problem's iterations whatever its size (tools/gg_profile.py: 70 iterations for 4032 NLPs whose median is below 20); the
# few problems that need more are handed to the continuation from their predecessor, which the reference applies to every point.
FIRST_PASS_ITER = int(os.environ.get("FL_GG_FIRST_ITER", 30))


def _options(tol=1.0e-8, max_iter=300):
    o = default_options()
    o.tol, o.constr_viol_tol, o.acceptable_tol = tol, tol, 1.0e-6        # steady_state.hpp:84-88
    o.refine_steps = int(os.environ.get("FL_GG_REFINE", 1))
    o.max_iter = max_iter     # an NLP of 8-13 variables that has not converged by then is retried by continuation instead
    return o


def solve_batch(params, spec, x0, device=0, bounds=None, max_iter=300):
    """Solves one batch of steady-state NLPs; returns the LapSolver result dict."""
    nlp = SteadyStateNLP(_model_of(params), params, spec, device=device)
    solver = LapSolver(nlp, options=_options(max_iter=max_iter), bounds=bounds)
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


def _base_bounds(params, brake=False, device=0):
    """Bounds of one steady-state NLP: the model's variable bounds (+ ax, ay) and constraint bounds."""
    if _model_of(params) == MODEL_F1:
        return f1_bounds(brake)
    nlp = SteadyStateNLP(MODEL_KART, params if np.ndim(params) == 1 else np.asarray(params)[0], np.zeros((1, 7)) + np.array([20.0, 0, 0, 0, 0, 0, 0]), device=device)
    b = nlp.bounds()
    nlp.close()
    return b


def speed_envelope(params, speeds, device=0):
    """Steps (1)-(3) of the reference's G-G computation for every speed: the 0 g state, the maximum lateral acceleration (with
    its longitudinal acceleration) and the maximum / minimum longitudinal acceleration at ay = 0."""
    speeds = np.asarray(speeds, dtype=np.float64)
    S = len(speeds)
    z, one = np.zeros(S), np.ones(S)
    f1 = _model_of(params) == MODEL_F1
    ia, iy = (11, 12) if f1 else (6, 7)
    out0 = solve_batch(params, np.stack([speeds, z, z, z, z, z, z], axis=1), np.tile(X0_F1 if f1 else X0, (S, 1)), device)
    x_0g = out0["x"].copy()
    x_0g[:, ia] = 0.0; x_0g[:, iy] = 0.0
    bnd_acc, bnd_brk = _base_bounds(params, False, device), _base_bounds(params, True, device)
    spec_lat = np.stack([speeds, z, z, one, one, z, -one], axis=1)
    spec_max = np.stack([speeds, z, z, one, z, -one, z], axis=1)
    spec_min = np.stack([speeds, z, z, one, z, one, z], axis=1)
    if f1:
        # (the F1's braking NLP has its own variable bounds, limebeer2014f1.h:62-80)
        out_lat = solve_batch(params, spec_lat, x_0g, device)
        lon_max = solve_batch(params, spec_max, x_0g, device, bounds=bnd_acc)
        lon_min = solve_batch(params, spec_min, x_0g, device, bounds=bnd_brk)
    else:
        # the three NLPs of a speed start from the same point and share the bounds: ONE batch of 3 S problems (a batch costs its
        # slowest problem's iterations whatever its size, profiles/r2_gg_launches.md)
        p3 = params if np.ndim(params) == 1 else np.tile(np.asarray(params), (3, 1))
        o3 = solve_batch(p3, np.concatenate([spec_lat, spec_max, spec_min]), np.tile(x_0g, (3, 1)), device)
        part = lambda a, b: {key: (val[a * S:b * S] if isinstance(val, np.ndarray) and len(val) == 3 * S else val) for key, val in o3.items()}
        out_lat, lon_max, lon_min = part(0, 1), part(1, 2), part(2, 3)
    return dict(x_0g=x_0g, zero_g=out0["x"].copy(), zero_g_solved=out0["status"] >= 1, x_lat=out_lat["x"].copy(), ay_max=out_lat["x"][:, iy].copy(),
                ax_lat=out_lat["x"][:, ia].copy(), ay_max_solved=out_lat["status"] >= 1, ay_max_iters=out_lat["iters"].copy(),
                ax_lon_max=lon_max["x"][:, ia].copy(), ax_lon_min=lon_min["x"][:, ia].copy(),
                lon_solved=(lon_max["status"] >= 1) & (lon_min["status"] >= 1), bounds_accelerate=bnd_acc, bounds_brake=bnd_brk)


def longitudinal_limits(params, env, k, ay, prev, device=0):
    """Steps (4)-(5) for a list of points: point b belongs to speed env[..][k[b]], has lateral acceleration ay[b] and the point
    prev[b] (an index into the list, or -1) as its predecessor along the fan of its speed.  Returns dict(ax_max, ax_min,
    ax_max_solved, ax_min_solved, iterations, feasible_solved)."""
    k, ay, prev = np.asarray(k, dtype=int), np.asarray(ay, dtype=np.float64), np.asarray(prev, dtype=int)
    B = len(k)
    f1 = _model_of(params) == MODEL_F1
    ia, iy = (11, 12) if f1 else (6, 7)
    unit = 9.81 if f1 else 1.0                       # acceleration_units
    v_b = env["speeds"][k]
    zb, ob = np.zeros(B), np.ones(B)
    p_b = params if np.ndim(params) == 1 else np.asarray(params)[k]
    # (4) a feasible point at (ax_lat ay / ay_max, ay), from the 0 g state; where that fails the predecessor's point is kept
    ax_f = env["ax_lat"][k] * ay / env["ay_max"][k]
    feas = solve_batch(p_b, np.stack([v_b, ax_f, ay, zb, zb, zb, zb], axis=1), env["x_0g"][k], device)
    xf, okf = feas["x"].copy(), feas["status"] >= 1
    axs = ax_f.copy()
    order = np.argsort(ay, kind="stable")
    for b in order:                                  # predecessors first
        if not okf[b]:
            if prev[b] >= 0:
                xf[b], axs[b] = xf[prev[b]], axs[prev[b]]
            else:
                xf[b], axs[b] = env["x_0g"][k[b]], 0.0
    x0 = xf.copy()
    x0[:, ia], x0[:, iy] = axs, ay

    def lap_bounds(base, lo, hi):
        xl, xu = np.tile(base[0], (B, 1)), np.tile(base[1], (B, 1))
        xl[:, ia], xu[:, ia] = lo[k], hi[k]
        return xl, xu, np.tile(base[2], (B, 1)), np.tile(base[3], (B, 1))

    b_max = lap_bounds(env["bounds_accelerate"], env["ax_lat"] - 0.5 / unit, env["ax_lon_max"] + 0.1 / unit)
    b_min = lap_bounds(env["bounds_brake"], env["ax_lon_min"] - 0.1 / unit, env["ax_lat"] + 0.1 / unit)
    res = dict(feasible_solved=okf)
    # the maximum and the minimum of every point are independent NLPs from the same start: ONE batch of 2 B problems with one set of
    # bounds per problem (first half: accelerating, second half: braking)
    two = lambda a: np.concatenate([a, a])
    bnd = tuple(np.concatenate([a, b]) for a, b in zip(b_max, b_min))
    spec = np.concatenate([np.stack([v_b, zb, ay, ob, zb, sg * ob, zb], axis=1) for sg in (-1.0, 1.0)])
    p2 = p_b if np.ndim(p_b) == 1 else two(p_b)
    ay2, prev2 = two(ay), np.concatenate([prev, np.where(prev >= 0, prev + B, -1)])
    start = two(x0)
    start[:, ia] = np.clip(start[:, ia], bnd[0][:, ia] + 1e-6, bnd[1][:, ia] - 1e-6)
    out = solve_batch(p2, spec, start, device, bounds=bnd, max_iter=FIRST_PASS_ITER)
    x, status, iters = out["x"].copy(), out["status"].copy(), out["iters"].copy()
    res["first_pass_iters"] = out["iters"].copy()
    # (5) second attempt from the predecessor's solution (the reference does it for the minimum, steady_state.hpp:801-809; the
    # same continuation serves the maximum here), swept along the fans until nothing changes
    for _ in range(B):
        todo = np.array([b for b in np.where(status < 1)[0] if prev2[b] >= 0 and status[prev2[b]] >= 1], dtype=int)
        if todo.size == 0:
            break
        st2 = x[prev2[todo]].copy()
        st2[:, iy] = ay2[todo]
        st2[:, ia] = np.clip(st2[:, ia], bnd[0][todo, ia] + 1e-6, bnd[1][todo, ia] - 1e-6)
        o2 = solve_batch(p2 if np.ndim(p2) == 1 else p2[todo], spec[todo], st2, device, bounds=tuple(a[todo] for a in bnd))
        good = o2["status"] >= 1
        x[todo[good]] = o2["x"][good]; status[todo[good]] = o2["status"][good]; iters[todo[good]] += o2["iters"][good]
        if not good.any():
            break
    if FIRST_PASS_ITER < 300 and (status < 1).any():
        # what neither the shortened first attempt nor the continuation solved gets the full iteration count from its own start
        todo = np.where(status < 1)[0]
        o2 = solve_batch(p2 if np.ndim(p2) == 1 else p2[todo], spec[todo], start[todo], device, bounds=tuple(a[todo] for a in bnd))
        good = o2["status"] >= 1
        x[todo[good]] = o2["x"][good]; status[todo[good]] = o2["status"][good]; iters[todo[good]] += o2["iters"][good]
    for name, half in (("ax_max", slice(0, B)), ("ax_min", slice(B, 2 * B))):
        res[name], res[name + "_solved"], res[name + "_iters"], res[name + "_x"] = x[half, ia].copy(), status[half] >= 1, iters[half], x[half]
    return res


def gg_diagram(params, speeds, n_lat=64, device=0, polish=0):
    """speeds [m/s].  Returns dict(ay [S][n_lat], ax_max, ax_min [S][n_lat], ay_max [S], solved masks); accelerations in the
    model's units (m/s2 for the kart, g for the F1: acceleration_units of lot2016kart.h:79 / limebeer2014f1.h:51).

    The reference's order of solves (steady_state.hpp:696-880), every step one batch over all speeds (and lateral accelerations):
      (1) 0 g;  (2) maximum lateral acceleration;  (3) maximum / minimum longitudinal acceleration at ay = 0;
      (4) for every lateral acceleration ay_i of the fan a FEASIBLE point at (ax_lat ay_i / ay_max, ay_i), solved from the 0 g
          state (where that fails the previous lateral acceleration's point is kept, :715-723);
      (5) from that point the maximum and the minimum longitudinal acceleration, with the reference's bounds on ax
          (accelerating: [ax_lat - 0.5, ax_max(0) + 0.1], braking: [ax_min(0) - 0.1, ax_lat + 0.1], :733-735, 787-789); a minimum
          that fails is solved again from the previous lateral acceleration's minimum (:801-809);
      (6) the last point of both curves is the maximum-lateral-acceleration solution (:857-870)."""
    speeds = np.asarray(speeds, dtype=np.float64)
    S, L = len(speeds), int(n_lat)
    env = speed_envelope(params, speeds, device)
    env["speeds"] = speeds
    n_f = L - 1
    t = np.linspace(0.0, 1.0, L)[:n_f]
    ay = t[None, :] * env["ay_max"][:, None]                             # [S][L-1]
    k = np.repeat(np.arange(S), n_f)
    idx = np.arange(S * n_f)
    prev = np.where(idx % n_f > 0, idx - 1, -1)
    lim = longitudinal_limits(params, env, k, ay.ravel(), prev, device)
    res = {}
    for name in ("ax_max", "ax_min"):
        res[name] = np.concatenate([lim[name].reshape(S, n_f), env["ax_lat"][:, None]], axis=1)                        # (6)
        res[name + "_solved"] = np.concatenate([lim[name + "_solved"].reshape(S, n_f), env["ay_max_solved"][:, None]], axis=1)
        res[name + "_iters"] = np.concatenate([lim[name + "_iters"].reshape(S, n_f), env["ay_max_iters"][:, None]], axis=1)
    res.update(ay=np.concatenate([ay, env["ay_max"][:, None]], axis=1), ay_max=env["ay_max"], ax_at_ay_max=env["ax_lat"], zero_g=env["zero_g"],
               zero_g_solved=env["zero_g_solved"], ay_max_solved=env["ay_max_solved"], feasible_solved=lim["feasible_solved"].reshape(S, n_f),
               ax_lon_max=env["ax_lon_max"], ax_lon_min=env["ax_lon_min"], lon_solved=env["lon_solved"])
    return res
