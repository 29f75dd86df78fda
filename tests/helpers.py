"""Shared test helpers: the CPU oracle behind the same LapProblem description the CUDA path takes."""
import os
import numpy as np

from fastest_lap_b200.problem import LapProblem
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """(problem, golden point).  Runs with a restricted integral quantity come with the point mapped to the node-wise formulation
    (augmented_point); "f1_catalunya_hypermesh" is the Catalunya golden with the brake bias on a four-stretch hypermesh, every
    stretch holding the vehicle's brake bias (a feasible point of that problem; there is no reference run of it with a vehicle of the database)."""
    if name == "f1_catalunya_hypermesh":
        p0, ex0 = LapProblem.load_npz(os.path.join(GOLDEN, "f1_catalunya_discrete.npz"))
        p = p0.with_hypermesh_brake_bias([0.0, 1000.0, 2500.0, 4000.0])
        ex = dict(ex0)
        for k in ("x", "zl", "zu"):
            v = np.zeros((p.n_points, p.nv))
            r = ex0[k].reshape(p.n_points, p.nv - 2)
            v[:, :p.aug_pos], v[:, p.aug_pos + 2:] = r[:, :p.aug_pos], r[:, p.aug_pos:]
            if k == "x":
                v[:, p.aug_pos] = p.params[18]
            ex[k] = v.ravel()
        row = p.ncpe - 6 - 4 - 1
        lam = np.zeros((p.n_elements, p.ncpe))
        lr = ex0["lam"].reshape(p.n_elements, p.ncpe - 1)
        lam[:, :row], lam[:, row + 1:] = lr[:, :row], lr[:, row:]
        ex["lam"] = lam.ravel()
        return p, ex
    p, ex = LapProblem.load_npz(os.path.join(GOLDEN, name + ".npz"))
    if "x_ref" in ex:
        ex = dict(ex)
        ex["x"], ex["lam"], ex["zl"], ex["zu"] = augmented_point(p, ex)
    return p, ex


def oracle_nlp(p, extended=False):
    """extended: the oracle built with its jets in long double (oracle/jet.hpp) -- the arbiter of ill-conditioned entries."""
    return orc.OracleNLP(p.model, p.form, p.closed, p.s, p.track_length, p.track_nodes, p.params, p.dissipation, p.sigma, p.elevation,
                         ctrl_default=p.ctrl_default, q0=p.q0, u0=p.u0, du0=p.du0, kart_direct_torque=p.kart_direct_torque,
                         node_params=getattr(p, "node_params", None), wind=getattr(p, "wind", (0.0, 0.0)), extended=extended, aug=getattr(p, "aug", None))


def equality_rows(p):
    """Row offsets (within an element) of the equality constraints: trapezoid + algebraic rows, control-rate rows."""
    n_eq = p.ncpe - 6 - (0 if p.form == 0 else 2)
    rows = list(range(n_eq))
    if p.form != 0:
        rows += [p.ncpe - 2, p.ncpe - 1]
    return np.array(rows)


def rel_err(a, b, floor=1.0):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0


# ---- 40-digit evaluation of the symbolic node program (the ground truth where fp64 conditioning limits entry-wise parity)
def truth_node(p, x, lam, obj, node, digits=40):
    """Exact (mpmath) gradient / Jacobian blocks / Hessian block of one collocation node, from the same expression DAG the
    CUDA code is generated from, evaluated with `digits` significant digits.  Returns dict of {(row, col): float}."""
    import mpmath as mp
    from fastest_lap_b200.codegen.nlpnode import Variant, NodeProgram
    from fastest_lap_b200.codegen import expr as X
    mp.mp.dps = digits
    key = (p.model, p.form, p.elevation, p.kart_direct_torque)
    prog = _PROGRAMS.get(key)
    if prog is None:
        prog = _PROGRAMS[key] = NodeProgram(Variant("f1" if p.model == 0 else "kart", "direct" if p.form == 0 else "derivative", p.elevation,
                                                    p.kart_direct_torque))
    g = prog.g
    NV, NCPE, n = prog.nv, prog.ncpe, p.n_points
    node0 = 0 if p.closed else 1
    xfull = np.zeros((n, NV))
    xfull[node0:] = np.asarray(x).reshape(-1, NV)
    if not p.closed:
        time_idx = (14 if p.model == 0 else 13) - 3
        q = [p.q0[j] for j in range(len(p.q0)) if j != time_idx]
        u = [p.u0[0], p.u0[2]] if p.model == 0 else [p.u0[0], p.u0[1]]
        xfull[0, :len(q) + 2] = q + u
        if p.form != 0:
            xfull[0, len(q) + 2:] = p.du0
    lam = np.asarray(lam).reshape(-1, NCPE)
    s, L = p.s, p.track_length
    hin = s[node] - s[node - 1] if node > 0 else (L - s[-1] if p.closed else 0.0)
    hout = s[node + 1] - s[node] if node + 1 < n else (L - s[-1] if p.closed else 0.0)
    prev, nxt = (node - 1) % n, (node + 1) % n
    env = {"x%d" % k: xfull[node, k] for k in range(NV)}
    th, mu, ph, dth, dmu, dph = p.track_nodes[node]
    trk = dict(kx=dph - np.sin(mu) * dth, ky=np.cos(ph) * dmu + np.cos(mu) * np.sin(ph) * dth, kz=-np.sin(ph) * dmu + np.cos(mu) * np.cos(ph) * dth,
               smu=np.sin(mu), cmu=np.cos(mu), sph=np.sin(ph), cph=np.cos(ph), hin=hin, hout=hout,
               ihin=1.0 / hin if hin > 0 else 0.0, ihout=1.0 / hout if hout > 0 else 0.0, wt=0.0, wn=0.0, wb=0.0)      # (no wind in the golden runs)
    if p.form != 0:
        penw = 0.0
        for e in range(p.n_elements):
            closing = e + 1 == n
            he = L - s[-1] if closing else s[e + 1] - s[e]
            penw += he * ((1 if (0 if closing else e + 1) == node else 0) + (1 if (n - 1 if closing else 0) == node else 0))
        trk["penw"] = penw
    for nm in prog.track_names + prog.geom_names:
        env["t_" + nm] = trk[nm]
    m = p.params[0]
    host = dict(sigma=p.sigma, diss0=p.dissipation[0], diss1=p.dissipation[1])
    if p.model == 0:
        host.update(f_wheel_scale=1.0 / m if p.params[22] < 1e-10 else 1.0, r_wheel_scale=1.0 / m if p.params[26] < 1e-10 else 1.0)
    for k, nm in enumerate(prog.param_names):
        env["p_" + nm] = p.params[k] if k < len(p.params) else host[nm]
    e_in = prev
    has_out = p.closed or node + 1 < n
    for r in range(NCPE):
        env["lin%d" % r] = lam[e_in, r]
        env["lout%d" % r] = lam[node, r] if has_out else 0.0
    env["obj"] = obj
    for c in range(len(prog.uprev)):
        env["uprev%d" % c] = xfull[prev, prog.ctrl_pos[c]]
        env["unext%d" % c] = xfull[nxt, prog.ctrl_pos[c]]
    fn = {X.SQRT: mp.sqrt, X.SIN: mp.sin, X.COS: mp.cos, X.ATAN: mp.atan, X.TANH: mp.tanh, X.EXP: mp.exp}
    groups = dict(hess=prog.hess, jacA=prog.jacA, jacB=prog.jacB, grad={(k, 0): e for k, e in enumerate(prog.gradf)})
    roots = [e for d in groups.values() for e in d.values()]
    val = {}
    for i in g.topo(roots):
        k, a, b = g.kind[i], g.a[i], g.b[i]
        if k == X.CONST: val[i] = mp.mpf(a)
        elif k == X.SYM: val[i] = mp.mpf(float(env[a]))
        elif k == X.ADD: val[i] = val[a] + val[b]
        elif k == X.SUB: val[i] = val[a] - val[b]
        elif k == X.MUL: val[i] = val[a] * val[b]
        elif k == X.NEG: val[i] = -val[a]
        elif k == X.RCP: val[i] = 1 / val[a]
        else: val[i] = fn[k](val[a])
    return {name: {key_: float(val[e]) for key_, e in d.items()} for name, d in groups.items()}


_PROGRAMS = {}


def state_control_deviation(x, x_ref, problem=None):
    """Largest deviation of a converged solution from the reference's stored one, relative to max(1, |value|): `north_star`
    asks for converged controls within 1e-6 relative; the reference's own comparator is EXPECT_NEAR(..., 1e-6) on every state
    and control (src/test/applications/f1_optimal_laptime_test.cpp:19-165).  With `problem` the result is the pair
    (states, controls): the full-mesh controls are the two columns after the states (and, derivative form, their rates)."""
    x, x_ref = np.asarray(x, dtype=np.float64).ravel(), np.asarray(x_ref, dtype=np.float64).ravel()
    d = np.abs(x - x_ref) / np.maximum(1.0, np.abs(x_ref))
    if problem is None:
        return float(np.max(d))
    D = d.reshape(-1, problem.nv)
    n_state = problem.nv - (2 if problem.form == 0 else 4)
    return float(np.max(D[:, :n_state])), float(np.max(D[:, n_state:n_state + 2]))


def oracle_kkt(p, out):
    """Stationarity, feasibility and complementarity of a returned point, all from the oracle's callbacks."""
    import scipy.sparse as sp
    from fastest_lap_b200.nlp import problem_bounds
    x, y, zl, zu = out["x"][0], out["y"][0], out["zl"][0], out["zu"][0]
    r = oracle_nlp(p).eval(x, y, 1.0)
    J = sp.csr_matrix((r["jac"][2], (r["jac"][0], r["jac"][1])), shape=(p.m, p.n))
    xl, xu, gl, gu = problem_bounds(p)
    stat = np.abs(r["grad"] + J.T @ y - zl + zu).max()
    g = r["g"]
    eq = gu <= gl
    viol_eq = np.abs(g[eq] - gl[eq]).max()
    # (bounds are relaxed by 1e-8 max(1, |b|) as in Ipopt: a point may sit that far outside)
    viol_bnd = max(np.maximum(gl - g, 0).max(), np.maximum(g - gu, 0).max(), np.maximum(xl - x, 0).max(), np.maximum(x - xu, 0).max())
    comp = max(np.abs((x - xl) * zl).max(), np.abs((xu - x) * zu).max())
    return stat, viol_eq, viol_bnd, comp


def augmented_point(p, ex):
    """The reference's converged point of a run with ONE restricted integral quantity (x_ref: 15 variables per node; lam_ref: 19
    multipliers per element; nu: the multiplier of the integral's row) in the node-wise formulation of the product and the oracle
    (codegen/nlpnode.py): per node [inputs, a, q, controls], per element one more trapezoid row after the model's.
    a = the integral accumulated up to the node (a_0 = the whole integral), obtained from the augmented rows themselves, which are
    linear in a; every augmented row carries the multiplier -nu, and a_0's active bound the multiplier |nu|."""
    nv, a_pos = p.nv, p.aug_pos
    x = np.zeros((p.n_points, nv))
    xr = ex["x_ref"].reshape(p.n_points, nv - 2)
    x[:, :a_pos] = xr[:, :a_pos]
    x[:, a_pos + 2:] = xr[:, a_pos:]
    row = p.ncpe - 6 - 4 - 1                                    # the last trapezoid row of the element (flat F1: 4 algebraic rows follow)
    g = oracle_nlp(p).eval(x.ravel(), derivs=False)["g"].reshape(p.n_elements, p.ncpe)
    c = -g[:, row]                                              # with a = 0: row = -(awout I_i + awin I_j)
    a = np.zeros(p.n_points)
    a[1:] = np.cumsum(c[:-1])
    a[0] = a[-1] + c[-1]
    x[:, a_pos] = a
    lam = np.zeros((p.n_elements, p.ncpe))
    lr = ex["lam_ref"].reshape(p.n_elements, p.ncpe - 1)
    lam[:, :row] = lr[:, :row]
    lam[:, row] = -float(ex["nu"])
    lam[:, row + 1:] = lr[:, row:]
    zl, zu = np.zeros((p.n_points, nv)), np.zeros((p.n_points, nv))
    for z, zr in ((zl, ex["zl_ref"]), (zu, ex["zu_ref"])):
        zr = zr.reshape(p.n_points, nv - 2)
        z[:, :a_pos] = zr[:, :a_pos]
        z[:, a_pos + 2:] = zr[:, a_pos:]
    (zu if float(ex["nu"]) > 0.0 else zl)[0, a_pos] = abs(float(ex["nu"]))
    return x.ravel(), lam.ravel(), zl.ravel(), zu.ravel()
