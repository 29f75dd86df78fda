"""Host front-end of the device-resident interior-point solver (include/fastestlap_b200.h, "interior-point solver").

Mirror of the reference's hand-off `ipopt_cppad_solve(options, x0, xl, xu, gl, gu, fg, result)` and of the fields it reads
back from `ipopt_cppad_result` (src/core/applications/optimal_laptime.hpp:546-556, 590-647): status, x, zl, zu, lambda,
objective, iteration count -- for every lap of a batch at once.  No CPU path: everything runs in libfastestlap_b200.so.
"""
import ctypes as C
import numpy as np

from .nlp import CollocationNLP, load_library, _ptr, _dp

STATUS = {0: "running", 1: "success", 2: "acceptable", -1: "max_iter", -2: "line_search_failed", -3: "regularisation_failed", -4: "not_finite"}


class IpmOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("constr_viol_tol", C.c_double), ("acceptable_tol", C.c_double), ("acceptable_iter", C.c_int), ("max_iter", C.c_int),
                ("mu_init", C.c_double), ("bound_push", C.c_double), ("bound_frac", C.c_double), ("bound_relax_factor", C.c_double),
                ("kappa_eps", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double), ("tau_min", C.c_double), ("kappa_sigma", C.c_double),
                ("s_max", C.c_double), ("gamma_theta", C.c_double), ("gamma_phi", C.c_double), ("delta", C.c_double), ("s_theta", C.c_double),
                ("s_phi", C.c_double), ("eta_phi", C.c_double), ("kappa_soc", C.c_double), ("alpha_red", C.c_double), ("max_soc", C.c_int),
                ("max_backtracks", C.c_int), ("delta_w_min", C.c_double), ("delta_w_0", C.c_double), ("delta_w_max", C.c_double),
                ("kappa_w_minus", C.c_double), ("kappa_w_plus", C.c_double), ("kappa_w_plus_bar", C.c_double), ("delta_c_bar", C.c_double),
                ("kappa_c", C.c_double), ("y_init_max", C.c_double), ("refine_steps", C.c_int)]


_ip = C.POINTER(C.c_int)


def _bind(lib):
    if getattr(lib, "_ipm_bound", False):
        return
    lib.fl_ipm_default_options.argtypes = [C.POINTER(IpmOptions)]
    lib.fl_ipm_create.restype = C.c_void_p
    lib.fl_ipm_create.argtypes = [C.c_void_p, C.POINTER(IpmOptions), _dp, _dp, _dp, _dp]
    lib.fl_ipm_create_ex.restype = C.c_void_p
    lib.fl_ipm_create_ex.argtypes = [C.c_void_p, C.POINTER(IpmOptions), _dp, _dp, _dp, _dp, C.c_int]
    lib.fl_ipm_destroy.argtypes = [C.c_void_p]
    lib.fl_ipm_solve.argtypes = [C.c_void_p, _dp, C.c_int]
    lib.fl_ipm_results.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp]
    lib.fl_ipm_stats.argtypes = [C.c_void_p, C.POINTER(C.c_long), _dp]
    lib.fl_kkt_factor_solve.argtypes = [C.c_void_p] + [_dp] * 8 + [C.c_int, C.c_int, _dp, _dp, _ip, _ip]
    lib.fl_ipm_set_partitioned.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.fl_ipm_set_warp_kkt.argtypes = [C.c_void_p, C.c_int]
    lib.fl_ipm_solve_warm.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int]
    lib.fl_ipm_sigma.argtypes = [C.c_void_p, _dp, _dp, _dp]
    lib._ipm_bound = True


def default_options():
    lib = load_library()
    _bind(lib)
    o = IpmOptions()
    lib.fl_ipm_default_options(C.byref(o))
    return o


class LapSolver:
    """Interior-point solver bound to one CollocationNLP (one batch of laps on one GPU)."""

    def __init__(self, nlp, options=None, bounds=None):
        assert isinstance(nlp, CollocationNLP)
        self.nlp, self._lib = nlp, load_library()
        _bind(self._lib)
        self.options = options
        args = [None] * 4
        per_lap = 0
        if bounds is not None:
            # one set for the batch ([n], [n], [m], [m]) or one per problem ([batch][n], ...)
            self._bounds = [np.ascontiguousarray(b, dtype=np.float64) for b in bounds]
            per_lap = int(self._bounds[0].size == nlp.batch * nlp.n and nlp.batch > 1)
            for b, sz in zip(self._bounds, (nlp.n, nlp.n, nlp.m, nlp.m)):
                if b.size != (nlp.batch * sz if per_lap else sz):
                    raise ValueError("bounds must have n, n, m, m entries (or batch times that, all four)")
            args = [_ptr(b) for b in self._bounds]
        self._h = self._lib.fl_ipm_create_ex(nlp._h, C.byref(options) if options is not None else None, *args, per_lap)
        if not self._h:
            raise RuntimeError("fl_ipm_create: " + self._lib.fl_last_error().decode())
        nlp._register_dependent(self)

    def set_partitioned(self, policy=0, max_laps=0):
        """Schedule of the KKT factorisation: 0 automatic, 1 always parallel in the stages, -1 always one block per lap.
        Returns the number of separator stages (0: not available for this problem)."""
        return int(self._lib.fl_ipm_set_partitioned(self._h, int(policy), int(max_laps)))

    def set_warp_kkt(self, mode=0):
        """Mapping of the unpartitioned factorisation: 1 one warp per lap (registers), -1 one block per lap, 0 automatic
        (warp per lap for batches of 64 laps or more).  Returns True if the warp-per-lap kernels are in use."""
        return bool(self._lib.fl_ipm_set_warp_kkt(self._h, int(mode)))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fl_ipm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, x0, verbose=False, warm=None, warm_push=1.0e-3, warm_mu=1.0e-4):
        """x0: [batch][n] (or [n], replicated).  warm = (lambda, zl, zu) of a previous solution starts the iteration from
        it (the reference's warm start, optimal_laptime.hpp:550-551).  Returns a dict of per-lap results."""
        nlp, B = self.nlp, self.nlp.batch
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.size == nlp.n:
            x0 = np.tile(x0.reshape(1, -1), (B, 1))
        if x0.size != B * nlp.n:
            raise ValueError("x0 must have %d or %d x %d entries" % (nlp.n, B, nlp.n))
        x0 = np.ascontiguousarray(x0.reshape(B, nlp.n))
        if warm is not None:
            w = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1, sz) if np.size(a) != sz else np.asarray(a, dtype=np.float64).reshape(1, sz), (B, sz)))
                 for a, sz in zip(warm, (nlp.m, nlp.n, nlp.n))]
            ok = self._lib.fl_ipm_solve_warm(self._h, _ptr(x0), _ptr(w[0]), _ptr(w[1]), _ptr(w[2]), float(warm_push), float(warm_mu), int(verbose))
        else:
            ok = self._lib.fl_ipm_solve(self._h, _ptr(x0), int(verbose))
        if not ok:
            raise RuntimeError("fl_ipm_solve: " + self._lib.fl_last_error().decode())
        x, y = np.empty((B, nlp.n)), np.empty((B, nlp.m))
        zl, zu = np.empty((B, nlp.n)), np.empty((B, nlp.n))
        obj, err = np.empty(B), np.empty((B, 4))
        status, iters = np.empty(B, np.int32), np.empty(B, np.int32)
        ip = lambda a: a.ctypes.data_as(_ip)
        if not self._lib.fl_ipm_results(self._h, _ptr(x), _ptr(y), _ptr(zl), _ptr(zu), _ptr(obj), ip(status), ip(iters), _ptr(err)):
            raise RuntimeError("fl_ipm_results: " + self._lib.fl_last_error().decode())
        stats = (C.c_long * 6)()
        ms = C.c_double()
        self._lib.fl_ipm_stats(self._h, stats, C.byref(ms))
        return dict(x=x, y=y, zl=zl, zu=zu, obj=obj, status=status, iters=iters, error=err[:, 0], constr_viol=err[:, 1], dual_inf=err[:, 2], mu=err[:, 3],
                    device_ms=ms.value, stats=dict(zip(("solver_launches", "factor_launches", "solve_launches", "full_rounds", "fg_rounds", "outer_rounds"), list(stats))))

    def sigma(self):
        """Barrier terms of the points the solver holds (after solve(): the returned ones): sigx [batch][n], sigs [batch][n_ineq],
        dc [batch] -- what kkt_factor_solve needs to solve with the KKT matrix of the solution (lap_sensitivities)."""
        nlp, B = self.nlp, self.nlp.batch
        sigx, sigs, dc = np.empty((B, nlp.n)), np.empty((B, nlp.problem.n_elements * 6)), np.empty(B)
        if not self._lib.fl_ipm_sigma(self._h, _ptr(sigx), _ptr(sigs), _ptr(dc)):
            raise RuntimeError("fl_ipm_sigma: " + self._lib.fl_last_error().decode())
        return sigx, sigs, dc

    def kkt_factor_solve(self, x, lam, sigx, sigs, dw, dc, r1, r2, refine_steps=0, ls_mode=False):
        """One factorisation + solve of the augmented system for every lap (building block / test entry)."""
        nlp, B = self.nlp, self.nlp.batch
        f64 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(shape if np.size(a) == np.prod(shape) else (1,) + shape[1:]), shape))
        x, lam = f64(x, (B, nlp.n)), f64(lam, (B, nlp.m))
        sigx, r1, r2 = f64(sigx, (B, nlp.n)), f64(r1, (B, nlp.n)), f64(r2, (B, nlp.m))
        sigs = f64(sigs, (B, np.size(sigs) // (B if np.ndim(sigs) > 1 else 1)))
        dw = np.ascontiguousarray(np.broadcast_to(np.asarray(dw, dtype=np.float64), (B,)))
        dc = np.ascontiguousarray(np.broadcast_to(np.asarray(dc, dtype=np.float64), (B,)))
        dx, dy = np.empty((B, nlp.n)), np.empty((B, nlp.m))
        neg, zero = np.empty(B, np.int32), np.empty(B, np.int32)
        ip = lambda a: a.ctypes.data_as(_ip)
        if not self._lib.fl_kkt_factor_solve(self._h, _ptr(x), _ptr(lam), _ptr(sigx), _ptr(sigs), _ptr(dw), _ptr(dc), _ptr(r1), _ptr(r2), int(refine_steps),
                                             int(ls_mode), _ptr(dx), _ptr(dy), ip(neg), ip(zero)):
            raise RuntimeError("fl_kkt_factor_solve: " + self._lib.fl_last_error().decode())
        return dx, dy, neg, zero


def lap_time(problem, x, obj):
    """Lap time of a closed lap's solution: the objective minus the control-rate penalty -- direct form
    diss ((u_j - u_i) / h)^2 h per element (optimal_laptime.hpp:1358-1368); derivative form 0.5 diss (du_l^2 + du_r^2) h with
    the left rate taken at node 0 on every non-closing element, as the reference does (:1553, :1611)."""
    from .problem import FORM_DIRECT
    X = np.asarray(x).reshape(-1, problem.nv)
    if not problem.closed:
        raise NotImplementedError("lap time from the objective: closed laps")
    h = np.diff(np.append(problem.s, problem.track_length))
    pen = 0.0
    if problem.form == FORM_DIRECT:
        n_in = problem.nv - 2
        for c in range(2):
            du = np.roll(X[:, n_in + c], -1) - X[:, n_in + c]
            pen += problem.dissipation[c] * np.sum(du * du / h)
    else:
        n_in = problem.nv - 4
        for c in range(2):
            rate = X[:, n_in + 2 + c]
            left = np.full(len(h), rate[0]); left[-1] = rate[-1]
            pen += 0.5 * problem.dissipation[c] * np.sum((left * left + np.roll(rate, -1) ** 2) * h)
    return float(obj) - pen


def lap_sensitivities(problem, x, lam, sigma, perturbed, device=0):
    """Derivatives of a converged lap with respect to vehicle parameters -- the reference's sensitivity analysis
    (Optimal_laptime with check_optimality, optimal_laptime.hpp:564-588: lion's Sensitivity_analysis solves the linearised
    optimality conditions for dx/dp).  Here: the KKT matrix of the solution, [W + Sigma_x, J^T; J, -D] with the barrier terms of the
    final iterate (`sigma` = LapSolver.sigma() taken before the solver was closed), is factorised once per parameter by the device
    kernels (kkt_factor_solve) and solved for  -d/dp [grad f + J^T lambda; g]  at fixed (x, lambda); that right-hand side is a central
    difference of the GPU callbacks over `perturbed` = [(problem with p + h, problem with p - h, 2 h), ...] -- the callbacks are
    closed-form in the parameters, a relative step of 1e-6 leaves a relative error of about 1e-9, three digits below the 1e-6 the
    reference's own test asks of the result (f1_sensitivity_analysis.cpp:24-26, against re-solved laps).
    Returns (dx_dp [P][n], dlambda_dp [P][m])."""
    x, lam = np.ascontiguousarray(x, dtype=np.float64).ravel(), np.ascontiguousarray(lam, dtype=np.float64).ravel()
    sigx, sigs, dc = sigma

    def grad_lagrangian(q):
        nlp = CollocationNLP(q, device=device)
        try:
            jr, jc, _, _ = nlp.structure()
            o = nlp.eval_all(x, lam, 1.0, want=("g", "grad", "jac"))
            gl = o["grad"][0] + np.bincount(jc, weights=o["jac"][0] * lam[jr], minlength=nlp.n)
            return gl, o["g"][0].copy()
        finally:
            nlp.close()

    nlp = CollocationNLP(problem, device=device)
    solver = LapSolver(nlp)
    dxs, dys = [], []
    try:
        for plus, minus, two_h in perturbed:
            (g1, c1), (g0, c0) = grad_lagrangian(plus), grad_lagrangian(minus)
            r1, r2 = -(g1 - g0) / two_h, -(c1 - c0) / two_h
            dx, dy, neg, zero = solver.kkt_factor_solve(x, lam, sigx[0], sigs[0], 0.0, dc[0], r1, r2, refine_steps=3)
            dxs.append(dx[0].copy()); dys.append(dy[0].copy())
    finally:
        solver.close(); nlp.close()
    return np.array(dxs), np.array(dys)


def dlaptime_dp(problem, x, dx):
    """Derivative of the lap time from dx/dp, as the reference integrates it (optimal_laptime.hpp:664-724): the trapezoid rule over
    d/dp of dt/ds = (1 - kz n) / (u cos(alpha) - v sin(alpha)), which depends on the parameters through the states only.
    Flat or 3-D tracks (kz = the z component of the road-frame curvature).  Closed laps: the closing element included; open sections:
    the first node is given, its derivative is zero (:651-659) and the lap time is the time of the last node.
    Returns (d laptime / dp, d/dp of dt/ds at every node of the mesh)."""
    from .problem import N_INPUTS
    X, DX = np.asarray(x).reshape(-1, problem.nv), np.asarray(dx).reshape(-1, problem.nv)
    iu, iv = (4, 5) if problem.model == 0 else (1, 2)
    i_n, i_a = N_INPUTS[problem.model] - 3, N_INPUTS[problem.model] - 2          # the last two inputs once the time is removed
    nd = problem.track_nodes if problem.closed else problem.track_nodes[1:]
    kz = -np.sin(nd[:, 2]) * nd[:, 4] + np.cos(nd[:, 1]) * np.cos(nd[:, 2]) * nd[:, 3]
    u, v, n, a = X[:, iu], X[:, iv], X[:, i_n], X[:, i_a]
    un, r = u * np.cos(a) - v * np.sin(a), 1.0 - kz * n
    d = (-kz * DX[:, i_n] / un - np.cos(a) * r / un ** 2 * DX[:, iu] + np.sin(a) * r / un ** 2 * DX[:, iv]
         + r * (v * np.cos(a) + u * np.sin(a)) / un ** 2 * DX[:, i_a])
    sg = problem.sigma
    if problem.closed:
        h = np.diff(np.append(problem.s, problem.track_length))
        return float(np.sum(h * ((1.0 - sg) * d + sg * np.roll(d, -1)))), d
    d = np.concatenate([[0.0], d])
    h = np.diff(problem.s)
    return float(np.sum(h * ((1.0 - sg) * d[:-1] + sg * d[1:]))), d


def solve_lap(problem, x_start, options=None, device=0, bounds=None, coarse_nodes=500, sequence_from=3000, verbose=False):
    """One closed lap.  Meshes of `sequence_from` nodes or more are solved by mesh sequencing: the lap on every k-th node (about
    `coarse_nodes` of them) from `x_start`, then the full mesh from that solution interpolated to it; coarser meshes directly,
    with the sequence as the second attempt.  (Measured on the F1 Catalunya lap, tools/mesh_study.py: from the steady-state
    start the 8000-node lap ends in a failed line search or -- 4000 nodes -- in a slower local optimum; from the 500-node lap
    both reach the optimum the mesh study converges to.  The solver has no restoration phase.)
    x_start: one node's variables, replicated.  Returns (LapSolver.solve's dict, list of the levels solved)."""
    from . import workloads
    node = np.asarray(x_start, dtype=np.float64).ravel()[:problem.nv]
    levels = []

    def run(p, x0, bnd):
        nlp = CollocationNLP(p, device=device)
        solver = LapSolver(nlp, options=options, bounds=bnd)
        out = solver.solve(x0, verbose=verbose)
        solver.close(); nlp.close()
        levels.append(dict(nodes=p.n_points, status=int(out["status"][0]), iterations=int(out["iters"][0])))
        return out

    def sequence():
        every = max(2, int(round(problem.n_points / float(coarse_nodes))))
        pc = problem.coarsened(every)
        bc = None if bounds is None else tuple(np.asarray(b).reshape(problem.n_points, -1)[::every].ravel() for b in bounds)
        oc = run(pc, np.tile(node, pc.n_points), bc)
        if oc["status"][0] < 1:
            return oc
        return run(problem, workloads.resample_point(problem, pc.s, oc["x"][0], problem.nv), bounds)

    if not problem.closed or problem.n_points < 4 * coarse_nodes // 2:
        return run(problem, np.tile(node, problem.n_points if problem.closed else problem.n_points - 1), bounds), levels
    if problem.n_points >= sequence_from:
        return sequence(), levels
    out = run(problem, np.tile(node, problem.n_points), bounds)
    if out["status"][0] < 1:
        out = sequence()
    return out, levels


def sweep_bounds(problem, params):
    """Bounds of a vehicle-parameter sweep.  The tyre-slip rows of the F1 are bounded by +-max(lambda_max(0), lambda_max(m g0)) of each
    lap's OWN vehicle (limebeer2014f1.h:232-262): when those differ between the laps of the batch the solver gets one set of bounds
    per lap; when they do not (the usual case: lambda_max(0) dominates and does not depend on the mass), None = one set for the batch."""
    from .nlp import problem_bounds
    ncpe = problem.ncpe
    if problem.model != 0:
        return None                                   # lot2016kart.h:194-198: the kart's slip rows are bounded by constants
    # the F1's four slip rows (csrc/fl_capi.cu fl_problem_bounds, limebeer2014f1.h:232-262): the parameters they read are the mass and,
    # per axle, the two reference loads and the two maximum slip angles of the tyre
    params = np.asarray(params, dtype=np.float64)
    deg, g0 = np.pi / 180.0, 9.81
    m = params[:, 0]
    eu = np.empty((params.shape[0], 4))
    for t in range(4):
        o = 32 if t < 2 else 45
        l1, l2, fz1, fz2 = params[:, o + 9] * deg, params[:, o + 10] * deg, params[:, o + 1], params[:, o + 2]
        a = (0.0 - fz1) * (l2 - l1) / (fz2 - fz1) + l1
        b = (m * g0 - fz1) * (l2 - l1) / (fz2 - fz1) + l1
        eu[:, t] = np.maximum(a, b)
    if np.all(np.abs(eu - eu[0]) <= 1e-15):
        return None
    n_eq = ncpe - 6 - (0 if problem.form == 0 else 2)
    xl, xu, gl, gu = problem_bounds(problem)
    B = params.shape[0]
    GL, GU = np.tile(gl, (B, 1)).reshape(B, problem.n_elements, ncpe), np.tile(gu, (B, 1)).reshape(B, problem.n_elements, ncpe)
    for t in range(4):
        GL[:, :, n_eq + t], GU[:, :, n_eq + t] = -eu[:, t, None], eu[:, t, None]
    return np.tile(xl, (B, 1)), np.tile(xu, (B, 1)), GL.reshape(B, -1), GU.reshape(B, -1)


def solve_sweep(problem, vehicle_params, options=None, device=0, start_speeds_kmh=(50.0, 100.0, 150.0), retry_mu_init=1.0, retry_max_iter=600):
    """Vehicle-parameter sweep: one lap per row of `vehicle_params`, all from the reference's starting point (steady state
    at start_speeds_kmh[0] replicated over the mesh, fastestlapc.cpp:1508-1545).  The interior-point method has no
    restoration phase: the few laps that do not converge are solved again in a small second (third) batch from another
    starting speed and a larger initial barrier parameter -- the retry-from-another-start loop the reference itself uses
    around Ipopt where it fails (steady_state.hpp:200-226, 786-800).  Returns LapSolver.solve's dict, plus `passes`."""
    import time
    from . import workloads
    t0 = time.perf_counter()
    params = np.ascontiguousarray(vehicle_params, dtype=np.float64)
    nlp = CollocationNLP(problem, vehicle_params=params, device=device)
    solver = LapSolver(nlp, options=options, bounds=sweep_bounds(problem, params))
    t1 = time.perf_counter()
    x_start = workloads.f1_start_point(problem, start_speeds_kmh[0], device=device)
    t2 = time.perf_counter()
    out = solver.solve(x_start)
    t3 = time.perf_counter()
    solver.close(); nlp.close()
    passes = [dict(laps=int(params.shape[0]), converged=int((out["status"] >= 1).sum()), device_ms=out["device_ms"])]
    phases = dict(allocate=t1 - t0, start_point=t2 - t1, first_pass=t3 - t2, retries=0.0)
    t4 = time.perf_counter()
    for speed in start_speeds_kmh[1:]:
        bad = np.where(out["status"] < 1)[0]
        if bad.size == 0:
            break
        o2 = default_options() if options is None else options
        keep = (o2.mu_init, o2.max_iter)
        o2.mu_init, o2.max_iter = retry_mu_init, max(retry_max_iter, o2.max_iter)
        nlp2 = CollocationNLP(problem, vehicle_params=params[bad], device=device)
        s2 = LapSolver(nlp2, options=o2)
        r = s2.solve(workloads.f1_start_point(problem, speed, device=device))
        s2.close(); nlp2.close()
        o2.mu_init, o2.max_iter = keep
        good = r["status"] >= 1
        for k in ("x", "y", "zl", "zu", "obj", "status", "error", "constr_viol", "dual_inf", "mu"):
            out[k][bad[good]] = r[k][good]
        out["iters"][bad[good]] += r["iters"][good]
        for k in out["stats"]:
            out["stats"][k] += r["stats"][k]
        passes.append(dict(laps=int(bad.size), converged=int(good.sum()), device_ms=r["device_ms"], start_speed_kmh=speed, mu_init=retry_mu_init))
    out["passes"] = passes
    phases["retries"] = time.perf_counter() - t4
    out["phases_s"] = phases
    return out
