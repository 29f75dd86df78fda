"""TEST INFRASTRUCTURE ONLY -- CPU statement of the interior-point iteration the device solver runs (fastest_lap_b200/csrc/ipm_kernels.cuh).

The reference hands its collocation NLP to Ipopt (lion::ipopt_cppad_solve, call site
/root/reference/src/core/applications/optimal_laptime.hpp:546-551; options :528-541: tol 1e-10, constr_viol_tol 1e-10,
acceptable_tol 1e-8, max_iter 3000, everything else Ipopt's default).  Ipopt is not in this container, so this file
restates the published algorithm (Waechter & Biegler, "On the implementation of an interior-point filter line-search
algorithm for large-scale nonlinear programming", Math. Program. 106, 2006 -- the equation numbers below are that
paper's) with numpy / scipy.sparse: slack reformulation of the inequality rows, monotone barrier update (7), fraction
to the boundary (15), filter line search (18)-(20) with second-order correction, primal-dual Hessian damping reset (16),
and inertia-correcting regularisation of the augmented system (Algorithm IC).  With a `layout` the augmented system goes
through oracle/kkt.py (the statement of the device's block factorisation, exact inertia from its pivot signs, Jacobian
regularisation always on); without one, through scipy's sparse LU with the inertia-free curvature test of Chiang &
Zavala (Comput. Optim. Appl. 64, 2016).  Deviations from Ipopt, shared with the device solver: no NLP scaling, no
restoration phase (here the step is refused and the status says so; the device solver restarts the barrier problem).

Parity: this solver is pinned on a converged lap of the reference in tests/test_oracle_kkt_ipm.py; the device solver is
pinned on the reference's laps directly (tests/test_gpu_ipm.py).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import kkt


class Options:
    tol = 1.0e-10
    constr_viol_tol = 1.0e-10
    acceptable_tol = 1.0e-8
    acceptable_iter = 15
    max_iter = 3000
    mu_init = 0.1
    bound_push = 1.0e-2          # kappa_1
    bound_frac = 1.0e-2          # kappa_2
    bound_relax_factor = 1.0e-8
    kappa_eps = 10.0
    kappa_mu = 0.2
    theta_mu = 1.5
    tau_min = 0.99
    kappa_sigma = 1.0e10
    s_max = 100.0
    # filter line search
    gamma_theta = 1.0e-5
    gamma_phi = 1.0e-8
    delta = 1.0
    s_theta = 1.1
    s_phi = 2.3
    eta_phi = 1.0e-8
    max_soc = 4
    kappa_soc = 0.99
    alpha_red = 0.5
    max_backtracks = 40
    # regularisation
    delta_w_min = 1.0e-20
    delta_w_0 = 1.0e-4
    delta_w_max = 1.0e40
    kappa_w_minus = 1.0 / 3.0
    kappa_w_plus = 8.0
    kappa_w_plus_bar = 100.0
    delta_c_bar = 1.0e-8
    kappa_c = 0.25
    curv_tol = 1.0e-11
    refine_steps = 2
    y_init_max = 1.0e3


class Result:
    pass


def solve(evaluate, x0, xl, xu, gl, gu, opts=None, y0=None, log=None, kkt_solver=None, layout=None):
    """evaluate(x, y, obj_factor, derivs) -> dict(f, g[, grad, J (csr m x n), H (csr n x n symmetric)])."""
    o = opts or Options()
    n, m = len(x0), len(gl)
    ineq = np.where(gu > gl)[0]
    eq = np.where(gu <= gl)[0]
    ni = len(ineq)
    is_ineq = np.zeros(m, bool); is_ineq[ineq] = True
    # relaxed bounds (Ipopt bound_relax_factor)
    rel = lambda b: o.bound_relax_factor * np.maximum(1.0, np.abs(b))
    xl, xu = xl - rel(xl), xu + rel(xu)
    sl, su = gl[ineq] - rel(gl[ineq]), gu[ineq] + rel(gu[ineq])

    def push(v, lo, hi):
        pl = np.minimum(o.bound_push * np.maximum(1.0, np.abs(lo)), o.bound_frac * (hi - lo))
        pu = np.minimum(o.bound_push * np.maximum(1.0, np.abs(hi)), o.bound_frac * (hi - lo))
        return np.minimum(np.maximum(v, lo + pl), hi - pu)

    x = push(np.array(x0, dtype=float), xl, xu)
    ev = evaluate(x, np.zeros(m), 1.0, True)
    s = push(ev["g"][ineq], sl, su)
    zl, zu = np.ones(n), np.ones(n)
    vl, vu = np.ones(ni), np.ones(ni)
    mu = o.mu_init
    y = np.zeros(m) if y0 is None else np.array(y0, dtype=float)

    def kkt_solve(H, J, sig_x, sig_s, dw, dc, r1, r2):
        """[H + sig_x + dw, J^T; J, -D] [dx; dy] = [r1; r2],  D = dc (equality rows), 1/(sig_s + dw) + dc (inequality rows)."""
        D = np.full(m, dc)
        D[ineq] += 1.0 / (sig_s + dw)
        if kkt_solver is not None:
            return kkt_solver(H, J, sig_x + dw, D, r1, r2)
        K = sp.bmat([[H + sp.diags(sig_x + dw), J.T], [J, -sp.diags(D)]], format="csc")
        lu = spla.splu(K)
        rhs = np.concatenate([r1, r2])
        sol = lu.solve(rhs)
        for _ in range(o.refine_steps):
            sol += lu.solve(rhs - K @ sol)
        return sol[:n], sol[n:]

    if y0 is None:
        # least-squares multipliers (Ipopt's default initialisation): min |grad f + J^T y - zl + zu|^2 + |-y_ineq - vl + vu|^2,
        # i.e. [I J^T; J -D][w; y] = -[grad f - zl + zu; 0] with D = 1 on the inequality rows (their slack columns)
        try:
            if layout is not None:
                Dv = np.full(m, o.delta_c_bar * mu ** o.kappa_c)
                Dv[ineq] = 1.0
                Dg, Lo = kkt.assemble_blocks(layout, sp.csr_matrix((n, n)), ev["J"], np.ones(n), Dv)
                F = kkt.BlockCyclicLDL(Dg, Lo, layout.closed)
                r1_, r2_ = -(ev["grad"] - zl + zu), np.zeros(m)
                _, yls = kkt.expand_solution(layout, ev["J"], Dv, r2_, F.solve(kkt.condense_rhs(layout, ev["J"], Dv, r1_, r2_)))
            else:
                _, yls = kkt_solve(sp.csr_matrix((n, n)), ev["J"], np.ones(n), np.ones(ni), 0.0, 0.0, -(ev["grad"] - zl + zu), np.zeros(m))
            y = yls if np.all(np.isfinite(yls)) and np.max(np.abs(yls)) <= o.y_init_max else np.zeros(m)
        except RuntimeError:
            y = np.zeros(m)

    def resid(x, s, g):
        c = g.copy()
        c[eq] -= gl[eq]
        c[ineq] -= s
        return c

    def barrier(f, x, s):
        return f - mu * (np.sum(np.log(x - xl)) + np.sum(np.log(xu - x)) + np.sum(np.log(s - sl)) + np.sum(np.log(su - s)))

    def error(mu_, ev, x, s, y, zl, zu, vl, vu):
        rx = ev["grad"] + ev["J"].T @ y - zl + zu
        rs = -y[ineq] - vl + vu
        c = resid(x, s, ev["g"])
        comp = max(np.max(np.abs((x - xl) * zl - mu_)), np.max(np.abs((xu - x) * zu - mu_)),
                   np.max(np.abs((s - sl) * vl - mu_)) if ni else 0.0, np.max(np.abs((su - s) * vu - mu_)) if ni else 0.0)
        nz = 2 * n + 2 * ni
        sd = max(o.s_max, (np.sum(np.abs(y)) + np.sum(zl) + np.sum(zu) + np.sum(vl) + np.sum(vu)) / (m + nz)) / o.s_max
        sc = max(o.s_max, (np.sum(zl) + np.sum(zu) + np.sum(vl) + np.sum(vu)) / nz) / o.s_max
        dual = max(np.max(np.abs(rx)), np.max(np.abs(rs)) if ni else 0.0)
        return max(dual / sd, np.max(np.abs(c)), comp / sc), dual, np.max(np.abs(c)), comp

    filt = []
    theta_max = theta_min = None
    dw_last = 0.0
    res = Result()
    res.status, res.iters, res.n_reg, res.n_fact, res.history = "max_iter", 0, 0, 0, []
    acceptable_count = 0

    for it in range(o.max_iter + 1):
        ev = evaluate(x, y, 1.0, True)
        f, g, grad, J, H = ev["f"], ev["g"], ev["grad"], ev["J"], ev["H"]
        e0, dual, cviol, comp = error(0.0, ev, x, s, y, zl, zu, vl, vu)
        res.history.append((it, f, cviol, dual, mu))
        if log:
            log("%4d  f %.10f  inf_pr %.2e  inf_du %.2e  compl %.2e  mu %.1e  dw %.1e" % (it, f, cviol, dual, comp, mu, dw_last))
        if e0 <= o.tol and cviol <= o.constr_viol_tol:
            res.status = "success"
            break
        acceptable_count = acceptable_count + 1 if e0 <= o.acceptable_tol else 0
        if acceptable_count >= o.acceptable_iter:
            res.status = "acceptable"
            break
        if it == o.max_iter:
            break
        # barrier parameter update (7); several reductions may follow each other
        changed = False
        while mu > o.tol / 10.0 and error(mu, ev, x, s, y, zl, zu, vl, vu)[0] <= o.kappa_eps * mu:
            mu = max(o.tol / 10.0, min(o.kappa_mu * mu, mu ** o.theta_mu))
            changed = True
        if changed:
            filt = []
        tau = max(o.tau_min, 1.0 - mu)

        c = resid(x, s, g)
        theta = np.sum(np.abs(c))
        if theta_max is None:
            theta_max, theta_min = 1.0e4 * max(1.0, theta), 1.0e-4 * max(1.0, theta)
        phi = barrier(f, x, s)
        sig_x = zl / (x - xl) + zu / (xu - x)
        sig_s = vl / (s - sl) + vu / (su - s)
        gphi_x = grad - mu / (x - xl) + mu / (xu - x)
        gphi_s = -mu / (s - sl) + mu / (su - s)
        rs_bar = gphi_s - y[ineq]
        r1 = -(gphi_x + J.T @ y)

        def direction(dw, dc, cvec):
            r2 = -cvec.copy()
            r2[ineq] -= rs_bar / (sig_s + dw)
            dx, dy = kkt_solve(H, J, sig_x, sig_s, dw, dc, r1, r2)
            ds = (dy[ineq] - rs_bar) / (sig_s + dw)
            return dx, ds, dy

        # regularisation
        if layout is not None:
            # Algorithm IC of the paper with the exact inertia of the block factorisation
            def factor(dw_, dc_):
                Dv = np.full(m, dc_)
                Dv[ineq] += 1.0 / (sig_s + dw_)
                Dg, Lo = kkt.assemble_blocks(layout, H, J, sig_x + dw_, Dv)
                F = kkt.BlockCyclicLDL(Dg, Lo, layout.closed)
                return F, Dv
            n_pos, n_neg = layout.n, layout.n_stages * layout.neq
            # closed laps with an even number of nodes and massless wheels have structurally dependent rows (the wheel rows
            # of consecutive elements read f_i + f_j = 0: their alternating sum vanishes), so the Jacobian regularisation
            # of IC-2 is always on -- what Ipopt's degeneracy heuristic arrives at after its first iterations
            dw, dc = 0.0, o.delta_c_bar * mu ** o.kappa_c
            F, Dv = factor(dw, dc)
            res.n_fact += 1
            while not (F.ok and F.inertia == (n_pos, n_neg, 0)):
                res.n_reg += 1
                if dw == 0.0:
                    dw = o.delta_w_0 if dw_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * dw_last)
                else:
                    dw *= o.kappa_w_plus_bar if dw_last == 0.0 else o.kappa_w_plus
                if dw > o.delta_w_max:
                    res.status = "regularisation_failed"
                    break
                F, Dv = factor(dw, dc)
                res.n_fact += 1
            if res.status == "regularisation_failed":
                break

            def direction(dw_, dc_, cvec):
                r2 = -cvec.copy()
                r2[ineq] -= rs_bar / (sig_s + dw_)
                b = kkt.condense_rhs(layout, J, Dv, r1, r2)
                z = F.solve(b)
                dx_, dy_ = kkt.expand_solution(layout, J, Dv, r2, z)
                for _ in range(o.refine_steps):
                    # iterative refinement on the full augmented system
                    e1 = r1 - (H @ dx_ + (sig_x + dw_) * dx_ + J.T @ dy_)
                    e2 = r2 - (J @ dx_ - Dv * dy_)
                    z = F.solve(kkt.condense_rhs(layout, J, Dv, e1, e2))
                    ddx, ddy = kkt.expand_solution(layout, J, Dv, e2, z)
                    dx_, dy_ = dx_ + ddx, dy_ + ddy
                ds_ = (dy_[ineq] - rs_bar) / (sig_s + dw_)
                return dx_, ds_, dy_
            dx, ds, dy = direction(dw, dc, c)
        else:
            # inertia-free curvature test along the computed direction (Chiang & Zavala)
            dc = o.delta_c_bar * mu ** o.kappa_c
            dw = 0.0
            while True:
                dx, ds, dy = direction(dw, dc, c)
                Hd = H @ dx
                curv = dx @ Hd + dx @ ((sig_x + dw) * dx) + ds @ ((sig_s + dw) * ds)
                if np.all(np.isfinite(dx)) and curv >= o.curv_tol * (dx @ dx + ds @ ds):
                    break
                res.n_reg += 1
                if dw == 0.0:
                    dw = o.delta_w_0 if dw_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * dw_last)
                else:
                    dw *= o.kappa_w_plus_bar if dw_last == 0.0 else o.kappa_w_plus
                if dw > o.delta_w_max:
                    res.status = "regularisation_failed"
                    break
            if res.status == "regularisation_failed":
                break
        if dw > 0.0:
            dw_last = dw

        def max_step(v, dv, lo, hi, tau_):
            a = 1.0
            neg = dv < 0
            if np.any(neg):
                a = min(a, np.min(-tau_ * (v[neg] - lo[neg]) / dv[neg]))
            pos = dv > 0
            if np.any(pos):
                a = min(a, np.min(tau_ * (hi[pos] - v[pos]) / dv[pos]))
            return a

        a_max = min(max_step(x, dx, xl, xu, tau), max_step(s, ds, sl, su, tau) if ni else 1.0)
        dzl = mu / (x - xl) - zl - zl / (x - xl) * dx
        dzu = mu / (xu - x) - zu + zu / (xu - x) * dx
        dvl = mu / (s - sl) - vl - vl / (s - sl) * ds
        dvu = mu / (su - s) - vu + vu / (su - s) * ds
        big = np.full(1, np.inf)

        def dual_step(z, dz):
            neg = dz < 0
            return min(1.0, np.min(-tau * z[neg] / dz[neg])) if np.any(neg) else 1.0

        a_dual = min(dual_step(zl, dzl), dual_step(zu, dzu), dual_step(vl, dvl) if ni else 1.0, dual_step(vu, dvu) if ni else 1.0)

        dphi = gphi_x @ dx + gphi_s @ ds
        a_min = 1.0e-13
        alpha = a_max
        accepted = False
        dx_c, ds_c = dx, ds
        for bt in range(o.max_backtracks):
            xt, st = x + alpha * dx_c, s + alpha * ds_c
            evt = evaluate(xt, None, 1.0, False)
            ct = resid(xt, st, evt["g"])
            theta_t = np.sum(np.abs(ct))
            phi_t = barrier(evt["f"], xt, st)
            ok_filter = theta_t <= theta_max and all(theta_t < (1.0 - o.gamma_theta) * th or phi_t < ph - o.gamma_phi * th for th, ph in filt) \
                and np.isfinite(phi_t)
            switching = dphi < 0 and alpha * (-dphi) ** o.s_phi > o.delta * theta ** o.s_theta and theta <= theta_min
            if ok_filter:
                if switching:
                    if phi_t <= phi + o.eta_phi * alpha * dphi:
                        accepted = True
                elif theta_t <= (1.0 - o.gamma_theta) * theta or phi_t <= phi - o.gamma_phi * theta:
                    accepted = True
            if accepted:
                break
            # second-order correction on the first trial step when infeasibility did not decrease
            if bt == 0 and theta_t >= theta:
                c_soc, theta_old = alpha * c + ct, theta_t
                for p_ in range(o.max_soc):
                    dxs, dss, dys = direction(dw, dc, c_soc)
                    a_s = min(max_step(x, dxs, xl, xu, tau), max_step(s, dss, sl, su, tau) if ni else 1.0)
                    xs_, ss_ = x + a_s * dxs, s + a_s * dss
                    evs = evaluate(xs_, None, 1.0, False)
                    cs_ = resid(xs_, ss_, evs["g"])
                    th_s, ph_s = np.sum(np.abs(cs_)), barrier(evs["f"], xs_, ss_)
                    okf = th_s <= theta_max and all(th_s < (1.0 - o.gamma_theta) * th or ph_s < ph - o.gamma_phi * th for th, ph in filt) and np.isfinite(ph_s)
                    acc = False
                    if okf:
                        if switching:
                            acc = ph_s <= phi + o.eta_phi * a_max * dphi
                        else:
                            acc = th_s <= (1.0 - o.gamma_theta) * theta or ph_s <= phi - o.gamma_phi * theta
                    if acc:
                        accepted, xt, st, alpha, dy = True, xs_, ss_, a_s, dys
                        theta_t, phi_t = th_s, ph_s
                        break
                    if th_s > o.kappa_soc * theta_old:
                        break
                    c_soc, theta_old = a_s * c_soc + cs_, th_s
                if accepted:
                    break
            alpha *= o.alpha_red
            if alpha < a_min:
                break
        if not accepted:
            res.status = "line_search_failed"
            break
        if not switching or not (phi_t <= phi + o.eta_phi * alpha * dphi):
            filt.append(((1.0 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
        x, s = xt, st
        y = y + alpha * dy
        zl, zu, vl, vu = zl + a_dual * dzl, zu + a_dual * dzu, vl + a_dual * dvl, vu + a_dual * dvu
        # (16): keep the bound multipliers within kappa_sigma of the primal estimate mu / slack
        def reset(z, slack):
            return np.maximum(np.minimum(z, o.kappa_sigma * mu / slack), mu / (o.kappa_sigma * slack))
        zl, zu, vl, vu = reset(zl, x - xl), reset(zu, xu - x), reset(vl, s - sl), reset(vu, su - s)
        res.iters = it + 1

    res.x, res.s, res.y, res.zl, res.zu, res.vl, res.vu, res.mu, res.f = x, s, y, zl, zu, vl, vu, mu, ev["f"]
    return res


def oracle_evaluator(onlp):
    """evaluate() closure over an oracle.OracleNLP."""
    n, m = onlp.n, onlp.m

    def evaluate(x, y, obj, derivs):
        r = onlp.eval(x, y, obj, derivs=derivs)
        out = {"f": r["f"], "g": r["g"]}
        if derivs:
            jr, jc, jv = r["jac"]
            hr, hc, hv = r["hess"]
            out["grad"] = r["grad"]
            out["J"] = sp.csr_matrix((jv, (jr, jc)), shape=(m, n))
            L = sp.csr_matrix((hv, (hr, hc)), shape=(n, n))
            out["H"] = L + sp.tril(L, -1).T
        return out
    return evaluate
