"""Symbolic statement of what ONE collocation node contributes to the NLP callbacks.

Restates Optimal_laptime::FG_direct / FG_derivative (src/core/applications/optimal_laptime.hpp:1226-1657) node-wise:
every term of f, g and the Lagrangian that involves node j depends on x_j only (plus, linearly, on the neighbours'
controls through the control-rate penalty), so eval_grad_f / eval_jac_g / eval_h split into independent per-node pieces:

  A-block  rows of the element that ENDS at node j   (j is the "right" node: states, sigma-weighted rates, algebraic rows, extras)
  B-block  rows of the element that STARTS at node j (j is the "left" node: trapezoid rows only)
  H-block  lower triangle of the Hessian of phi_j(x_j) = obj*f_j + sum(lambda * rows touching x_j)
  C-slots  control-control couplings (u_j, u_{j-1}) of the direct form's control-rate penalty.

The variable order inside a node is the reference's (optimal_laptime.hpp:1269-1285): inputs without `time`, then the
full-mesh controls, then (derivative form) their rates.
"""
import os

from .expr import Graph, E
from . import models


class Variant:
    def __init__(self, model, form, elevation=False, direct_torque=False, aug=False):
        assert model in ("f1", "kart") and form in ("direct", "derivative", "steady")
        self.model, self.form, self.elevation, self.direct_torque = model, form, bool(elevation), bool(direct_torque)
        # one augmented state per node (see NodeProgram): a control held constant over stretches of the lap (hypermesh / constant
        # controls) or the running value of a restricted integral quantity
        self.aug = bool(aug)
        if aug:
            assert model == "f1" and form == "direct", "the augmented state is offered for the F1, direct form"
        if model == "kart":
            assert not elevation, "lot2016kart rejects tracks with elevation (lot2016kart.h:63-68)"

    @property
    def name(self):
        if self.form == "steady":
            return "%s_steady" % self.model
        n = "%s_%s" % (self.model, self.form)
        if self.elevation:
            n += "_3d"
        if self.direct_torque:
            n += "_torque"
        if self.aug:
            n += "_aug"
        return n


class NodeProgram:
    """All symbols and output expressions of one variant."""

    def __init__(self, variant):
        v = self.variant = variant
        g = self.g = Graph()
        S = lambda name, cls: E(g, g.sym(name, cls))
        if v.form == "steady":
            (self._init_steady if v.model == "kart" else self._init_steady_f1)(S)
            return
        deriv = v.form == "derivative"

        if v.model == "f1":
            self.param_names = models.F1_PARAMS + models.F1_HOST_PARAMS
            n_in, time_idx = 14, 11
            self.n_fm = 2
            td_ids = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13] if v.elevation else [0, 1, 2, 3, 4, 5, 6, 12, 13]
            alg_ids = [] if v.elevation else [7, 8, 9, 10]       # limebeer2014f1.h:185-225
        else:
            self.param_names = models.KART_PARAMS + models.KART_HOST_PARAMS
            n_in, time_idx = 13, 10
            self.n_fm = 2
            td_ids = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12]
            alg_ids = []
        self.param_names = self.param_names + ["diss0", "diss1"]
        # Augmented state (Variant.aug).  The reference keeps a hypermesh / constant control as ONE variable per stretch, appended to
        # x (optimal_laptime.hpp:378-404, control_array_at_s: detail/optimal_laptime_control_variables.h:190-222), and a restricted
        # integral quantity as ONE dense constraint row (optimal_laptime.hpp:1316-1317, 1352-1353, 1409-1416): both couple every
        # node of the lap.  Here they are restated node-wise so that the KKT system stays block-tridiagonal: every node carries
        #   a  the control's value at the node / the integral accumulated up to the node,
        #   q  a free variable that enters the row of the element ENDING at the node with the per-node constant rj (0 or 1),
        # and every element one more trapezoid row  a_j + rj q_j - bout_i a_i - (awout_i I_i + awin_j I_j) = 0, I = integrand dt/ds.
        # Control: I = 0, bout = 1, rj = 1 where a new stretch starts (free jump), 0 inside a stretch (a_j = a_i).  Integral on a
        # closed lap: rj = 0, bout = 0 at node 0 only -- the accumulation starts from zero there and the closing element makes a_0 the
        # value of the whole integral, restricted by a_0's own bounds [lower, upper]; q is inert.
        # Same feasible set and objective as the reference's formulation, hence the same optimum.
        self.n_aug = 1 if v.aug else 0
        if v.aug:
            self.param_names = self.param_names + ["aug_bb", "aug_w_engine", "aug_w_tfl", "aug_w_tfr", "aug_w_trl", "aug_w_trr"]
        self.td_ids, self.alg_ids = td_ids, alg_ids
        self.n_td, self.n_alg, self.n_ext = len(td_ids) + self.n_aug, len(alg_ids), 6
        self.n_ctrl_rows = self.n_fm if deriv else 0
        self.ncpe = self.n_td + self.n_alg + self.n_ext + self.n_ctrl_rows
        self.nv = (n_in - 1) + 2 * self.n_aug + self.n_fm + (self.n_fm if deriv else 0)
        # the full-mesh controls stay the LAST variables of the node (the warp-per-lap KKT schedule relies on it): a, q sit between
        # the inputs and the controls
        self.aug_pos = [n_in - 1, n_in] if v.aug else []
        self.ctrl_pos = [n_in - 1 + 2 * self.n_aug, n_in + 2 * self.n_aug]
        self.dctrl_pos = [n_in + 1, n_in + 2] if deriv else []

        # ---- symbols
        self.x = [S("x%d" % k, "x") for k in range(self.nv)]
        P = self.P = {n: S("p_" + n, "p") for n in self.param_names}
        # (wt, wn, wb: wind along the track's tangent, normal, binormal -- per-node constants, chassis.hpp:269-274)
        if v.elevation:
            self.track_names = ["kx", "ky", "kz", "smu", "cmu", "sph", "cph", "wt", "wn", "wb"]
            trk = {n: S("t_" + n, "t") for n in self.track_names}
        elif os.environ.get("FL_NO_WIND"):          # experiment: the node programs without the wind terms
            self.track_names = ["kz"]
            trk = dict(kx=0.0, ky=0.0, kz=S("t_kz", "t"), smu=0.0, cmu=1.0, sph=0.0, cph=1.0, wt=0.0, wn=0.0, wb=0.0)
        else:
            self.track_names = ["kz", "wt", "wn"]
            trk = dict(kx=0.0, ky=0.0, kz=S("t_kz", "t"), smu=0.0, cmu=1.0, sph=0.0, cph=1.0, wt=S("t_wt", "t"), wn=S("t_wn", "t"), wb=0.0)
        if v.aug:
            self.track_names = self.track_names + ["rj", "bout", "awin", "awout"]
            rj, bout, awin, awout = S("t_rj", "t"), S("t_bout", "t"), S("t_awin", "t"), S("t_awout", "t")
        # per-node mesh constants: lengths of the element that ends / starts at the node; direct form: their inverses
        # (0 where the element does not exist); derivative form: total length weighting the node's control-rate penalty
        self.geom_names = ["hin", "hout"] + (["penw"] if deriv else ["ihin", "ihout"])
        hin, hout = S("t_hin", "t"), S("t_hout", "t")
        penw = S("t_penw", "t") if deriv else None
        ihin = None if deriv else S("t_ihin", "t")
        ihout = None if deriv else S("t_ihout", "t")
        self.lam_in = [S("lin%d" % r, "w") for r in range(self.ncpe)]
        self.lam_out = [S("lout%d" % r, "w") for r in range(self.n_td)] + [None] * (self.n_alg + self.n_ext) + \
                       [S("lout%d" % (self.n_td + self.n_alg + self.n_ext + c), "w") for c in range(self.n_ctrl_rows)]
        obj = S("obj", "w")
        self.uprev = [S("uprev%d" % c, "n") for c in range(self.n_fm)] if not deriv else []
        self.unext = [S("unext%d" % c, "n") for c in range(self.n_fm)] if not deriv else []
        sigma = P["sigma"]
        diss = [P["diss0"], P["diss1"]]

        # ---- the vehicle at this node
        q = list(self.x[:time_idx]) + [E(g, g.const(0.0))] + list(self.x[time_idx:n_in - 1])
        ctrl = [self.x[p] for p in self.ctrl_pos]
        if v.model == "f1":
            # controls: delta, boost, throttle, brake-bias; boost and brake-bias keep their default values
            # (Control_variables::control_array_at_s, detail/optimal_laptime_control_variables.h:190-222)
            bias = P["brake_bias"]
            if v.aug:
                a_var, q_var = self.x[self.aug_pos[0]], self.x[self.aug_pos[1]]
                bias = P["aug_bb"] * a_var + (1.0 - P["aug_bb"]) * bias
            u = [ctrl[0], E(g, g.const(0.0)), ctrl[1], bias]
            out = models.f1_model(q, u, P, trk)
        else:
            out = models.kart_model(q, ctrl, P, trk, direct_torque=v.direct_torque, curvilinear=True)
        states, dstates, dtds, extra = out["states"], out["dstates"], out["dtds"], out["extra"]
        self.outputs = out["outputs"]          # named output channels of the model at this node (get_outputs_map of the reference)

        Q = [states[i] for i in td_ids]
        QD = [dstates[i] for i in td_ids]
        QA = []
        if v.aug:
            # integrands of limebeer2014f1::compute_integral_quantities (limebeer2014f1.h:291-306) times dt/ds (optimal_laptime.hpp:1352)
            iq = out["integrands"]
            integrand = P["aug_w_engine"] * iq[0] + P["aug_w_tfl"] * iq[1] + P["aug_w_tfr"] * iq[2] + P["aug_w_trl"] * iq[3] + P["aug_w_trr"] * iq[4]
            Q.append(bout * a_var)
            QD.append(integrand * dtds)
            QA.append(a_var + rj * q_var)
        ALG = [dstates[i] / dtds for i in alg_ids]              # optimal_laptime.hpp:1341 (= the un-scaled equation)
        EXT = list(extra)
        CDOT = [self.x[p] * dtds for p in self.dctrl_pos]       # optimal_laptime.hpp:1580-1581
        self.Q, self.QD, self.ALG, self.EXT, self.CDOT, self.DTDS = Q, QD, ALG, EXT, CDOT, dtds

        # ---- node values handed to the element assembly (g, f)
        self.values = Q + QD + ALG + EXT + [dtds] + CDOT + QA
        self.value_names = (["Q%d" % r for r in range(self.n_td)] + ["QD%d" % r for r in range(self.n_td)] +
                            ["ALG%d" % r for r in range(self.n_alg)] + ["EXT%d" % r for r in range(self.n_ext)] + ["DTDS"] +
                            ["CDOT%d" % c for c in range(self.n_ctrl_rows)] + ["QA%d" % r for r in range(self.n_aug)])

        # ---- rows as functions of x_j: A (j is the right node of its element), B (j is the left node)
        one_m_sigma = 1.0 - sigma
        rowsA, rowsB = [], []
        for r in range(self.n_td - self.n_aug):
            rowsA.append(Q[r] - hin * sigma * QD[r])
            rowsB.append(-Q[r] - hout * one_m_sigma * QD[r])
        for r in range(self.n_aug):
            # weights of the node's integrand in the two elements it belongs to: per-node constants (the closing element of the
            # reference swaps them, optimal_laptime.hpp:1409)
            k = self.n_td - self.n_aug + r
            rowsA.append(QA[r] - awin * QD[k])
            rowsB.append(-Q[k] - awout * QD[k])
        for r in range(self.n_alg):
            rowsA.append(ALG[r]); rowsB.append(None)
        for r in range(self.n_ext):
            rowsA.append(EXT[r]); rowsB.append(None)
        for c in range(self.n_ctrl_rows):
            rowsA.append(ctrl[c] - hin * sigma * CDOT[c])
            rowsB.append(-ctrl[c] - hout * one_m_sigma * CDOT[c])
        self.rowsA, self.rowsB = rowsA, rowsB

        # ---- local objective: every term of f that involves x_j (its x_j-gradient is d f / d x_j)
        fobj = (hin * sigma + hout * one_m_sigma) * dtds
        if deriv:
            for c in range(self.n_fm):
                du = self.x[self.dctrl_pos[c]]
                fobj = fobj + 0.5 * diss[c] * penw * du * du                   # optimal_laptime.hpp:1553, 1611
        else:
            for c in range(self.n_fm):
                d_in, d_out = ctrl[c] - self.uprev[c], self.unext[c] - ctrl[c]
                fobj = fobj + diss[c] * (d_in * d_in * ihin + d_out * d_out * ihout)   # optimal_laptime.hpp:1362-1366, 1401-1402
        self.fobj = fobj

        # ---- Lagrangian restricted to x_j
        phi = obj * fobj
        for r in range(self.ncpe):
            phi = phi + self.lam_in[r] * rowsA[r]
            if rowsB[r] is not None:
                phi = phi + self.lam_out[r] * rowsB[r]
        self.phi = phi

        self._differentiate(rowsA, rowsB, fobj, phi, obj, diss if not deriv else None, ihin if not deriv else None)

    def _init_steady(self, S):
        """Steady-state (G-G diagram) NLP of the kart as a one-node "lap": lot2016kart::steady_state_equations
        (src/core/vehicles/lot2016kart.h:124-182) behind the functors Solve / Max_lat_acc / Max_lon_acc / Min_lon_acc of
        src/core/applications/steady_state.h:99-284, in the 8-variable layout of Max_lat_acc (steady_state.hpp:899):
        x = [w_axle, z, phi, mu, psi, delta, ax, ay].  Per-problem data travel as extra parameters: speed, given ax / ay,
        switches that make ax / ay a variable or a given value, and the objective weights (f = cax ax + cay ay).
        The 6 equations are the node's "algebraic" rows, the 6 tyre-slip rows its inequality rows; there are no trapezoid
        rows and no controls, and the element assembly's objective h * dt/ds carries f (h = 1)."""
        g = self.g
        self.param_names = models.KART_PARAMS + models.KART_HOST_PARAMS + ["diss0", "diss1"] + \
            ["ss_v", "ss_ax", "ss_ay", "ss_free_ax", "ss_free_ay", "ss_cax", "ss_cay"]
        self.n_fm, self.td_ids, self.alg_ids = 0, [], list(range(6))
        self.n_td, self.n_alg, self.n_ext, self.n_ctrl_rows = 0, 6, 6, 0
        self.ncpe, self.nv = 12, 8
        self.ctrl_pos, self.dctrl_pos = [0, 0], []
        self.x = [S("x%d" % k, "x") for k in range(self.nv)]
        P = self.P = {n: S("p_" + n, "p") for n in self.param_names}
        self.track_names = ["kz"]
        self.geom_names = ["hin", "hout", "ihin", "ihout"]
        self.lam_in = [S("lin%d" % r, "w") for r in range(self.ncpe)]
        self.lam_out = [None] * self.ncpe
        obj = S("obj", "w")
        self.uprev, self.unext = [], []
        x, v = self.x, P["ss_v"]
        ax = x[6] * P["ss_free_ax"] + (1.0 - P["ss_free_ax"]) * P["ss_ax"]
        ay = x[7] * P["ss_free_ay"] + (1.0 - P["ss_free_ay"]) * P["ss_ay"]
        psi = x[4]
        cps, sps = models.cos(psi), models.sin(psi)
        zero = E(g, g.const(0.0))
        q = [(x[0] + 1.0) * (v / 0.139), v * cps, -(v * sps), ay / v, x[1], x[2], x[3], zero, zero, zero, zero, zero, psi]
        out = models.kart_model(q, [x[5], zero], P, None, direct_torque=True, curvilinear=False)
        d = out["dstates"]
        du, dv = d[1], d[2]
        eqs = [d[7], d[3], d[8], d[9], du * sps + dv * cps, ax - du * cps + dv * sps]
        kap, lam = out["kappa"], out["lam"]
        ext = [kap[2], kap[3], lam[2], lam[3], lam[0], lam[1]]
        fobj = x[6] * P["ss_cax"] + x[7] * P["ss_cay"]
        self.Q, self.QD, self.ALG, self.EXT, self.CDOT, self.DTDS = [], [], eqs, ext, [], fobj
        self.values = eqs + ext + [fobj]
        self.value_names = ["ALG%d" % r for r in range(6)] + ["EXT%d" % r for r in range(6)] + ["DTDS"]
        rowsA, rowsB = eqs + ext, [None] * 12
        self.rowsA, self.rowsB, self.fobj = rowsA, rowsB, fobj
        phi = obj * fobj
        for r in range(self.ncpe):
            phi = phi + self.lam_in[r] * rowsA[r]
        self.phi = phi
        self._differentiate(rowsA, rowsB, fobj, phi, obj, None, None)

    def _init_steady_f1(self, S):
        """Steady-state NLP of the F1: limebeer2014f1::steady_state_equations (src/core/vehicles/limebeer2014f1.h:103-178)
        behind the same Steady_state functors; x = [kappa x4, psi / 10 deg, Fz x4 (in m g), delta / 10 deg, throttle, ax, ay]
        with ax, ay in g (acceleration_units = g0, limebeer2014f1.h:51): 11 equations, then lambda_i / lambda_max_i in [-1, 1.2]."""
        import math
        g = self.g
        self.param_names = models.F1_PARAMS + models.F1_HOST_PARAMS + ["diss0", "diss1"] + \
            ["ss_v", "ss_ax", "ss_ay", "ss_free_ax", "ss_free_ay", "ss_cax", "ss_cay"]
        self.n_fm, self.td_ids, self.alg_ids = 0, [], list(range(11))
        self.n_td, self.n_alg, self.n_ext, self.n_ctrl_rows = 0, 11, 4, 0
        self.ncpe, self.nv = 15, 13
        self.ctrl_pos, self.dctrl_pos = [0, 0], []
        self.x = [S("x%d" % k, "x") for k in range(self.nv)]
        P = self.P = {n: S("p_" + n, "p") for n in self.param_names}
        self.track_names = ["kz"]
        self.geom_names = ["hin", "hout", "ihin", "ihout"]
        self.lam_in = [S("lin%d" % r, "w") for r in range(self.ncpe)]
        self.lam_out = [None] * self.ncpe
        obj = S("obj", "w")
        self.uprev, self.unext = [], []
        x, v = self.x, P["ss_v"]
        G0 = models.G0
        ax = x[11] * P["ss_free_ax"] + (1.0 - P["ss_free_ax"]) * P["ss_ax"]
        ay = x[12] * P["ss_free_ay"] + (1.0 - P["ss_free_ay"]) * P["ss_ay"]
        max_angle = 10.0 * math.pi / 180.0
        psi = x[4] * max_angle
        cps, sps = models.cos(psi), models.sin(psi)
        zero = E(g, g.const(0.0))
        q = [x[0], x[1], x[2], x[3], v * cps, -(v * sps), ay * (G0 / v), x[5], x[6], x[7], x[8], zero, zero, zero]
        u = [x[9] * max_angle, zero, x[10], P["brake_bias"]]
        trk = dict(kx=0.0, ky=0.0, kz=0.0, smu=0.0, cmu=1.0, sph=0.0, cph=1.0)
        out = models.f1_model(q, u, P, trk)
        r = out["rates"]
        m = P["mass"]
        inv_g, inv_mg = 1.0 / G0, 1.0 / (m * G0)
        eqs = [r["dw_g"], r["dLroll_mg"], r["dLpitch_mg"], r["roll_balance_mg"],
               (r["du"] * sps + r["dv"] * cps) * inv_g, ax + (r["dv"] * sps - r["du"] * cps) * inv_g, r["domega"] * inv_g] + \
              [r["dL"][i] * inv_mg for i in range(4)]
        ext = []
        for i in range(4):
            pre = "ft_" if i < 2 else "rt_"
            Fz = -(x[5 + i] * (m * G0))                       # maximum_lambda of the load fed to the model (limebeer2014f1.h:123-126)
            lmax = (Fz - P[pre + "Fz1"]) * ((P[pre + "lmax2_deg"] * models.DEG - P[pre + "lmax1_deg"] * models.DEG) * (1.0 / (P[pre + "Fz2"] - P[pre + "Fz1"]))) \
                + P[pre + "lmax1_deg"] * models.DEG
            ext.append(r["lam"][i] / lmax)
        fobj = x[11] * P["ss_cax"] + x[12] * P["ss_cay"]
        self.Q, self.QD, self.ALG, self.EXT, self.CDOT, self.DTDS = [], [], eqs, ext, [], fobj
        self.values = eqs + ext + [fobj]
        self.value_names = ["ALG%d" % k for k in range(11)] + ["EXT%d" % k for k in range(4)] + ["DTDS"]
        rowsA, rowsB = eqs + ext, [None] * 15
        self.rowsA, self.rowsB, self.fobj = rowsA, rowsB, fobj
        phi = obj * fobj
        for k in range(self.ncpe):
            phi = phi + self.lam_in[k] * rowsA[k]
        self.phi = phi
        self._differentiate(rowsA, rowsB, fobj, phi, obj, None, None)

    def _differentiate(self, rowsA, rowsB, fobj, phi, obj, diss, ihin):
        g = self.g
        deriv = diss is None
        xs = [xx.i for xx in self.x]
        # Jacobian blocks (forward sweeps), gradient of f
        fw = []
        roots = [e.i for e in rowsA] + [e.i for e in rowsB if e is not None] + [fobj.i]
        for k in range(self.nv):
            fw.append(g.forward(roots, xs[k]))
        nA = len(rowsA)
        idxB = [r for r in range(self.ncpe) if rowsB[r] is not None]
        zero = g.const(0.0)
        self.jacA = {}      # (row, var) -> expr id
        self.jacB = {}
        self.gradf = []
        for k in range(self.nv):
            for r in range(nA):
                if fw[k][r] != zero:
                    self.jacA[(r, k)] = fw[k][r]
            for q_, r in enumerate(idxB):
                if fw[k][nA + q_] != zero:
                    self.jacB[(r, k)] = fw[k][nA + q_]
            self.gradf.append(fw[k][nA + len(idxB)])

        # Hessian of phi: reverse sweep for the gradient, forward sweeps over it for the columns (forward-over-reverse)
        grad_phi = g.reverse(phi.i, xs)
        self.hess = {}      # (row >= col) -> expr id
        for c in range(self.nv):
            col = g.forward(grad_phi, xs[c])
            for r in range(c, self.nv):
                if col[r] != zero:
                    self.hess[(r, c)] = col[r]
        # coupling slots of the direct-form control-rate penalty: d2 f / d u_j d u_{j-1} = -2 diss / h_in
        self.coupling = []
        if not deriv:
            for c in range(self.n_fm):
                self.coupling.append((obj * (-2.0) * diss[c] * ihin).i)
