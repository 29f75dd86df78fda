// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// CPU restatement of the collocation NLP the reference hands to Ipopt:
//   Optimal_laptime::FG_direct<closed>::compute      /root/reference/src/core/applications/optimal_laptime.hpp:1226-1420
//   Optimal_laptime::FG_derivative<closed>::compute  /root/reference/src/core/applications/optimal_laptime.hpp:1426-1657
// together with the derivative objects CppAD would produce from its tape (eval_grad_f, eval_jac_g, eval_h of
// lion::ipopt_cppad_solve, call site optimal_laptime.hpp:546-551).  Entries are emitted as (row, col, value)
// triplets; parity is always checked by (row, col), never by position.
#ifndef ORACLE_NLP_HPP
#define ORACLE_NLP_HPP

#include <vector>
#include <omp.h>
#include <cstring>
#include "jet.hpp"
#include "f1_model.hpp"
#include "kart_model.hpp"

namespace oracle {

enum { MODEL_F1 = 0, MODEL_KART = 1 };
enum { FORM_DIRECT = 0, FORM_DERIVATIVE = 1 };

struct Desc
{
    int model, form, closed, elevation, n_points, kart_direct_torque;
    const double* s;            // [n_points]
    double track_length;
    const double* track;        // [n_points][6]: theta, mu, phi, dtheta/ds, dmu/ds, dphi/ds
    const double* params;       // model parameter vector
    double sigma;
    double dissipation[2];      // of the two full-mesh controls (F1: delta, throttle; kart: delta, throttle/torque)
    const double* q0;           // open problems: inputs of node 0   [NIN]
    const double* u0;           // open problems: controls of node 0 [NCTRL]
    const double* du0;          // open problems, derivative form: dcontrols_dt of the full-mesh controls at node 0 [2]
    const double* ctrl_default; // default value of every control [NCTRL] (used for the not-optimised ones)
    const double* node_params = nullptr;   // [n_points][n_params]: parameters that vary with the arclength (Model_parameter::operator()(s),
    int n_params = 0;                      //  model_parameter.h:51-77, applied per node at dynamic_model_car.hpp:62-63), or null
    const double* params_at(int node) const { return node_params ? node_params + (size_t)node * n_params : params; }
    double wind[2] = { 0.0, 0.0 };         // northward, eastward wind velocity [m/s] (chassis.h:276-277)
    // One quantity that couples the whole lap in the reference, restated node-wise (the formulation of the product, see
    // fastest_lap_b200/codegen/nlpnode.py): 1 = the brake bias as a control held constant over stretches of the lap
    // (Control_variables::control_array_at_s with a hypermesh / constant control, detail/optimal_laptime_control_variables.h:190-222),
    // 2 = a restricted integral quantity (limebeer2014f1.h:284-306; optimal_laptime.hpp:1316-1317, 1352-1353, 1409-1416).
    // Every node gains the variables (a, q) after its inputs and every element one trapezoid row
    //     a_j + rj q_j - bout_i a_i - (awout_i I_i + awin_j I_j) = 0,     I = integrand dt/ds.
    int aug_kind = 0;
    const double* aug_jump = nullptr;      // kind 1: [n_points], 1 where the node starts a new stretch
    double aug_weights[5] = { 0, 0, 0, 0, 0 };   // kind 2: engine-energy, tire-fl, -fr, -rl, -rr energy
};

constexpr int MAX_ROWS = 14;

// What one node contributes: everything is a function of the node's own variables only.
template <int NV>
struct NodeRows
{
    Jet2<NV> q[MAX_ROWS];      // states of the time-derivative equations
    Jet2<NV> qd[MAX_ROWS];     // dstates_dt * dt/ds of the time-derivative equations
    Jet2<NV> alg[4];           // algebraic rows: dstates_dt[eq] / (dt/ds)
    Jet2<NV> ext[6];           // extra constraints
    Jet2<NV> dtds;
    Jet2<NV> cdot[2];          // derivative form: dcontrol_dt * dt/ds of the full-mesh controls
    Jet2<NV> ctrl[2];          // the full-mesh controls themselves
    Jet2<NV> dctrl[2];         // derivative form: dcontrols_dt variables
    Jet2<NV> qa;               // augmented row: the value seen by the element that ENDS at the node (a + rj q)
};

struct Layout
{
    int nv;          // variables per node
    int n_td, n_alg, n_ext, n_fm, n_ctrl_rows;
    int ncpe;        // constraints per element
    int td_ids[MAX_ROWS], alg_ids[4];
    int ctrl_pos[2]; // position of the full-mesh controls within the node variables
    int dctrl_pos[2];
    int n_nodes_var; // nodes that carry variables
    int n_elements;
    int n, m;
    int n_aug, aug_pos;  // augmented state: count (its row is the last trapezoid row), position of `a` in the node (q follows)
};

inline Layout make_layout(const Desc& d)
{
    Layout L;
    std::memset(&L, 0, sizeof(L));
    const bool deriv = d.form == FORM_DERIVATIVE;
    if (d.model == MODEL_F1)
    {
        // limebeer2014f1.h:185-225 ; test/vehicles/limebeer2014f1_test.cpp:176-198
        if (d.elevation)
        {
            const int ids[13] = {0,1,2,3,4,5,6,7,8,9,10,12,13};
            L.n_td = 13; for (int i = 0; i < 13; ++i) L.td_ids[i] = ids[i];
            L.n_alg = 0;
        }
        else
        {
            const int ids[9] = {0,1,2,3,4,5,6,12,13};
            L.n_td = 9; for (int i = 0; i < 9; ++i) L.td_ids[i] = ids[i];
            L.n_alg = 4; for (int i = 0; i < 4; ++i) L.alg_ids[i] = 7 + i;
        }
        L.n_ext = F1_NEXTRA;
        L.n_fm = 2;
        L.n_aug = d.aug_kind ? 1 : 0;
        L.aug_pos = F1_NIN - 1;
        L.n_td += L.n_aug;
        L.nv = (F1_NIN - 1) + 2 * L.n_aug + 2 + (deriv ? 2 : 0);
        L.ctrl_pos[0] = 13 + 2 * L.n_aug; L.ctrl_pos[1] = 14 + 2 * L.n_aug;
        L.dctrl_pos[0] = 15; L.dctrl_pos[1] = 16;
    }
    else
    {
        // dynamic_model_car.h:170-188: all kart equations are time derivatives except time itself
        const int ids[12] = {0,1,2,3,4,5,6,7,8,9,11,12};
        L.n_td = 12; for (int i = 0; i < 12; ++i) L.td_ids[i] = ids[i];
        L.n_alg = 0;
        L.n_ext = KART_NEXTRA;
        L.n_fm = 2;
        L.nv = (KART_NIN - 1) + 2 + (deriv ? 2 : 0);
        L.ctrl_pos[0] = 12; L.ctrl_pos[1] = 13;
        L.dctrl_pos[0] = 14; L.dctrl_pos[1] = 15;
    }
    L.n_ctrl_rows = deriv ? L.n_fm : 0;
    L.ncpe = L.n_td + L.n_alg + L.n_ext + L.n_ctrl_rows;
    L.n_nodes_var = d.closed ? d.n_points : d.n_points - 1;
    L.n_elements = d.closed ? d.n_points : d.n_points - 1;      // optimal_laptime.hpp:73-74
    L.n = L.n_nodes_var * L.nv;
    L.m = L.n_elements * L.ncpe;
    return L;
}

// Evaluate one node.  xn = the node's NV variables (or nullptr for the fixed first node of an open problem).
template <int NV>
inline void eval_node(const Desc& d, const Layout& L, int i, const double* xn, NodeRows<NV>& R)
{
    typedef Jet2<NV> J;
    const bool deriv = d.form == FORM_DERIVATIVE;
    TrackNode tk;
    const double* t = d.track + 6 * i;
    tk.theta = t[0]; tk.mu = t[1]; tk.phi = t[2]; tk.dtheta = t[3]; tk.dmu = t[4]; tk.dphi = t[5];
    tk.wind_north = d.wind[0]; tk.wind_east = d.wind[1];

    if (d.model == MODEL_F1)
    {
        J q[F1_NIN], u[F1_NCTRL];
        for (int c = 0; c < F1_NCTRL; ++c) u[c] = J(d.ctrl_default[c]);
        if (xn)
        {
            // optimal_laptime.hpp:1269-1285: inputs before time, inputs after time, full-mesh controls
            int k = 0;
            for (int j = 0; j < F1_ITIME; ++j, ++k) q[j] = J::variable(xn[k], k);
            q[F1_ITIME] = J(0.0);
            for (int j = F1_ITIME + 1; j < F1_NIN; ++j, ++k) q[j] = J::variable(xn[k], k);
            J a_var(0.0), q_var(0.0);
            if (L.n_aug)
            {
                a_var = J::variable(xn[k], k); ++k;
                q_var = J::variable(xn[k], k); ++k;
                if (d.aug_kind == 1) u[F1_CBIAS] = a_var;
                const bool integral = d.aug_kind == 2;
                const double rj = integral ? 0.0 : (d.aug_jump[i] != 0.0 ? 1.0 : 0.0);
                const double bout = (integral && i == 0) ? 0.0 : 1.0;
                R.q[L.n_td - 1] = bout * a_var;
                R.qa = a_var + rj * q_var;
            }
            u[F1_CDELTA] = J::variable(xn[k], k); ++k;
            u[F1_CTHROTTLE] = J::variable(xn[k], k); ++k;
            if (deriv) { R.dctrl[0] = J::variable(xn[k], k); ++k; R.dctrl[1] = J::variable(xn[k], k); ++k; }
        }
        else
        {
            for (int j = 0; j < F1_NIN; ++j) q[j] = J(d.q0[j]);
            for (int c = 0; c < F1_NCTRL; ++c) u[c] = J(d.u0[c]);
            if (deriv) { R.dctrl[0] = J(d.du0[0]); R.dctrl[1] = J(d.du0[1]); }
        }
        F1Out<J> o;
        f1_car(q, u, d.params_at(i), tk, o);
        for (int r = 0; r < L.n_td - L.n_aug; ++r) { R.q[r] = o.states[L.td_ids[r]]; R.qd[r] = o.dstates[L.td_ids[r]]; }
        if (L.n_aug)
        {
            // integrand times dt/ds (optimal_laptime.hpp:1317, 1352)
            J I(0.0);
            for (int k = 0; k < 5; ++k) if (d.aug_kind == 2 && d.aug_weights[k] != 0.0) I = I + d.aug_weights[k] * o.integrands[k];
            R.qd[L.n_td - 1] = I * o.dstates[F1_ITIME];
        }
        for (int r = 0; r < L.n_alg; ++r) R.alg[r] = o.dstates[L.alg_ids[r]] / o.dtds;      // optimal_laptime.hpp:1341
        for (int r = 0; r < L.n_ext; ++r) R.ext[r] = o.extra[r];
        R.dtds = o.dstates[F1_ITIME];
        R.ctrl[0] = u[F1_CDELTA]; R.ctrl[1] = u[F1_CTHROTTLE];
    }
    else
    {
        J q[KART_NIN], u[KART_NCTRL];
        for (int c = 0; c < KART_NCTRL; ++c) u[c] = J(d.ctrl_default[c]);
        if (xn)
        {
            int k = 0;
            for (int j = 0; j < KART_ITIME; ++j, ++k) q[j] = J::variable(xn[k], k);
            q[KART_ITIME] = J(0.0);
            for (int j = KART_ITIME + 1; j < KART_NIN; ++j, ++k) q[j] = J::variable(xn[k], k);
            u[0] = J::variable(xn[k], k); ++k;
            u[1] = J::variable(xn[k], k); ++k;
            if (deriv) { R.dctrl[0] = J::variable(xn[k], k); ++k; R.dctrl[1] = J::variable(xn[k], k); ++k; }
        }
        else
        {
            for (int j = 0; j < KART_NIN; ++j) q[j] = J(d.q0[j]);
            for (int c = 0; c < KART_NCTRL; ++c) u[c] = J(d.u0[c]);
            if (deriv) { R.dctrl[0] = J(d.du0[0]); R.dctrl[1] = J(d.du0[1]); }
        }
        KartOut<J> o;
        kart_car(q, u, d.params_at(i), tk, d.kart_direct_torque != 0, true, o);
        for (int r = 0; r < L.n_td; ++r) { R.q[r] = o.states[L.td_ids[r]]; R.qd[r] = o.dstates[L.td_ids[r]]; }
        for (int r = 0; r < L.n_ext; ++r) R.ext[r] = o.extra[r];
        R.dtds = o.dstates[KART_ITIME];
        R.ctrl[0] = u[0]; R.ctrl[1] = u[1];
    }
    if (deriv)
        for (int c = 0; c < L.n_fm; ++c) R.cdot[c] = R.dctrl[c] * R.dtds;     // optimal_laptime.hpp:1580-1581
}

struct Triplets
{
    std::vector<int> row, col;
    std::vector<double> val;
    void add(int r, int c, double v) { row.push_back(r); col.push_back(c); val.push_back(v); }
};

struct EvalOut
{
    double f;
    std::vector<double> grad, g;
    Triplets jac, hess;     // hess: lower triangle (row >= col), duplicates must be summed
};

template <int NV>
void eval_nlp_t(const Desc& d, const double* x, const double* lambda, double obj_factor, bool want_derivs, EvalOut& out)
{
    typedef Jet2<NV> J;
    const Layout L = make_layout(d);
    const int np = d.n_points;
    const bool deriv = d.form == FORM_DERIVATIVE;
    const double sigma = d.sigma;
    auto var0 = [&](int node) { return d.closed ? node * NV : (node - 1) * NV; };     // -NV for the fixed node of open problems
    // node rows are evaluated where the assembly needs them (one pass over the lap per thread chunk, two live nodes per
    // thread), not stored for the whole lap: a lap of 8000 nodes would be 435 MB of second-order jets
    auto node_rows = [&](int i, NodeRows<NV>& Rn) {
        const double* xn = (d.closed || i > 0) ? x + var0(i) : nullptr;
        eval_node<NV>(d, L, i, xn, Rn);
    };
    // derivative form: the control rates of nodes 0 and n-1 enter the penalty of every element (reference quirk, see below)
    std::vector<NodeRows<NV>> Rends(deriv ? 2 : 0);
    if (deriv) { node_rows(0, Rends[0]); node_rows(np - 1, Rends[1]); }

    out.f = 0.0;
    out.g.assign(L.m, 0.0);
    out.grad.assign(L.n, 0.0);
    out.jac = Triplets();
    out.hess = Triplets();

    // Element assembly: the elements are cut into contiguous chunks, one per thread.  Every chunk accumulates the gradient and
    // the per-node Hessian blocks of ITS nodes (first left node .. last right node) in private arrays and emits its triplets
    // into private lists; the chunks are then joined in element order, so the result does not depend on the thread count
    // except for the order of the two additions at a node shared by two chunks.
    struct Chunk
    {
        int e0 = 0, e1 = 0;
        double f = 0.0;
        std::vector<oreal> grad, H;      // [(e1 - e0 + 1) nodes][NV], [..][NV*NV]
        Triplets jac, hess;
    };
    int n_chunks = 1;
    #pragma omp parallel
    {
        #pragma omp single
        n_chunks = omp_get_num_threads();
    }
    if (n_chunks > L.n_elements) n_chunks = L.n_elements > 0 ? L.n_elements : 1;
    std::vector<Chunk> chunks(n_chunks);

    #pragma omp parallel for schedule(static, 1)
    for (int ck = 0; ck < n_chunks; ++ck)
    {
        Chunk& C = chunks[ck];
        C.e0 = (int)((long)L.n_elements * ck / n_chunks);
        C.e1 = (int)((long)L.n_elements * (ck + 1) / n_chunks);
        const int nn = C.e1 - C.e0 + 1;
        if (want_derivs) { C.grad.assign((size_t)nn * NV, 0.0); C.H.assign((size_t)nn * NV * NV, 0.0); }
        // local slot of a node of this chunk: left node of element e is slot e - e0, right node slot e - e0 + 1
        auto add_grad = [&](int slot, int node, const J& a, double c) {
            if (!want_derivs) return;
            if (!d.closed && node == 0) return;
            oreal* gr = C.grad.data() + (size_t)slot * NV;
            for (int k = 0; k < NV; ++k) gr[k] += c * a.g[k];
        };
        auto add_hess = [&](int slot, int node, const J& a, double c) {
            if (!want_derivs) return;
            if (!d.closed && node == 0) return;
            oreal* Hn = C.H.data() + (size_t)slot * NV * NV;
            for (int r = 0; r < NV; ++r) for (int cc = 0; cc <= r; ++cc) Hn[r * NV + cc] += c * a.h[r][cc];
        };
        auto add_jac = [&](int row, int node, const J& a, double c) {
            if (!want_derivs) return;
            if (!d.closed && node == 0) return;
            const int c0 = var0(node);
            for (int k = 0; k < NV; ++k) C.jac.add(row, c0 + k, c * a.g[k]);
        };

        std::vector<NodeRows<NV>> pair(2);
        int cur = 0;
        if (C.e1 > C.e0) node_rows(C.e0, pair[0]);         // (element e starts at node e)
        for (int e = C.e0; e < C.e1; ++e, cur ^= 1)
        {
            // element e joins node i -> j ; the closing element of a closed lap joins n-1 -> 0 with h = L - s.back()
            const bool closing = d.closed && e == np - 1;
            const int i = closing ? np - 1 : e;
            const int j = closing ? 0 : e + 1;
            const int si = e - C.e0, sj = si + 1;
            const double h = closing ? d.track_length - d.s[np - 1] : d.s[j] - d.s[i];
            node_rows(j, pair[cur ^ 1]);
            const NodeRows<NV>& A = pair[cur];
            const NodeRows<NV>& B = pair[cur ^ 1];
            int row = e * L.ncpe;

            // objective: integral of time (optimal_laptime.hpp:1333, 1378)
            C.f += h * ((1.0 - sigma) * A.dtds.v + sigma * B.dtds.v);
            add_grad(si, i, A.dtds, h * (1.0 - sigma));
            add_grad(sj, j, B.dtds, h * sigma);
            add_hess(si, i, A.dtds, obj_factor * h * (1.0 - sigma));
            add_hess(sj, j, B.dtds, obj_factor * h * sigma);

            // control penalisation
            for (int c = 0; c < L.n_fm; ++c)
            {
                const double diss = d.dissipation[c];
                if (!deriv)
                {
                    // optimal_laptime.hpp:1362-1366, 1401-1402
                    const double du = B.ctrl[c].v - A.ctrl[c].v;
                    const double derivative = du / h;
                    C.f += diss * (derivative * derivative) * h;
                    if (want_derivs)
                    {
                        const int pi = var0(i) + L.ctrl_pos[c], pj = var0(j) + L.ctrl_pos[c];
                        const bool vi = d.closed || i > 0;
                        if (vi) C.grad[(size_t)si * NV + L.ctrl_pos[c]] += -2.0 * diss * du / h;
                        C.grad[(size_t)sj * NV + L.ctrl_pos[c]] += 2.0 * diss * du / h;
                        const double hv = obj_factor * 2.0 * diss / h;
                        if (vi) C.hess.add(pi, pi, hv);
                        C.hess.add(pj, pj, hv);
                        if (vi) C.hess.add(pi > pj ? pi : pj, pi > pj ? pj : pi, -hv);
                    }
                }
                else
                {
                    // optimal_laptime.hpp:1553 (note the reference's dcontrols_dt[i-i] = dcontrols_dt[0]) and :1611
                    const int il = closing ? np - 1 : 0;
                    const J& dl = Rends[closing ? 1 : 0].dctrl[c];
                    const J& dr = B.dctrl[c];
                    C.f += 0.5 * diss * (dl.v * dl.v + dr.v * dr.v) * h;
                    if (want_derivs)
                    {
                        // node `il` may lie outside this chunk: its gradient entry and Hessian diagonal go out as triplets
                        const int pl = var0(il) + L.dctrl_pos[c], pr = var0(j) + L.dctrl_pos[c];
                        const bool vl = d.closed || il > 0;
                        if (vl) { C.hess.add(pl, pl, obj_factor * diss * h); C.hess.add(-1 - pl, 0, diss * dl.v * h); }
                        C.grad[(size_t)sj * NV + L.dctrl_pos[c]] += diss * dr.v * h;
                        C.hess.add(pr, pr, obj_factor * diss * h);
                    }
                }
            }

            // trapezoidal rows (optimal_laptime.hpp:1337-1338, 1383-1384)
            for (int r = 0; r < L.n_td - L.n_aug; ++r, ++row)
            {
                out.g[row] = B.q[r].v - A.q[r].v - h * ((1.0 - sigma) * A.qd[r].v + sigma * B.qd[r].v);
                const double lam = lambda ? lambda[row] : 0.0;
                add_jac(row, i, A.q[r], -1.0);  add_jac(row, i, A.qd[r], -h * (1.0 - sigma));
                add_jac(row, j, B.q[r], 1.0);   add_jac(row, j, B.qd[r], -h * sigma);
                add_hess(si, i, A.q[r], -lam);      add_hess(si, i, A.qd[r], -lam * h * (1.0 - sigma));
                add_hess(sj, j, B.q[r], lam);       add_hess(sj, j, B.qd[r], -lam * h * sigma);
            }
            if (L.n_aug)
            {
                // augmented row; the closing element weights its two integrands the other way round (optimal_laptime.hpp:1409)
                const int r = L.n_td - 1;
                const double wa = h * (closing ? sigma : 1.0 - sigma), wb = h * (closing ? 1.0 - sigma : sigma);
                out.g[row] = B.qa.v - A.q[r].v - (wa * A.qd[r].v + wb * B.qd[r].v);
                const double lam = lambda ? lambda[row] : 0.0;
                add_jac(row, i, A.q[r], -1.0);  add_jac(row, i, A.qd[r], -wa);
                add_jac(row, j, B.qa, 1.0);     add_jac(row, j, B.qd[r], -wb);
                add_hess(si, i, A.q[r], -lam);  add_hess(si, i, A.qd[r], -lam * wa);
                add_hess(sj, j, B.qa, lam);     add_hess(sj, j, B.qd[r], -lam * wb);
                ++row;
            }
            // algebraic rows (optimal_laptime.hpp:1340-1341, 1386-1387)
            for (int r = 0; r < L.n_alg; ++r, ++row)
            {
                out.g[row] = B.alg[r].v;
                const double lam = lambda ? lambda[row] : 0.0;
                add_jac(row, j, B.alg[r], 1.0);
                add_hess(sj, j, B.alg[r], lam);
            }
            // extra constraints (optimal_laptime.hpp:1344-1347, 1390-1394)
            for (int r = 0; r < L.n_ext; ++r, ++row)
            {
                out.g[row] = B.ext[r].v;
                const double lam = lambda ? lambda[row] : 0.0;
                add_jac(row, j, B.ext[r], 1.0);
                add_hess(sj, j, B.ext[r], lam);
            }
            // derivative form: control rate rows (optimal_laptime.hpp:1573-1583, 1632-1641)
            for (int c = 0; c < L.n_ctrl_rows; ++c, ++row)
            {
                out.g[row] = B.ctrl[c].v - A.ctrl[c].v - h * ((1.0 - sigma) * A.cdot[c].v + sigma * B.cdot[c].v);
                const double lam = lambda ? lambda[row] : 0.0;
                add_jac(row, i, A.ctrl[c], -1.0); add_jac(row, i, A.cdot[c], -h * (1.0 - sigma));
                add_jac(row, j, B.ctrl[c], 1.0);  add_jac(row, j, B.cdot[c], -h * sigma);
                add_hess(si, i, A.cdot[c], -lam * h * (1.0 - sigma));
                add_hess(sj, j, B.cdot[c], -lam * h * sigma);
            }
        }

        if (want_derivs)
        {
            // merge the duplicate jacobian triplets (q and qd contributions of the same (row, col))
            Triplets merged;
            const size_t nt = C.jac.val.size();
            size_t k = 0;
            while (k < nt)
            {
                // contributions of one (row, node) come in runs of NV, possibly two consecutive runs for the same node
                if (k + 2 * NV <= nt && C.jac.row[k] == C.jac.row[k + NV] && C.jac.col[k] == C.jac.col[k + NV])
                {
                    for (int q = 0; q < NV; ++q) merged.add(C.jac.row[k + q], C.jac.col[k + q], C.jac.val[k + q] + C.jac.val[k + NV + q]);
                    k += 2 * NV;
                }
                else
                {
                    for (int q = 0; q < NV; ++q) merged.add(C.jac.row[k + q], C.jac.col[k + q], C.jac.val[k + q]);
                    k += NV;
                }
            }
            C.jac = merged;
        }
    }

    // join the chunks in element order
    std::vector<oreal> H;
    if (want_derivs) H.assign((size_t)np * NV * NV, 0.0);
    for (int ck = 0; ck < n_chunks; ++ck)
    {
        const Chunk& C = chunks[ck];
        out.f += C.f;
        if (!want_derivs) continue;
        for (int slot = 0; slot <= C.e1 - C.e0 && C.e1 > C.e0; ++slot)
        {
            int node = C.e0 + slot;
            if (node >= np) node -= np;                       // right node of the closing element
            if (!d.closed && node == 0) continue;
            double* gr = out.grad.data() + var0(node);
            for (int k = 0; k < NV; ++k) gr[k] += C.grad[(size_t)slot * NV + k];
            oreal* Hn = H.data() + (size_t)node * NV * NV;
            const oreal* Hc = C.H.data() + (size_t)slot * NV * NV;
            for (int k = 0; k < NV * NV; ++k) Hn[k] += Hc[k];
        }
        out.jac.row.insert(out.jac.row.end(), C.jac.row.begin(), C.jac.row.end());
        out.jac.col.insert(out.jac.col.end(), C.jac.col.begin(), C.jac.col.end());
        out.jac.val.insert(out.jac.val.end(), C.jac.val.begin(), C.jac.val.end());
        for (size_t k = 0; k < C.hess.val.size(); ++k)
        {
            if (C.hess.row[k] < 0) out.grad[-1 - C.hess.row[k]] += C.hess.val[k];      // gradient entry of a node outside the chunk
            else out.hess.add(C.hess.row[k], C.hess.col[k], C.hess.val[k]);
        }
    }

    if (want_derivs)
    {
        for (int node = 0; node < np; ++node)
        {
            if (!d.closed && node == 0) continue;
            const oreal* Hn = H.data() + (size_t)node * NV * NV;
            const int c0 = var0(node);
            for (int r = 0; r < NV; ++r) for (int c = 0; c <= r; ++c) out.hess.add(c0 + r, c0 + c, Hn[r * NV + c]);
        }
    }
}

inline void eval_nlp(const Desc& d, const double* x, const double* lambda, double obj_factor, bool want_derivs, EvalOut& out)
{
    const Layout L = make_layout(d);
    switch (L.nv)
    {
    case 14: eval_nlp_t<14>(d, x, lambda, obj_factor, want_derivs, out); break;
    case 15: eval_nlp_t<15>(d, x, lambda, obj_factor, want_derivs, out); break;
    case 16: eval_nlp_t<16>(d, x, lambda, obj_factor, want_derivs, out); break;
    case 17: eval_nlp_t<17>(d, x, lambda, obj_factor, want_derivs, out); break;
    default: break;
    }
}

} // namespace oracle
#endif
