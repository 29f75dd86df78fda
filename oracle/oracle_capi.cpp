// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline leg.  The product (fastest_lap_b200/) never links or loads this file.
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>
#include "nlp.hpp"
#include "steady.hpp"

using namespace oracle;

extern "C" {

struct orc_desc
{
    int model, form, closed, elevation, n_points, kart_direct_torque;
    const double* s;
    double track_length;
    const double* track;
    const double* params;
    double sigma;
    double dissipation[2];
    const double* q0;
    const double* u0;
    const double* du0;
    const double* ctrl_default;
    const double* node_params;     // [n_points][orc_nparams(model)] or null
    double wind[2];                // northward, eastward [m/s]
    int aug_kind;                  // see oracle::Desc
    const double* aug_jump;
    double aug_weights[5];
};

static Desc to_desc(const orc_desc* c)
{
    Desc d;
    d.model = c->model; d.form = c->form; d.closed = c->closed; d.elevation = c->elevation; d.n_points = c->n_points;
    d.kart_direct_torque = c->kart_direct_torque;
    d.s = c->s; d.track_length = c->track_length; d.track = c->track; d.params = c->params; d.sigma = c->sigma;
    d.dissipation[0] = c->dissipation[0]; d.dissipation[1] = c->dissipation[1];
    d.q0 = c->q0; d.u0 = c->u0; d.du0 = c->du0; d.ctrl_default = c->ctrl_default;
    d.node_params = c->node_params; d.n_params = c->model == MODEL_F1 ? (int)F1_NPARAMS : (int)K_NPARAMS;
    d.wind[0] = c->wind[0]; d.wind[1] = c->wind[1];
    d.aug_kind = c->aug_kind; d.aug_jump = c->aug_jump;
    for (int k = 0; k < 5; ++k) d.aug_weights[k] = c->aug_weights[k];
    return d;
}

int orc_nparams(int model) { return model == MODEL_F1 ? (int)F1_NPARAMS : (int)K_NPARAMS; }

void orc_sizes(const orc_desc* c, int* n, int* m, int* nv, int* ncpe, long* max_jac, long* max_hess)
{
    const Desc d = to_desc(c);
    const Layout L = make_layout(d);
    *n = L.n; *m = L.m; *nv = L.nv; *ncpe = L.ncpe;
    *max_jac = (long)L.m * 2 * L.nv;
    *max_hess = (long)d.n_points * (L.nv * (L.nv + 1) / 2) + (long)L.n_elements * 3 * L.n_fm;
}

// Full evaluation: f, grad f, g, Jacobian triplets, Hessian-of-the-Lagrangian triplets (lower triangle, duplicates summed by the caller)
int orc_eval(const orc_desc* c, const double* x, const double* lambda, double obj_factor, int want_derivs, int n_threads,
             double* f, double* grad, double* g,
             int* jrow, int* jcol, double* jval, long* njac,
             int* hrow, int* hcol, double* hval, long* nhess)
{
    const Desc d = to_desc(c);
    if (n_threads > 0) omp_set_num_threads(n_threads);
    EvalOut out;
    eval_nlp(d, x, lambda, obj_factor, want_derivs != 0, out);
    *f = out.f;
    std::memcpy(g, out.g.data(), out.g.size() * sizeof(double));
    if (want_derivs)
    {
        std::memcpy(grad, out.grad.data(), out.grad.size() * sizeof(double));
        *njac = (long)out.jac.val.size();
        std::memcpy(jrow, out.jac.row.data(), out.jac.row.size() * sizeof(int));
        std::memcpy(jcol, out.jac.col.data(), out.jac.col.size() * sizeof(int));
        std::memcpy(jval, out.jac.val.data(), out.jac.val.size() * sizeof(double));
        *nhess = (long)out.hess.val.size();
        std::memcpy(hrow, out.hess.row.data(), out.hess.row.size() * sizeof(int));
        std::memcpy(hcol, out.hess.col.data(), out.hess.col.size() * sizeof(int));
        std::memcpy(hval, out.hess.val.data(), out.hess.val.size() * sizeof(double));
    }
    return 0;
}

// Timing leg for bench.py's cpu_baseline: `reps` full callback rounds (f, grad, g, Jacobian, Hessian) over the lap,
// returns seconds per round.
double orc_time_eval(const orc_desc* c, const double* x, const double* lambda, double obj_factor, int reps, int n_threads)
{
    const Desc d = to_desc(c);
    if (n_threads > 0) omp_set_num_threads(n_threads);
    EvalOut out;
    const double t0 = omp_get_wtime();
    for (int r = 0; r < reps; ++r) eval_nlp(d, x, lambda, obj_factor, true, out);
    const double t1 = omp_get_wtime();
    return (t1 - t0) / reps;
}

// Output channels of the model at every node of a lap: out [45][n_points]; x as the NLP takes it (open problems: without node 0)
int orc_outputs(const orc_desc* c, const double* x, double* out)
{
    const Desc d = to_desc(c);
    const Layout L = make_layout(d);
    const int np = d.n_points;
    for (int i = 0; i < np; ++i)
    {
        const double* xn = (d.closed || i > 0) ? x + (size_t)(d.closed ? i : i - 1) * L.nv : nullptr;
        TrackNode tk;
        const double* t = d.track + 6 * i;
        tk.theta = t[0]; tk.mu = t[1]; tk.phi = t[2]; tk.dtheta = t[3]; tk.dmu = t[4]; tk.dphi = t[5];
        tk.wind_north = d.wind[0]; tk.wind_east = d.wind[1];
        const int nin = d.model == MODEL_F1 ? (int)F1_NIN : (int)KART_NIN, nct = d.model == MODEL_F1 ? (int)F1_NCTRL : (int)KART_NCTRL;
        const int itime = d.model == MODEL_F1 ? (int)F1_ITIME : (int)KART_ITIME;
        double q[16], u[4];
        for (int k = 0; k < nct; ++k) u[k] = d.ctrl_default[k];
        if (xn)
        {
            int k = 0;
            for (int j = 0; j < nin; ++j) q[j] = (j == itime) ? 0.0 : xn[k++];
            if (L.n_aug) { if (d.aug_kind == 1) u[F1_CBIAS] = xn[k]; k += 2; }
            if (d.model == MODEL_F1) { u[F1_CDELTA] = xn[k++]; u[F1_CTHROTTLE] = xn[k++]; }
            else { u[0] = xn[k++]; u[1] = xn[k++]; }
        }
        else
        {
            for (int j = 0; j < nin; ++j) q[j] = d.q0[j];
            for (int k = 0; k < nct; ++k) u[k] = d.u0[k];
        }
        if (d.model == MODEL_F1)
        {
            F1Out<double> o;
            f1_car<double>(q, u, d.params_at(i), tk, o);
            for (int k = 0; k < 45; ++k) out[(size_t)k * np + i] = o.channels[k];
        }
        else
        {
            KartOut<double> o;
            kart_car<double>(q, u, d.params_at(i), tk, d.kart_direct_torque != 0, true, o);
            for (int k = 0; k < 45; ++k) out[(size_t)k * np + i] = o.channels[k];
        }
    }
    return 0;
}

int orc_max_threads() { return omp_get_max_threads(); }

// wind used by the single-node entry points below (orc_f1_car, orc_kart_car): northward, eastward [m/s]
static double g_wind[2] = { 0.0, 0.0 };
void orc_set_wind(double northward, double eastward) { g_wind[0] = northward; g_wind[1] = eastward; }

// One node of the F1, values only (unit-level checks against the reference's closed-form tests)
void orc_f1_car(const double* q, const double* u, const double* P, const double* track6,
                double* states, double* dstates, double* dtds, double* extra, double* diag /* kappa[4], lambda[4], omega_w[4], Fx[4], Fy[4], N[4] */)
{
    TrackNode tk; tk.theta = track6[0]; tk.mu = track6[1]; tk.phi = track6[2]; tk.dtheta = track6[3]; tk.dmu = track6[4]; tk.dphi = track6[5];
    tk.wind_north = g_wind[0]; tk.wind_east = g_wind[1];
    F1Out<double> o;
    f1_car<double>(q, u, P, tk, o);
    for (int i = 0; i < F1_NIN; ++i) { states[i] = o.states[i]; dstates[i] = o.dstates[i]; }
    *dtds = o.dtds;
    for (int i = 0; i < F1_NEXTRA; ++i) extra[i] = o.extra[i];
    for (int i = 0; i < 4; ++i)
    {
        diag[i] = o.kappa[i]; diag[4 + i] = o.lambda[i]; diag[8 + i] = o.omega_w[i];
        diag[12 + i] = o.Fx_t[i]; diag[16 + i] = o.Fy_t[i]; diag[20 + i] = o.N[i];
    }
}

// One node of the kart, values only
void orc_kart_car(const double* q, const double* u, const double* P, const double* track6, int direct_torque, int curvilinear,
                  double* states, double* dstates, double* dtds, double* extra, double* diag /* kappa[4], lambda[4], N[4], Fx[4], Fy[4], w[4], force[3], torque[3] */)
{
    TrackNode tk; tk.theta = track6[0]; tk.mu = track6[1]; tk.phi = track6[2]; tk.dtheta = track6[3]; tk.dmu = track6[4]; tk.dphi = track6[5];
    tk.wind_north = g_wind[0]; tk.wind_east = g_wind[1];
    KartOut<double> o;
    kart_car<double>(q, u, P, tk, direct_torque != 0, curvilinear != 0, o);
    for (int i = 0; i < KART_NIN; ++i) { states[i] = o.states[i]; dstates[i] = o.dstates[i]; }
    *dtds = o.dtds;
    for (int i = 0; i < KART_NEXTRA; ++i) extra[i] = o.extra[i];
    for (int i = 0; i < 4; ++i)
    {
        diag[i] = o.kappa[i]; diag[4 + i] = o.lambda[i]; diag[8 + i] = o.N[i];
        diag[12 + i] = o.Fx_t[i]; diag[16 + i] = o.Fy_t[i]; diag[20 + i] = o.w[i];
    }
    for (int i = 0; i < 3; ++i) { diag[24 + i] = o.force[i]; diag[27 + i] = o.torque[i]; }
}

// Steady-state NLP of the kart at one point: f, c[12], grad f [8], dense Jacobian [12][8], dense Hessian of the Lagrangian [8][8]
// spec = {v, ax, ay, free_ax, free_ay, cax, cay}
void orc_kart_steady(const double* x, const double* P, const double* spec, const double* lambda, double obj_factor,
                     double* f, double* c, double* grad, double* jac, double* hess)
{
    SteadySpec sp{spec[0], spec[1], spec[2], spec[3], spec[4], spec[5], spec[6]};
    using J = Jet2<KSS_NX>;
    J xj[KSS_NX], cj[KSS_NC], fj;
    for (int k = 0; k < KSS_NX; ++k) xj[k] = J::variable(x[k], k);
    kart_steady<J>(xj, P, sp, cj, fj);
    *f = fj.v;
    for (int r = 0; r < KSS_NC; ++r) c[r] = cj[r].v;
    if (!grad) return;
    for (int k = 0; k < KSS_NX; ++k) grad[k] = fj.g[k];
    for (int r = 0; r < KSS_NC; ++r) for (int k = 0; k < KSS_NX; ++k) jac[r * KSS_NX + k] = cj[r].g[k];
    for (int i = 0; i < KSS_NX; ++i)
        for (int k = 0; k <= i; ++k)
        {
            double h = obj_factor * fj.h[i][k];
            for (int r = 0; r < KSS_NC; ++r) h += lambda[r] * cj[r].h[i][k];
            hess[i * KSS_NX + k] = h; hess[k * KSS_NX + i] = h;
        }
}

// The same for the F1 (13 variables, 15 rows)
void orc_f1_steady(const double* x, const double* P, const double* spec, const double* lambda, double obj_factor,
                   double* f, double* c, double* grad, double* jac, double* hess)
{
    SteadySpec sp{spec[0], spec[1], spec[2], spec[3], spec[4], spec[5], spec[6]};
    using J = Jet2<FSS_NX>;
    J xj[FSS_NX], cj[FSS_NC], fj;
    for (int k = 0; k < FSS_NX; ++k) xj[k] = J::variable(x[k], k);
    f1_steady<J>(xj, P, sp, cj, fj);
    *f = fj.v;
    for (int r = 0; r < FSS_NC; ++r) c[r] = cj[r].v;
    if (!grad) return;
    for (int k = 0; k < FSS_NX; ++k) grad[k] = fj.g[k];
    for (int r = 0; r < FSS_NC; ++r) for (int k = 0; k < FSS_NX; ++k) jac[r * FSS_NX + k] = cj[r].g[k];
    for (int i = 0; i < FSS_NX; ++i)
        for (int k = 0; k <= i; ++k)
        {
            double h = obj_factor * fj.h[i][k];
            for (int r = 0; r < FSS_NC; ++r) h += lambda[r] * cj[r].h[i][k];
            hess[i * FSS_NX + k] = h; hess[k * FSS_NX + i] = h;
        }
}

} // extern "C"
