// Host driver and C-ABI of the device-resident interior-point solver for batches of laps (include/fastestlap_b200.h,
// section "interior-point solver").  Replaces, for the collocation NLP, what Ipopt + MUMPS do behind
// lion::ipopt_cppad_solve in the reference (src/core/applications/optimal_laptime.hpp:528-556).  The host sequences
// phases and reads a few counters per phase; every vector, matrix and per-lap decision lives on the device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <chrono>
#include <vector>

#include "fl_internal.h"
#include "ipm_kernels.cuh"

namespace {

#define FLI_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fl_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return 0; } } while (0)

template <class T> T* dalloc(size_t count, std::vector<void*>& owned)
{
    void* p = nullptr;
    if (fl_dev_malloc(&p, (count ? count : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, (count ? count : 1) * sizeof(T));
    owned.push_back(p);
    return static_cast<T*>(p);
}

template <class T> T* dupload(const std::vector<T>& v, std::vector<void*>& owned)
{
    T* p = dalloc<T>(v.size(), owned);
    if (p && !v.empty()) cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}

} // namespace

struct fl_ipm
{
    fl_nlp* nlp = nullptr;
    int B = 0, NV = 0, NEQ = 0, NB = 0;
    fl_ipm_dev I{};
    fl_kkt_dev K{};
    std::vector<void*> owned;
    double *wx = nullptr, *wy = nullptr;          // solution of one solve
    int* d_neg = nullptr; int* d_zero = nullptr;
    long launches = 0;
    long n_factor_launches = 0, n_solve_launches = 0, n_eval_full = 0, n_eval_fg = 0, n_rounds = 0;
    double ms_factor = 0.0, ms_solve = 0.0, ms_eval = 0.0, ms_total = 0.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // partitioned (parallel-in-the-stages) mode
    fl_kkt_part Q{};
    int P = 0;                                   // separators (0: the mode is not available for this problem)
    int vec_block = IPM_BLOCK;                   // threads per lap of the vector kernels (IPM_BLOCK_MAX for small batches)
    size_t vec_smem = (size_t)IPM_NCPEMAX * IPM_BLOCK * sizeof(double);
    int vec_cluster = 1;                         // blocks per lap of the vector kernels (a thread-block cluster when > 1: long laps, few of them)
    int* d_list[2] = { nullptr, nullptr };       // running laps of this round / of the round before (compact launches of the vector kernels)
    int n_list = 0;                              // blocks of a compact launch (0: one block per lap of the batch)
    int vec_block_created = IPM_BLOCK;           // the choice made for the whole batch; vec_block follows the number of laps still running
    int part_policy = 0;                         // 0 auto (few active laps), 1 always, -1 never
    int part_max_laps = 512;
    bool part_round = false;                     // the current factorisation is a partitioned one
    double *rDinv = nullptr, *rF = nullptr, *rLst = nullptr, *rAq = nullptr, *rywork = nullptr;
    int *d_redneg = nullptr, *d_redzero = nullptr;
    long n_part_factor = 0;
    bool warp_kkt = false;                       // the warp-per-lap schedule of kkt_warp_kernels.cuh serves the unpartitioned rounds
    bool warp_kkt_ok = false;                    // ... and is available for this variant
};

namespace {

// The whole chain of every lap: one warp per lap with the stage matrix in registers (kkt_warp_kernels.cuh) -- the batch schedule --
// or one block of four warps per lap (kkt_kernels.cuh), which also serves the reduced system of the partitioned mode.
template <int NV, int NEQ> void launch_factor(fl_ipm* s, const fl_kkt_dev& K)
{
    const int nb = K.lap_list ? s->n_list : s->B;          // blocks: one per lap of the batch, or one per running lap (compact launch)
    if (s->warp_kkt && !K.sep) k_kkt_factor_w<NV, NEQ><<<nb, 32, sizeof(kw_smem<NV, NEQ>), s->nlp->stream>>>(K);
    else k_kkt_factor<NV, NEQ><<<nb, KT, sizeof(kkt_smem<NV, NEQ>), s->nlp->stream>>>(K);
}
template <int NV, int NEQ> void launch_solve(fl_ipm* s, const fl_kkt_dev& K, const double* r1, const double* r2, double* dx, double* dy)
{
    const int nb = K.lap_list ? s->n_list : s->B;
    if (s->warp_kkt && !K.sep)
    {
        // the ring version (factor blocks through shared memory by bulk asynchronous copies) where the stage's blocks are multiples of
        // 16 bytes; FL_KWS_RING=0 keeps the register-prefetch kernel (comparison runs)
        static const bool ring = !(std::getenv("FL_KWS_RING") && std::atoi(std::getenv("FL_KWS_RING")) == 0);
        if constexpr (kws_ring<NV, NEQ>::ok)
        {
            if (ring)
            {
                static unsigned long long attr_set = 0;          // per device
                if (!((attr_set >> s->nlp->device) & 1ull))
                {
                    cudaFuncSetAttribute(k_kkt_solve_w2<NV, NEQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kws_ring<NV, NEQ>::bytes);
                    // (the largest shared-memory carve-out, so that the ring, not the default L1 / shared split, decides the laps per SM)
                    cudaFuncSetAttribute(k_kkt_solve_w2<NV, NEQ>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
                    attr_set |= 1ull << s->nlp->device;
                }
                k_kkt_solve_w2<NV, NEQ><<<nb, 32, kws_ring<NV, NEQ>::bytes, s->nlp->stream>>>(K, r1, r2, dx, dy);
                return;
            }
        }
        k_kkt_solve_w<NV, NEQ><<<nb, 32, 0, s->nlp->stream>>>(K, r1, r2, dx, dy);
    }
    else k_kkt_solve<NV, NEQ><<<nb, KT, 0, s->nlp->stream>>>(K, r1, r2, dx, dy);
}

#define FLI_DISPATCH(fn, ...) \
    do { \
        if (s->NV == 15 && s->NEQ == 13) fn<15, 13>(s, __VA_ARGS__); \
        else if (s->NV == 17 && s->NEQ == 15) fn<17, 15>(s, __VA_ARGS__); \
        else if (s->NV == 17 && s->NEQ == 14) fn<17, 14>(s, __VA_ARGS__); \
        else if (s->NV == 14 && s->NEQ == 12) fn<14, 12>(s, __VA_ARGS__); \
        else if (s->NV == 16 && s->NEQ == 14) fn<16, 14>(s, __VA_ARGS__); \
        else if (s->NV == 8 && s->NEQ == 6) fn<8, 6>(s, __VA_ARGS__); \
        else if (s->NV == 13 && s->NEQ == 11) fn<13, 11>(s, __VA_ARGS__); \
        else { fl_set_error("no KKT kernel for this stage size"); return 0; } \
    } while (0)

template <int NV, int NEQ> void launch_factor_seg(fl_ipm* s, const fl_kkt_dev& K)
{
    k_kkt_factor_seg<NV, NEQ><<<dim3(s->P, K.lap_list ? s->n_list : s->B), KT, sizeof(kkt_smem<NV, NEQ>), s->nlp->stream>>>(K, s->Q);
}
template <int NV, int NEQ> void launch_solve_seg_fwd(fl_ipm* s, const fl_kkt_dev& K, const double* r1, const double* r2)
{
    k_kkt_solve_seg_fwd<NV, NEQ><<<dim3(s->P, K.lap_list ? s->n_list : s->B), KT, 0, s->nlp->stream>>>(K, s->Q, r1, r2);
}
template <int NV, int NEQ> void launch_solve_seg_bwd(fl_ipm* s, const fl_kkt_dev& K, const double* r2, double* dx, double* dy)
{
    k_kkt_solve_seg_bwd<NV, NEQ><<<dim3(s->P, K.lap_list ? s->n_list : s->B), KT, 0, s->nlp->stream>>>(K, s->Q, s->rywork, r2, dx, dy);
}

// the warp-per-lap factorisation keeps ~20 KB of shared memory per lap: ask for the largest shared-memory carve-out, so that
// the registers, not the default L1 / shared split, decide how many laps an SM holds
template <int NV, int NEQ> void carveout_w(fl_ipm*, int* ok)
{
    *ok = cudaFuncSetAttribute(k_kkt_factor_w<NV, NEQ>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared) == cudaSuccess;
}
int set_warp_carveout(fl_ipm* s)
{
    int ok = 0;
    FLI_DISPATCH(carveout_w, &ok);
    if (!ok) { fl_set_error(std::string("cudaFuncSetAttribute(PreferredSharedMemoryCarveout) failed: ") + cudaGetErrorString(cudaGetLastError())); return 0; }
    return 1;
}

// vector kernels: one block per lap, or one cluster of `vec_cluster` blocks per lap
template <class... KArgs, class... Args>
static void vec_launch(fl_ipm* s, void (*kernel)(KArgs...), Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((s->I.lap_list ? s->n_list : s->B) * s->vec_cluster), 1, 1);
    cfg.blockDim = dim3((unsigned)s->vec_block, 1, 1);
    cfg.dynamicSmemBytes = s->vec_smem;
    cfg.stream = s->nlp->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)s->vec_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->vec_cluster > 1 ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// FL_IPM_PROFILE=1: synchronise around the phases and accumulate their device time (diagnosis only)
bool profiling() { static const bool on = std::getenv("FL_IPM_PROFILE") != nullptr; return on; }
struct phase_timer
{
    fl_ipm* s; double* acc; cudaEvent_t a, b;
    phase_timer(fl_ipm* s_, double* acc_) : s(s_), acc(acc_) { if (profiling()) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s->nlp->stream); } }
    ~phase_timer() { if (profiling()) { cudaEventRecord(b, s->nlp->stream); cudaEventSynchronize(b); float ms = 0.f; cudaEventElapsedTime(&ms, a, b); *acc += ms; cudaEventDestroy(a); cudaEventDestroy(b); } }
};

// with_ic: the inertia-correcting loop runs inside the kernel on the lap records (the interior-point iteration);
// otherwise one pass with the dw / dc arrays, pivot counts reported
int kkt_factor(fl_ipm* s, const int* active, int ls_mode, bool with_ic)
{
    phase_timer pt(s, &s->ms_factor);
    fl_kkt_dev K = s->K;
    K.active = active; K.ls_mode = ls_mode; K.laps = with_ic ? s->I.lap : nullptr; K.o = s->I.o; K.lap_list = s->I.lap_list;
    FLI_DISPATCH(launch_factor, K);
    s->launches += 1; s->n_factor_launches += 1;
    FLI_CUDA(cudaGetLastError());
    return 1;
}

// the reduced system over the separators, as the sequential kernels see it
fl_kkt_dev reduced_system(const fl_ipm* s, const fl_kkt_dev& K)
{
    fl_kkt_dev R = K;
    R.sep = s->Q.sep; R.P = s->P; R.redP = s->Q.redP; R.redT = s->Q.redT; R.redL = s->Q.redL; R.segcarry = s->Q.segcarry; R.segrb = s->Q.segrb;
    R.Dinv = s->rDinv; R.F = s->rF; R.Lst = s->rLst; R.Aq = s->rAq; R.ywork = s->rywork; R.n_neg = s->d_redneg; R.n_zero = s->d_redzero;
    return R;
}

// one pass of the partitioned factorisation with the dw / dc arrays (pivot counts per segment and of the reduced system)
int kkt_factor_part(fl_ipm* s, const int* active, int ls_mode)
{
    phase_timer pt(s, &s->ms_factor);
    fl_kkt_dev K = s->K;
    K.active = active; K.ls_mode = ls_mode; K.laps = nullptr; K.o = s->I.o; K.lap_list = s->I.lap_list;
    FLI_DISPATCH(launch_factor_seg, K);
    const fl_kkt_dev R = reduced_system(s, K);
    FLI_DISPATCH(launch_factor, R);
    s->launches += 2; s->n_factor_launches += 1; s->n_part_factor += 1;
    FLI_CUDA(cudaGetLastError());
    return 1;
}

int kkt_solve(fl_ipm* s, const int* active, int ls_mode, const double* r1, const double* r2, double* dx, double* dy)
{
    phase_timer pt(s, &s->ms_solve);
    fl_kkt_dev K = s->K;
    K.active = active; K.ls_mode = ls_mode; K.laps = nullptr; K.lap_list = s->I.lap_list;
    if (s->part_round)
    {
        FLI_DISPATCH(launch_solve_seg_fwd, K, r1, r2);
        const fl_kkt_dev R = reduced_system(s, K);
        FLI_DISPATCH(launch_solve, R, r1, r2, dx, dy);
        FLI_DISPATCH(launch_solve_seg_bwd, K, r2, dx, dy);
        s->launches += 3; s->n_solve_launches += 1;
        FLI_CUDA(cudaGetLastError());
        return 1;
    }
    FLI_DISPATCH(launch_solve, K, r1, r2, dx, dy);
    s->launches += 1; s->n_solve_launches += 1;
    FLI_CUDA(cudaGetLastError());
    return 1;
}

// solve K (tx, ty) = (I.r1, I.r2) with iterative refinement on the full augmented system; result into (ox, oy)
int solve_refined(fl_ipm* s, const int* active, int ls_mode, double* ox, double* oy, int refine_steps = -1)
{
    const fl_ipm_dev& I = s->I;
    cudaStream_t st = s->nlp->stream;
    if (!kkt_solve(s, active, ls_mode, I.r1, I.r2, s->wx, s->wy)) return 0;
    vec_launch(s, k_ipm_axpy_sol, I, active, s->wx, s->wy, I.tx, I.ty, 1);
    if (refine_steps < 0) refine_steps = I.o.refine_steps;
    for (int k = 0; k < refine_steps; ++k)
    {
        // the refinement solve runs for the laps whose residual asks for it (Ipopt's residual_ratio_max; Ipopt always takes
        // one step, min_refinement_steps = 1: FL_IPM_REFINE_ALWAYS=1 does the same)
        static const double ratio_max = std::getenv("FL_IPM_REFINE_ALWAYS") ? -1.0 : 1.0e-10;
        vec_launch(s, k_ipm_residual, I, active, ls_mode, ratio_max);
        if (!kkt_solve(s, I.act_refine, ls_mode, I.e1, I.e2, s->wx, s->wy)) return 0;
        vec_launch(s, k_ipm_axpy_sol, I, (const int*)I.act_refine, s->wx, s->wy, I.tx, I.ty, 0);
        s->launches += 2;
    }
    vec_launch(s, k_ipm_axpy_sol, I, active, I.tx, I.ty, ox, oy, 1);
    s->launches += 2;
    FLI_CUDA(cudaGetLastError());
    return 1;
}

int read_counters(fl_ipm* s, int* host4)
{
    FLI_CUDA(cudaMemcpyAsync(host4, s->I.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->nlp->stream));
    FLI_CUDA(cudaStreamSynchronize(s->nlp->stream));
    return 1;
}

int zero_counters(fl_ipm* s)
{
    FLI_CUDA(cudaMemsetAsync(s->I.counters, 0, 8 * sizeof(int), s->nlp->stream));
    return 1;
}

} // namespace

// Releases the solver's device resources and detaches it from its problem handle (called by fl_ipm_destroy, and by
// fl_nlp_destroy for the solvers still bound to a handle that is going away).
void fl_ipm_orphan(fl_ipm* s)
{
    if (!s || !s->nlp) return;
    fl_nlp* nlp = s->nlp;
    cudaSetDevice(nlp->device);
    if (nlp->stream) cudaStreamSynchronize(nlp->stream);
    for (void* p : s->owned) fl_dev_free(p);
    s->owned.clear();
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    s->ev0 = s->ev1 = nullptr;
    for (size_t k = 0; k < nlp->dependents.size(); ++k)
        if (nlp->dependents[k] == s) { nlp->dependents.erase(nlp->dependents.begin() + (long)k); break; }
    s->nlp = nullptr;
}

#define FLI_ALIVE(s) do { if (!(s) || !(s)->nlp) { fl_set_error("the solver's problem handle has been destroyed"); return 0; } } while (0)

extern "C" {

void fl_ipm_default_options(fl_ipm_opts* o)
{
    // Ipopt defaults, with the tolerances the reference sets (optimal_laptime.hpp:528-541)
    o->tol = 1.0e-10; o->constr_viol_tol = 1.0e-10; o->acceptable_tol = 1.0e-8; o->acceptable_iter = 15; o->max_iter = 3000;
    o->mu_init = 0.1; o->bound_push = 1.0e-2; o->bound_frac = 1.0e-2; o->bound_relax_factor = 1.0e-8;
    o->kappa_eps = 10.0; o->kappa_mu = 0.2; o->theta_mu = 1.5; o->tau_min = 0.99; o->kappa_sigma = 1.0e10; o->s_max = 100.0;
    o->gamma_theta = 1.0e-5; o->gamma_phi = 1.0e-8; o->delta = 1.0; o->s_theta = 1.1; o->s_phi = 2.3; o->eta_phi = 1.0e-8;
    o->kappa_soc = 0.99; o->alpha_red = 0.5; o->max_soc = 4; o->max_backtracks = 40;
    o->delta_w_min = 1.0e-20; o->delta_w_0 = 1.0e-4; o->delta_w_max = 1.0e40; o->kappa_w_minus = 1.0 / 3.0; o->kappa_w_plus = 8.0;
    o->kappa_w_plus_bar = 100.0; o->delta_c_bar = 1.0e-8; o->kappa_c = 0.25; o->y_init_max = 1.0e3; o->refine_steps = 1;
}

void fl_ipm_destroy(fl_ipm* s)
{
    if (!s) return;
    fl_ipm_orphan(s);
    delete s;
}

fl_ipm* fl_ipm_create(fl_nlp* nlp, const fl_ipm_opts* opts, const double* x_l, const double* x_u, const double* g_l, const double* g_u)
{
    return fl_ipm_create_ex(nlp, opts, x_l, x_u, g_l, g_u, 0);
}

// bounds_per_lap != 0: x_l, x_u are [batch][n] and g_l, g_u [batch][m] -- one set per problem of the batch (G-G sweeps: the
// reference narrows the bounds of the acceleration per speed, steady_state.hpp:733-735, 787-789)
fl_ipm* fl_ipm_create_ex(fl_nlp* nlp, const fl_ipm_opts* opts, const double* x_l, const double* x_u, const double* g_l, const double* g_u, int bounds_per_lap)
{
    if (!nlp) { fl_set_error("null nlp"); return nullptr; }
    if (cudaSetDevice(nlp->device) != cudaSuccess) { fl_set_error("cudaSetDevice failed"); return nullptr; }
    static_assert(sizeof(fl_ipm_opts) == sizeof(fl_ipm_options), "public and device option structs must match");
    const fl_variant* v = nlp->v;
    const fl_dev& P = nlp->P;
    if (v->NV > IPM_NVMAX || v->NCPE > IPM_NCPEMAX) { fl_set_error("variant too large for the interior-point kernels"); return nullptr; }
    fl_ipm* s = new fl_ipm;
    s->nlp = nlp;
    s->B = P.batch;
    fl_ipm_opts o;
    if (opts) o = *opts; else fl_ipm_default_options(&o);
    const int NV = v->NV, NCPE = v->NCPE, NINEQ = v->NEXT, NEQ = NCPE - NINEQ;
    s->NV = NV; s->NEQ = NEQ; s->NB = NV + NEQ;
    const int n = P.n, m = P.m, S = P.n_var_nodes, n_el = P.n_elements, ni = n_el * NINEQ, B = P.batch;
    // bounds: the caller's, or the model defaults; relaxed by bound_relax_factor as Ipopt does
    const bool have_bounds = x_l && x_u && g_l && g_u;
    if (bounds_per_lap && !have_bounds) { fl_set_error("bounds per lap need x_l, x_u, g_l, g_u"); delete s; return nullptr; }
    const int NBND = bounds_per_lap ? B : 1;         // sets of bounds
    std::vector<double> xl((size_t)NBND * n), xu((size_t)NBND * n), gl((size_t)NBND * m), gu((size_t)NBND * m);
    if (have_bounds) { xl.assign(x_l, x_l + xl.size()); xu.assign(x_u, x_u + xu.size()); gl.assign(g_l, g_l + gl.size()); gu.assign(g_u, g_u + gu.size()); }
    else if (!fl_nlp_bounds(nlp, xl.data(), xu.data(), gl.data(), gu.data())) { delete s; return nullptr; }
    // (Ipopt also caps this relaxation at constr_viol_tol, 1e-10 with the reference's options; the cap is not applied here --
    //  listed with the deviations in ipm_kernels.cuh: a converged iterate may sit up to 1e-8 max(1, |b|) outside a bound)
    auto rel = [&](double b) { return o.bound_relax_factor * std::fmax(1.0, std::fabs(b)); };
    // row classification: the NEXT "extra" rows of every element are the inequalities (optimal_laptime.hpp:447-476)
    std::vector<int> eq_of(NCPE, -1), ineq_of(NCPE, -1), eq_rows, ineq_rows;
    for (int r = 0; r < NCPE; ++r)
    {
        const bool ineq = r >= v->NTD + v->NALG && r < v->NTD + v->NALG + v->NEXT;
        if (ineq) { ineq_of[r] = (int)ineq_rows.size(); ineq_rows.push_back(r); }
        else { eq_of[r] = (int)eq_rows.size(); eq_rows.push_back(r); }
    }
    for (size_t i = 0; i < gl.size(); ++i)
    {
        const bool ineq = ineq_of[(i % (size_t)m) % NCPE] >= 0;
        if (ineq != (gu[i] > gl[i])) { fl_set_error("constraint bounds do not match the row classification of the variant"); delete s; return nullptr; }
    }
    std::vector<double> sl((size_t)NBND * ni), su((size_t)NBND * ni);
    for (int b = 0; b < NBND; ++b)
        for (int e = 0; e < n_el; ++e)
            for (int t = 0; t < NINEQ; ++t)
            {
                const size_t row = (size_t)b * m + (size_t)e * NCPE + ineq_rows[t];
                sl[(size_t)b * ni + e * NINEQ + t] = gl[row] - rel(gl[row]);
                su[(size_t)b * ni + e * NINEQ + t] = gu[row] + rel(gu[row]);
            }
    for (size_t i = 0; i < xl.size(); ++i)
    {
        if (!(xu[i] > xl[i])) { fl_set_error("every variable needs lower < upper bound"); delete s; return nullptr; }
        xl[i] -= rel(xl[i]); xu[i] += rel(xu[i]);
    }
    auto& own = s->owned;
    fl_ipm_dev& I = s->I;
    I.n = n; I.m = m; I.ni = ni; I.S = S; I.NV = NV; I.NCPE = NCPE; I.NINEQ = NINEQ; I.NEQ = NEQ; I.n_el = n_el; I.node0 = P.node0; I.n_points = P.n_points;
    I.NJA = v->NJA; I.NJB = v->NJB; I.NH = v->NH; I.NCPL = v->NCPL; I.nB = P.nB; I.nC = P.nC; I.closed = P.closed;
    I.nnz_jac = P.nnz_jac; I.nnz_hess = P.nnz_hess;
    auto ivec = [&](const int* p, int cnt) { return dupload(std::vector<int>(p, p + cnt), own); };
    I.t.ja_row = ivec(v->JA_ROW, v->NJA); I.t.ja_col = ivec(v->JA_COL, v->NJA);
    I.t.jb_row = ivec(v->JB_ROW, v->NJB); I.t.jb_col = ivec(v->JB_COL, v->NJB);
    I.t.h_row = ivec(v->H_ROW, v->NH); I.t.h_col = ivec(v->H_COL, v->NH);
    I.t.ctrl_pos = ivec(v->CTRL_POS, 2);
    I.t.eq_of_row = dupload(eq_of, own); I.t.ineq_of_row = dupload(ineq_of, own);
    I.t.eq_rows = dupload(eq_rows, own); I.t.ineq_rows = dupload(ineq_rows, own);
    std::memcpy(&I.o, &o, sizeof(o));
    I.nx = nlp->d_xuser; I.nlam = nlp->d_lambda; I.nf = nlp->d_f; I.ng = nlp->d_g; I.ngrad = nlp->d_grad; I.njac = nlp->d_jac; I.nhess = nlp->d_hess;
    I.xl = dupload(xl, own); I.xu = dupload(xu, own); I.gl = dupload(gl, own); I.sl = dupload(sl, own); I.su = dupload(su, own);
    I.bn = bounds_per_lap ? (size_t)n : 0; I.bm = bounds_per_lap ? (size_t)m : 0; I.bi = bounds_per_lap ? (size_t)ni : 0;
    const size_t Bn = (size_t)B * n, Bm = (size_t)B * m, Bi = (size_t)B * ni;
    I.x = dalloc<double>(Bn, own); I.s = dalloc<double>(Bi, own); I.y = dalloc<double>(Bm, own);
    I.zl = dalloc<double>(Bn, own); I.zu = dalloc<double>(Bn, own); I.vl = dalloc<double>(Bi, own); I.vu = dalloc<double>(Bi, own);
    I.dx = dalloc<double>(Bn, own); I.ds = dalloc<double>(Bi, own); I.dy = dalloc<double>(Bm, own);
    I.dxs = dalloc<double>(Bn, own); I.dss = dalloc<double>(Bi, own); I.dys = dalloc<double>(Bm, own); I.st = dalloc<double>(Bi, own);
    I.sigx = dalloc<double>(Bn, own); I.sigs = dalloc<double>(Bi, own); I.r1 = dalloc<double>(Bn, own); I.r2 = dalloc<double>(Bm, own);
    I.c = dalloc<double>(Bm, own); I.csoc = dalloc<double>(Bm, own); I.rsbar = dalloc<double>(Bi, own); I.gx = dalloc<double>(Bn, own);
    I.e1 = dalloc<double>(Bn, own); I.e2 = dalloc<double>(Bm, own); I.tx = dalloc<double>(Bn, own); I.ty = dalloc<double>(Bm, own);
    s->wx = dalloc<double>(Bn, own); s->wy = dalloc<double>(Bm, own);
    I.dw = dalloc<double>(B, own); I.dc = dalloc<double>(B, own);
    I.act_factor = dalloc<int>(B, own); I.act_soc = dalloc<int>(B, own); I.act_run = dalloc<int>(B, own);
    I.act_refine = dalloc<int>(B, own);
    I.counters = dalloc<int>(8, own);
    s->d_neg = dalloc<int>(B, own); s->d_zero = dalloc<int>(B, own);
    I.n_neg = s->d_neg; I.n_zero = s->d_zero;
    I.lap = dalloc<fl_ipm_lap>(B, own);
    s->d_list[0] = dalloc<int>(B, own); s->d_list[1] = dalloc<int>(B, own);
    I.lap_list = nullptr; I.lap_list_out = nullptr;
    fl_kkt_dev& K = s->K;
    K.S = S; K.n_el = n_el; K.node0 = P.node0; K.closed = P.closed; K.nB = P.nB; K.nC = P.nC; K.n_points = P.n_points;
    K.NJA = v->NJA; K.NJB = v->NJB; K.NH = v->NH; K.NCPL = v->NCPL; K.NCPE = NCPE; K.NINEQ = NINEQ; K.n = n; K.m = m;
    K.nnz_jac = P.nnz_jac; K.nnz_hess = P.nnz_hess; K.t = I.t;
    K.jac = nlp->d_jac; K.hess = nlp->d_hess; K.sigx = I.sigx; K.sigs = I.sigs; K.dw = I.dw; K.dc = I.dc; K.active = nullptr;
    const size_t NB2 = (size_t)s->NB * s->NB;
    K.Dinv = dalloc<double>((size_t)B * S * NB2, own);
    K.F = (P.closed && S > 2) ? dalloc<double>((size_t)B * S * NB2, own) : nullptr;
    K.Aq = dalloc<double>((size_t)B * S * NINEQ * (NV + 1), own);
    K.ywork = dalloc<double>((size_t)B * S * s->NB, own);
    K.Lst = dalloc<double>((size_t)B * S * s->NB * NV, own);
    K.laps = nullptr;
    K.n_neg = s->d_neg; K.n_zero = s->d_zero; K.ls_mode = 0;
    if (!I.lap || !K.Dinv || !K.ywork || !K.Lst || !s->wy || (P.closed && S > 2 && !K.F))
    {
        fl_set_error("out of device memory for the interior-point workspace");
        fl_ipm_destroy(s);
        return nullptr;
    }
    // partitioned mode: about sqrt(S) separators
    if (S >= 64)
    {
        int Pn = (int)std::lround(std::sqrt((double)S));
        Pn = Pn < 3 ? 3 : (Pn > 128 ? 128 : Pn);
        std::vector<int> sep;
        for (int k = 0; k < Pn; ++k) { const int v0 = (int)((long)k * S / Pn); if (sep.empty() || v0 > sep.back()) sep.push_back(v0); }
        s->P = (int)sep.size();
        const size_t BP = (size_t)B * s->P;
        s->Q.sep = dupload(sep, own); s->Q.P = s->P;
        s->Q.redP = dalloc<double>(BP * NV * NV, own); s->Q.redT = dalloc<double>(BP * NB2, own); s->Q.redL = dalloc<double>(BP * s->NB * NV, own);
        s->Q.segcarry = dalloc<double>(BP * NV, own); s->Q.segrb = dalloc<double>(BP * s->NB, own);
        s->Q.segneg = dalloc<int>(BP, own); s->Q.segzero = dalloc<int>(BP, own);
        s->rDinv = dalloc<double>(BP * NB2, own); s->rF = dalloc<double>(BP * NB2, own); s->rLst = dalloc<double>(BP * s->NB * NV, own);
        s->rAq = dalloc<double>(BP * NINEQ * (NV + 1), own); s->rywork = dalloc<double>(BP * s->NB, own);
        s->d_redneg = dalloc<int>(B, own); s->d_redzero = dalloc<int>(B, own);
        if (!K.F) K.F = dalloc<double>((size_t)B * S * NB2, own);          // open laps: the segments' fill blocks
        if (!s->rywork || !s->rDinv || !K.F) { fl_set_error("out of device memory for the partitioned factorisation"); fl_ipm_destroy(s); return nullptr; }
        // automatic schedule: parallel in the stages while at most this many laps are running.  Measured on the 4096-lap sweep of 500-node
        // laps (first pass, s): 512 laps 12.99, 1024 12.71, 1536 12.58, 2048 12.81, 2560 13.10 -- the warp-per-lap schedule needs its second
        // wave from 1629 laps on and pays the slowest lap's inertia-correction attempts below that; scaled with the stages of a lap
        s->part_max_laps = std::max(256, std::min(2048, 768000 / std::max(1, S)));
        if (const char* e = std::getenv("FL_IPM_PART_MAX")) s->part_max_laps = std::atoi(e);
    }
    // vector kernels: one block per lap; wide blocks when the batch alone cannot fill the machine
    if (S == 1) s->vec_block = 32;                // one-node problems (steady-state NLPs): a warp per problem, its reductions need no barrier tree
    else if (B <= 64) s->vec_block = IPM_BLOCK_MAX;
    else if (B <= 512) s->vec_block = 256;
    s->vec_smem = (size_t)IPM_NCPEMAX * s->vec_block * sizeof(double);
    s->vec_block_created = s->vec_block;
    // long laps, few of them: a cluster of blocks per lap (each block strides over 1 / vec_cluster of the lap's vectors)
    if (const char* e = std::getenv("FL_IPM_CLUSTER")) s->vec_cluster = std::max(1, std::min(8, std::atoi(e)));
    else
    {
        int n_sm = 0;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, nlp->device) != cudaSuccess || n_sm <= 0) { fl_set_error("cannot read the SM count of the device"); fl_ipm_destroy(s); return nullptr; }
        while (s->vec_cluster < 8 && (size_t)B * s->vec_cluster * 2 <= (size_t)n_sm && S >= s->vec_cluster * 3 * s->vec_block / 2) s->vec_cluster *= 2;
    }
    {
        // The attribute is per device and process-wide: it is set to the largest request any solver can make (never to this
        // solver's own size, which would lower the limit under a live solver created for a smaller batch).
        const void* kernels[] = { (const void*)k_ipm_init, (const void*)k_ipm_init_slacks, (const void*)k_ipm_take_multipliers, (const void*)k_ipm_state,
                                  (const void*)k_ipm_rhs, (const void*)k_ipm_residual, (const void*)k_ipm_axpy_sol, (const void*)k_ipm_direction,
                                  (const void*)k_ipm_soc_direction, (const void*)k_ipm_line_search, (const void*)k_ipm_init_warm, (const void*)k_ipm_init_slacks_warm };
        const int smem_max = (int)((size_t)IPM_NCPEMAX * IPM_BLOCK_MAX * sizeof(double));
        for (const void* kf : kernels)
            if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) != cudaSuccess)
            {
                fl_set_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ") + cudaGetErrorString(cudaGetLastError()));
                fl_ipm_destroy(s);
                return nullptr;
            }
    }
    // warp-per-lap KKT kernels: the control couplings must sit on the last two variables of a node (direct form), and the compact
    // fill block holds 16 rows.  FL_KKT_WARP=0 keeps the block-per-lap kernels (comparison runs).
    s->warp_kkt_ok = (v->NCPL == 0 || (v->NCPL == 2 && v->CTRL_POS[0] == NV - 2 && v->CTRL_POS[1] == NV - 1)) && NEQ + v->NCPL <= 16;
    // automatic choice: batches.  A lap alone (or a handful) is latency-bound on its chain of stages, where four warps per lap
    // (and the parallel-in-the-stages schedule) are the better mapping.  fl_ipm_set_warp_kkt / FL_KKT_WARP override it.
    s->warp_kkt = s->warp_kkt_ok && B >= 64;
    if (const char* e = std::getenv("FL_KKT_WARP")) s->warp_kkt = s->warp_kkt_ok && std::atoi(e) != 0;
    if (s->warp_kkt_ok && !set_warp_carveout(s)) { fl_ipm_destroy(s); return nullptr; }
    cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);
    if (P.closed && S == 2) { fl_set_error("closed laps need at least three nodes"); fl_ipm_destroy(s); return nullptr; }
    nlp->dependents.push_back(s);
    return s;
}

// Test / building-block entry: one factorisation and solve of the augmented system at the callbacks' current output.
// Host arrays: sigx [batch][n], sigs [batch][n_el*NINEQ] (element-major), dw, dc [batch], r1 [batch][n], r2 [batch][m].
int fl_kkt_factor_solve(fl_ipm* s, const double* x, const double* lambda, const double* sigx, const double* sigs, const double* dw, const double* dc,
                        const double* r1, const double* r2, int refine_steps, int ls_mode, double* dx, double* dy, int* n_neg, int* n_zero)
{
    FLI_ALIVE(s);
    fl_nlp* nlp = s->nlp;
    FLI_CUDA(cudaSetDevice(nlp->device));
    const fl_ipm_dev& I = s->I;
    const size_t B = s->B;
    cudaStream_t st = nlp->stream;
    s->I.lap_list = nullptr; s->I.lap_list_out = nullptr; s->n_list = 0;          // one block per lap of the batch
    FLI_CUDA(cudaMemcpyAsync(nlp->d_xuser, x, B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(nlp->d_lambda, lambda, B * I.m * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!fl_eval_device(nlp, FL_EVAL_FG | FL_EVAL_JAC | FL_EVAL_HESS, 1.0)) return 0;
    FLI_CUDA(cudaMemcpyAsync(I.sigx, sigx, B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(I.sigs, sigs, B * I.ni * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(I.dw, dw, B * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(I.dc, dc, B * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(I.r1, r1, B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
    FLI_CUDA(cudaMemcpyAsync(I.r2, r2, B * I.m * sizeof(double), cudaMemcpyHostToDevice, st));
    // the refinement kernel reads dw, dc from the lap records
    std::vector<fl_ipm_lap> laps(B);
    for (size_t b = 0; b < B; ++b) { std::memset(&laps[b], 0, sizeof(fl_ipm_lap)); laps[b].dw = dw[b]; laps[b].dc = dc[b]; }
    FLI_CUDA(cudaMemcpyAsync(I.lap, laps.data(), B * sizeof(fl_ipm_lap), cudaMemcpyHostToDevice, st));
    s->part_round = s->P > 0 && s->part_policy > 0;
    if (s->part_round) { if (!kkt_factor_part(s, nullptr, ls_mode)) return 0; }
    else if (!kkt_factor(s, nullptr, ls_mode, false)) return 0;
    fl_ipm_options keep = s->I.o;
    s->I.o.refine_steps = refine_steps;
    const int ok = solve_refined(s, nullptr, ls_mode, I.dx, I.dy);
    s->I.o = keep;
    if (!ok) return 0;
    FLI_CUDA(cudaMemcpyAsync(dx, I.dx, B * I.n * sizeof(double), cudaMemcpyDeviceToHost, st));
    FLI_CUDA(cudaMemcpyAsync(dy, I.dy, B * I.m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (n_neg) FLI_CUDA(cudaMemcpyAsync(n_neg, s->part_round ? s->d_redneg : s->d_neg, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_zero) FLI_CUDA(cudaMemcpyAsync(n_zero, s->part_round ? s->d_redzero : s->d_zero, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    FLI_CUDA(cudaStreamSynchronize(st));
    if (s->part_round)
    {
        // add the pivots of the segment interiors
        std::vector<int> sn(B * s->P), sz(B * s->P);
        FLI_CUDA(cudaMemcpy(sn.data(), s->Q.segneg, sn.size() * sizeof(int), cudaMemcpyDeviceToHost));
        FLI_CUDA(cudaMemcpy(sz.data(), s->Q.segzero, sz.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < B; ++b)
            for (int k = 0; k < s->P; ++k) { if (n_neg) n_neg[b] += sn[b * s->P + k]; if (n_zero) n_zero[b] += sz[b * s->P + k]; }
        s->part_round = false;
    }
    return 1;
}

static int ipm_solve_impl(fl_ipm* s, const double* x0, const double* y0, const double* zl0, const double* zu0, double warm_push, double warm_mu, int verbose);

int fl_ipm_solve(fl_ipm* s, const double* x0, int verbose) { return ipm_solve_impl(s, x0, nullptr, nullptr, nullptr, 0.0, 0.0, verbose); }

// Warm start from a previous solution: x0 [batch][n], lambda0 [batch][m], zl0, zu0 [batch][n] (what the reference passes to
// ipopt_cppad_solve when warm_start is set, optimal_laptime.hpp:550-551).  push: Ipopt's warm_start_bound_push (1e-3);
// mu_init: starting barrier parameter (e.g. 1e-4 close to a solution).
int fl_ipm_solve_warm(fl_ipm* s, const double* x0, const double* lambda0, const double* zl0, const double* zu0, double push, double mu_init, int verbose)
{
    if (!lambda0 || !zl0 || !zu0 || !(push > 0.0) || !(mu_init > 0.0)) { fl_set_error("warm start needs lambda, zl, zu, push > 0 and mu_init > 0"); return 0; }
    return ipm_solve_impl(s, x0, lambda0, zl0, zu0, push, mu_init, verbose);
}

} // extern "C"

static int ipm_solve_impl(fl_ipm* s, const double* x0, const double* y0, const double* zl0, const double* zu0, double warm_push, double warm_mu, int verbose)
{
    FLI_ALIVE(s);
    fl_nlp* nlp = s->nlp;
    FLI_CUDA(cudaSetDevice(nlp->device));
    const fl_ipm_dev& I = s->I;
    const int B = s->B;
    cudaStream_t st = nlp->stream;
    const int ALL = FL_EVAL_FG | FL_EVAL_JAC | FL_EVAL_HESS;
    int cnt[4];
    // counters of THIS solve (fl_ipm_stats)
    s->launches = 0; s->n_factor_launches = 0; s->n_solve_launches = 0; s->n_eval_full = 0; s->n_eval_fg = 0; s->n_rounds = 0; s->n_part_factor = 0;
    s->ms_factor = s->ms_solve = s->ms_eval = 0.0;
    s->vec_block = s->vec_block_created;
    s->vec_smem = (size_t)IPM_NCPEMAX * s->vec_block * sizeof(double);
    s->I.lap_list = nullptr; s->I.lap_list_out = nullptr; s->n_list = 0;
    cudaEventRecord(s->ev0, st);
    FLI_CUDA(cudaMemcpyAsync(nlp->d_xuser, x0, (size_t)B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (y0)
    {
        FLI_CUDA(cudaMemcpyAsync(I.y, y0, (size_t)B * I.m * sizeof(double), cudaMemcpyHostToDevice, st));
        FLI_CUDA(cudaMemcpyAsync(I.zl, zl0, (size_t)B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
        FLI_CUDA(cudaMemcpyAsync(I.zu, zu0, (size_t)B * I.n * sizeof(double), cudaMemcpyHostToDevice, st));
        vec_launch(s, k_ipm_init_warm, I, warm_push, warm_mu);
        if (!fl_eval_device(nlp, FL_EVAL_FG, 1.0)) return 0;
        vec_launch(s, k_ipm_init_slacks_warm, I, warm_push);
        s->launches += 2; s->n_eval_fg += 1;
    }
    else
    {
        vec_launch(s, k_ipm_init, I);
        if (!fl_eval_device(nlp, ALL, 1.0)) return 0;
        vec_launch(s, k_ipm_init_slacks, I);
        s->launches += 2; s->n_eval_full += 1;
        // least-squares multipliers (Ipopt's default start): [I J^T; J -D][w; y] = -[grad f - zl + zu; 0]
        s->part_round = false;
        if (!kkt_factor(s, nullptr, 1, false)) return 0;
        if (!solve_refined(s, nullptr, 1, I.dx, I.dy)) return 0;
        vec_launch(s, k_ipm_take_multipliers, I);
        s->launches += 1;
    }
    const int blocks_b = (B + 127) / 128;
    // FL_IPM_TRACE=<file>: one line per lock-step round (running laps, schedule of the factorisation, host clock at the round's first
    // synchronisation point) -- where the time of a batch goes as its laps converge; and one comment line "#ls,round,trial rounds,laps
    // evaluated in them" per iteration (measured on 1024 laps: 18 trial rounds per iteration for 1.1 evaluations per lap -- a few laps
    // backtrack to the cap while the others wait; a round with few laps costs ~0.1 ms against 130 ms per iteration)
    std::FILE* trace = nullptr;
    if (const char* e = std::getenv("FL_IPM_TRACE")) { trace = std::fopen(e, "a"); if (trace) std::fprintf(trace, "round,running,partitioned,ms,ms_factor,ms_solve,ms_eval\n"); }
    const auto t_start = std::chrono::steady_clock::now();
    for (long round = 0;; ++round)
    {
        s->n_rounds = round;
        // callbacks at (x, y), state of the iterate
        { phase_timer pt(s, &s->ms_eval); if (!fl_eval_device_masked(nlp, ALL, 1.0, round == 0 ? nullptr : I.act_run)) return 0; }
        s->n_eval_full += 1;
        if (!zero_counters(s)) return 0;
        // (k_ipm_state runs over the laps listed in the round before -- a superset of those still running -- and the list of this
        // round is written into the other buffer)
        vec_launch(s, k_ipm_state, I);
        s->I.lap_list_out = s->d_list[round & 1];
        const bool no_compact = std::getenv("FL_IPM_NO_COMPACT") != nullptr;          // (comparison runs: tests/test_gpu_ipm.py)
        const bool listed = !no_compact && B >= 64 && I.S > 1;
        if (listed) { k_ipm_count_running<<<blocks_b, 128, 0, st>>>(I, B, 0); k_ipm_count_running<<<blocks_b, 128, 0, st>>>(I, B, 1); s->launches += 1; }
        else k_ipm_count_running<<<blocks_b, 128, 0, st>>>(I, B, -1);
        s->launches += 2;
        if (!read_counters(s, cnt)) return 0;
        // batches of laps: every kernel of the round is launched over the list of running laps, the costly ones first
        if (listed && cnt[0] > 0) { s->I.lap_list = s->d_list[round & 1]; s->n_list = cnt[0]; }
        else { s->I.lap_list = nullptr; s->n_list = 0; }
        if (verbose)
        {
            fl_ipm_lap l0;
            cudaMemcpy(&l0, I.lap, sizeof(fl_ipm_lap), cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "%4ld  running %d  lap0: it %d f %.10f inf_pr %.2e inf_du %.2e mu %.1e dw %.1e alpha %.2e status %d\n", round, cnt[0], l0.iter, l0.f,
                         l0.cviol, l0.dual_inf, l0.mu, l0.dw_last, l0.alpha, l0.status);
        }
        if (trace)
            std::fprintf(trace, "%ld,%d,%d,%.3f,%.3f,%.3f,%.3f\n", round, cnt[0], (int)(s->P > 0 && (s->part_policy > 0 || (s->part_policy == 0 && cnt[0] <= s->part_max_laps))),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(), s->ms_factor, s->ms_solve, s->ms_eval);
        if (cnt[0] == 0) break;
        // The vector kernels give one block to a lap.  Its width follows the laps still RUNNING: the stragglers of a large batch get the
        // wide blocks a small batch is created with (a 128-thread block walks a 500-node lap's vectors in 0.23 ms per line-search trial,
        // 0.7 - 1.2 ms per state / residual kernel, whatever the number of laps -- profiles/r2_solve1024_launches.md -- which is what a
        // round costs once few laps are left)
        if (I.S > 1 && !std::getenv("FL_IPM_FIXED_BLOCK"))
        {
            const int want = cnt[0] <= 64 ? IPM_BLOCK_MAX : (cnt[0] <= 512 ? 256 : IPM_BLOCK);
            s->vec_block = std::max(s->vec_block_created, want);
            s->vec_smem = (size_t)IPM_NCPEMAX * s->vec_block * sizeof(double);
        }
        // factorisation.  Many active laps: one block per lap walks the whole chain, the inertia correction (Algorithm IC)
        // runs inside the kernel.  Few active laps: the chain is cut into segments factorised by their own blocks, and IC
        // runs between the attempts.
        s->part_round = s->P > 0 && (s->part_policy > 0 || (s->part_policy == 0 && cnt[0] <= s->part_max_laps));
        if (!s->part_round) { if (!kkt_factor(s, I.act_run, 0, true)) return 0; }
        else
            for (int attempt = 0; attempt < 200; ++attempt)
            {
                if (!kkt_factor_part(s, I.act_factor, 0)) return 0;
                if (!zero_counters(s)) return 0;
                k_ipm_inertia_part<<<blocks_b, 128, 0, st>>>(I, s->Q, s->d_redneg, s->d_redzero, B);
                s->launches += 1;
                if (!read_counters(s, cnt)) return 0;
                if (cnt[1] == 0) break;
            }
        // search direction
        vec_launch(s, k_ipm_rhs, I, 0);
        s->launches += 1;
        if (!solve_refined(s, I.act_run, 0, I.dx, I.dy)) return 0;
        if (!zero_counters(s)) return 0;
        vec_launch(s, k_ipm_direction, I);
        s->launches += 1;
        if (!read_counters(s, cnt)) return 0;
        // filter line search, all laps in lock step
        int guard = 0;
        long ls_rounds = 0, ls_laps = 0;          // trace: trial rounds of this iteration and laps evaluated in them
        while (cnt[2] > 0 && guard++ < 400)
        {
            ls_rounds += 1; ls_laps += cnt[2];
            if (cnt[3] > 0)
            {
                vec_launch(s, k_ipm_rhs, I, 1);
                if (!solve_refined(s, I.act_soc, 0, I.dxs, I.dys, 0)) return 0;
                vec_launch(s, k_ipm_soc_direction, I);
                s->launches += 2;
            }
            k_ipm_mask_trials<<<blocks_b, 128, 0, st>>>(I, B);
            s->launches += 1;
            { phase_timer pt(s, &s->ms_eval); if (!fl_eval_device_masked(nlp, FL_EVAL_FG, 1.0, I.act_factor)) return 0; }
            s->n_eval_fg += 1;
            vec_launch(s, k_ipm_line_search, I);          // (its counters were zeroed by k_ipm_mask_trials)
            s->launches += 1;
            if (!read_counters(s, cnt)) return 0;
        }
        if (trace) std::fprintf(trace, "#ls,%ld,%ld,%ld\n", round, ls_rounds, ls_laps);
    }
    if (trace) std::fclose(trace);
    cudaEventRecord(s->ev1, st);
    FLI_CUDA(cudaEventSynchronize(s->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    s->ms_total = ms;
    if (profiling() && s->warp_kkt)
    {
        std::vector<fl_ipm_lap> laps(B);
        cudaMemcpy(laps.data(), I.lap, (size_t)B * sizeof(fl_ipm_lap), cudaMemcpyDeviceToHost);
        long long sf = 0, so = 0, sh = 0, nf = 0;
        for (const fl_ipm_lap& l : laps) { sf += l.stages_fail; so += l.stages_ok; sh += l.fail_shift; nf += l.n_fail; }
        std::fprintf(stderr, "ipm profile: warp-per-lap factorisation walked %lld stages in successful attempts and %lld in %lld failing ones (%.0f stages each of %d); "
                             "consecutive failures of a lap were %.1f stages apart on average\n", so, sf, nf, nf ? (double)sf / nf : 0.0, s->K.S, nf ? (double)sh / nf : 0.0);
    }
    if (profiling()) std::fprintf(stderr, "ipm profile: total %.1f ms, factor %.1f ms (%ld launches), solve %.1f ms (%ld), callbacks %.1f ms (%ld + %ld)\n", s->ms_total,
                                  s->ms_factor, s->n_factor_launches, s->ms_solve, s->n_solve_launches, s->ms_eval, s->n_eval_full, s->n_eval_fg);
    FLI_CUDA(cudaGetLastError());
    return 1;
}

extern "C" {

int fl_ipm_results(fl_ipm* s, double* x, double* y, double* zl, double* zu, double* obj, int* status, int* iters, double* errors)
{
    FLI_ALIVE(s);
    fl_nlp* nlp = s->nlp;
    FLI_CUDA(cudaSetDevice(nlp->device));
    const fl_ipm_dev& I = s->I;
    const size_t B = s->B;
    if (x) FLI_CUDA(cudaMemcpy(x, I.x, B * I.n * sizeof(double), cudaMemcpyDeviceToHost));
    if (y) FLI_CUDA(cudaMemcpy(y, I.y, B * I.m * sizeof(double), cudaMemcpyDeviceToHost));
    if (zl) FLI_CUDA(cudaMemcpy(zl, I.zl, B * I.n * sizeof(double), cudaMemcpyDeviceToHost));
    if (zu) FLI_CUDA(cudaMemcpy(zu, I.zu, B * I.n * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<fl_ipm_lap> laps(B);
    FLI_CUDA(cudaMemcpy(laps.data(), I.lap, B * sizeof(fl_ipm_lap), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; ++b)
    {
        if (obj) obj[b] = laps[b].f;
        if (status) status[b] = laps[b].status;
        if (iters) iters[b] = laps[b].iter;
        if (errors) { errors[4 * b] = laps[b].err0; errors[4 * b + 1] = laps[b].cviol; errors[4 * b + 2] = laps[b].dual_inf; errors[4 * b + 3] = laps[b].mu; }
    }
    return 1;
}

// Barrier terms of the iterates the solver holds (after a solve: of the returned points): sigx [batch][n] = zl / (x - xl) + zu / (xu - x),
// sigs [batch][n_elements * NINEQ] = vl / (s - sl) + vu / (su - s) with the solver's own (relaxed) bounds, slacks and slack multipliers,
// and the constraint regularisation dc [batch] of the last iteration.  With them fl_kkt_factor_solve solves systems with the KKT matrix
// of the solution -- parameter sensitivities (optimal_laptime.hpp:564-588).  Any output may be NULL.
int fl_ipm_sigma(fl_ipm* s, double* sigx, double* sigs, double* dc)
{
    FLI_ALIVE(s);
    FLI_CUDA(cudaSetDevice(s->nlp->device));
    FLI_CUDA(cudaStreamSynchronize(s->nlp->stream));
    const fl_ipm_dev& I = s->I;
    const size_t B = s->B, n = I.n, ni = I.ni;
    auto down = [&](const double* d, size_t count, std::vector<double>& h) { h.resize(count); return cudaMemcpy(h.data(), d, count * sizeof(double), cudaMemcpyDeviceToHost); };
    if (sigx)
    {
        std::vector<double> x, zl, zu, xl, xu;
        FLI_CUDA(down(I.x, B * n, x)); FLI_CUDA(down(I.zl, B * n, zl)); FLI_CUDA(down(I.zu, B * n, zu));
        FLI_CUDA(down(I.xl, (I.bn ? B : 1) * n, xl)); FLI_CUDA(down(I.xu, (I.bn ? B : 1) * n, xu));
        for (size_t b = 0; b < B; ++b)
            for (size_t i = 0; i < n; ++i)
            {
                const size_t k = b * n + i, kb = b * I.bn + i;
                sigx[k] = zl[k] / (x[k] - xl[kb]) + zu[k] / (xu[kb] - x[k]);
            }
    }
    if (sigs)
    {
        std::vector<double> sv, vl, vu, sl, su;
        FLI_CUDA(down(I.s, B * ni, sv)); FLI_CUDA(down(I.vl, B * ni, vl)); FLI_CUDA(down(I.vu, B * ni, vu));
        FLI_CUDA(down(I.sl, (I.bi ? B : 1) * ni, sl)); FLI_CUDA(down(I.su, (I.bi ? B : 1) * ni, su));
        for (size_t b = 0; b < B; ++b)
            for (size_t i = 0; i < ni; ++i)
            {
                const size_t k = b * ni + i, kb = b * I.bi + i;
                sigs[k] = vl[k] / (sv[k] - sl[kb]) + vu[k] / (su[kb] - sv[k]);
            }
    }
    if (dc)
    {
        std::vector<fl_ipm_lap> laps(B);
        FLI_CUDA(cudaMemcpy(laps.data(), I.lap, B * sizeof(fl_ipm_lap), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < B; ++b) dc[b] = laps[b].dc;
    }
    return 1;
}

// Partitioned (parallel-in-the-stages) factorisation: policy 0 = automatic (when at most `max_laps` laps are active), 1 = always,
// -1 = never.  Returns the number of separators (0: not available for this problem).
int fl_ipm_set_partitioned(fl_ipm* s, int policy, int max_laps)
{
    s->part_policy = policy;
    if (max_laps > 0) s->part_max_laps = max_laps;
    return s->P;
}

// Schedule of the unpartitioned KKT factorisation and solves: 1 = one warp per lap with the stage matrix in registers
// (kkt_warp_kernels.cuh), -1 = one block of four warps per lap (kkt_kernels.cuh), 0 = automatic (warp per lap for batches of
// 64 laps or more).  Returns 1 if the warp-per-lap kernels are in use afterwards.
int fl_ipm_set_warp_kkt(fl_ipm* s, int mode)
{
    if (mode > 0) s->warp_kkt = s->warp_kkt_ok;
    else if (mode < 0) s->warp_kkt = false;
    else s->warp_kkt = s->warp_kkt_ok && s->B >= 64;
    return s->warp_kkt ? 1 : 0;
}

// counters of the last solve: [0] kernel launches of the solver (callbacks excluded), [1] factorisation launches, [2] solve launches,
// [3] full callback rounds, [4] f/g-only callback rounds, [5] outer rounds; ms = device time of the whole solve
void fl_ipm_stats(const fl_ipm* s, long stats[6], double* ms)
{
    stats[0] = s->launches; stats[1] = s->n_factor_launches; stats[2] = s->n_solve_launches; stats[3] = s->n_eval_full; stats[4] = s->n_eval_fg;
    stats[5] = s->n_rounds;
    if (ms) *ms = s->ms_total;
}

} // extern "C"
