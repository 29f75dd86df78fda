// Interior-point iteration for batches of independent collocation NLPs, device resident, sm_100a fp64.
//
// The reference hands its NLP to Ipopt (lion::ipopt_cppad_solve, src/core/applications/optimal_laptime.hpp:546-551, with
// the options of :528-541: tol 1e-10, constr_viol_tol 1e-10, acceptable_tol 1e-8, max_iter 3000, Ipopt defaults
// otherwise).  Ipopt is a third-party dependency outside the reference tree; these kernels implement its published
// algorithm (Waechter & Biegler, Math. Program. 106 (2006); equation numbers below are the paper's): slack
// reformulation of the inequality rows, monotone barrier update (7), fraction to the boundary (15), filter line search
// (18)-(20) with second-order correction, multiplier reset (16), inertia-correcting regularisation (Algorithm IC) on the
// exact inertia reported by the block factorisation of kkt_kernels.cuh.  oracle/ipm.py is the numpy statement of the
// same iteration.  Deviations from Ipopt: no NLP scaling; no restoration phase (when the line search finds no acceptable
// step, or after 8 consecutive steps shorter than 0.05, the barrier problem is restarted from the current primal point
// with a larger mu, an empty filter, bound multipliers on the central path, equality multipliers kept unless they have
// blown up, and a cap on the constraint violation tied to the current one; at most IPM_MAX_RESTARTS times, see
// ipm_restart); iterative refinement on demand (Ipopt's residual_ratio_max without its minimum of one step); Jacobian
// regularisation dc always on (closed laps with an even node count have dependent rows); bounds relaxed by
// bound_relax_factor max(1, |b|) = 1e-8 without Ipopt's cap at constr_viol_tol (1e-10 with the reference's options).
//
// One block per lap -- or, for long laps that would leave the GPU idle, one thread-block CLUSTER per lap (up to 8 blocks
// on neighbouring SMs): the blocks stride over the lap's vectors together, reductions and the leader's decisions cross
// the cluster through distributed shared memory and cluster barriers.  Every reduction is a fixed-order tree (results
// do not depend on scheduling).  The host only sequences phases and reads a few counters per phase: all laps advance in
// lock step, each with its own state.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <cooperative_groups.h>
#include "kkt_kernels.cuh"
#include "kkt_part_kernels.cuh"
#include "kkt_warp_kernels.cuh"

constexpr int IPM_BLOCK = 128;          // threads per lap when many laps share the GPU
constexpr int IPM_BLOCK_MAX = 1024;     // ... when few do (fl_ipm.cu picks per launch; kernels use blockDim.x)
constexpr int IPM_NVMAX = 17, IPM_NCPEMAX = 21, IPM_MAX_RESTARTS = 3;

struct fl_ipm_dev
{
    int n, m, ni, S, NV, NCPE, NINEQ, NEQ, n_el, node0, n_points, NJA, NJB, NH, NCPL, nB, nC, closed;
    long nnz_jac, nnz_hess;
    fl_kkt_tables t;
    fl_ipm_options o;
    // the NLP's buffers (fl_nlp): inputs x, lambda; outputs f, g, grad, jac, hess
    double* nx; double* nlam; const double* nf; const double* ng; const double* ngrad; const double* njac; const double* nhess;
    // bounds (relaxed), shared by the batch
    const double *xl, *xu, *gl, *sl, *su;          // [n], [n], [m], [ni], [ni] -- or one set per lap, see the strides
    size_t bn, bm, bi;                             // lap strides of the bound arrays: 0 (shared by the batch) or n, m, ni
    // iterate and work vectors, [batch][...]
    double *x, *s, *y, *zl, *zu, *vl, *vu;
    double *dx, *ds, *dy, *dxs, *dss, *dys, *st;   // step, second-order-corrected step, trial slacks
    double *sigx, *sigs, *r1, *r2, *c, *csoc, *rsbar, *gx, *e1, *e2, *tx, *ty;
    double *dw, *dc;                               // [batch] mirrors of lap.dw / lap.dc for the factorisation
    int* act_refine;                               // [batch] mask: the residual of the last solve asks for a refinement step
    int *act_factor, *act_soc, *act_run;           // [batch] masks: act_run = lap still iterating; act_soc = needs a second-order-correction solve;
                                                   // act_factor = pending factorisation (partitioned schedule), reused as "trial point pending" in the line search
    int* counters;                                 // [8]: 0 running laps, 1 pending factorisations, 2 laps in line search, 3 pending SOC solves
    const int* n_neg; const int* n_zero;
    fl_ipm_lap* lap;
    // compact launches: when few laps of a large batch are still running the vector kernels get one block per RUNNING lap, block b
    // serving lap lap_list[b] (written by k_ipm_count_running; null: block b serves lap b)
    const int* lap_list;
    int* lap_list_out;
};

// ---- the lap's thread group: one block, or the blocks of one cluster ---------------------------------------------------
__device__ __forceinline__ unsigned ipm_nranks() { unsigned v; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(v)); return v; }
__device__ __forceinline__ unsigned ipm_rank() { unsigned v; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v)); return v; }
// barrier over the lap's threads; global and shared-memory writes before it are visible after it
__device__ __forceinline__ void ipm_sync()
{
    if (ipm_nranks() > 1) cooperative_groups::this_cluster().sync();
    else __syncthreads();
}
// value the leader block (rank 0) left in its shared variable `sh` before this call
template <class T> __device__ __forceinline__ T ipm_bcast(T* sh)
{
    ipm_sync();
    const T v = (ipm_nranks() > 1) ? *cooperative_groups::this_cluster().map_shared_rank(sh, 0) : *sh;
    ipm_sync();          // the variable may be rewritten, the leader may exit
    return v;
}

struct ipm_verdict { double a, b; int code; };     // what the leader thread decided, for the rest of the lap's threads

// ---- reductions over the lap's threads (fixed order) --------------------------------------------------------------------
template <class OP> __device__ __forceinline__ double ipm_reduce(double v, double* red, OP op)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = op(r, red[w]);
    const unsigned nrk = ipm_nranks();
    if (nrk == 1) return r;
    __shared__ double part;
    if (threadIdx.x == 0) part = r;
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    cl.sync();
    r = *cl.map_shared_rank(&part, 0);
    for (unsigned q = 1; q < nrk; ++q) r = op(r, *cl.map_shared_rank(&part, q));
    cl.sync();
    return r;
}
__device__ __forceinline__ double ipm_block_sum(double v, double* red) { return ipm_reduce(v, red, [](double a, double b) { return a + b; }); }
__device__ __forceinline__ double ipm_block_max(double v, double* red) { return ipm_reduce(v, red, [](double a, double b) { return fmax(a, b); }); }
__device__ __forceinline__ double ipm_block_min(double v, double* red) { return -ipm_block_max(-v, red); }

__device__ __forceinline__ int ipm_row_of_ineq(const fl_ipm_dev& I, int k) { return (k / I.NINEQ) * I.NCPE + I.t.ineq_rows[k % I.NINEQ]; }

// ---- sparse products with the slot-major Jacobian / Hessian of one lap ------------------------------------------------
// The matrix slots and the vector multiplied are read through the non-coherent path (__ldg: neither is written by the
// calling kernel), so that the loads of several slots are in flight together.
// out[n] (+)= J^T w : thread per variable node, accumulators in shared memory (acc[NVMAX][(int)blockDim.x])
__device__ void ipm_JT_mul(const fl_ipm_dev& I, const double* jac, const double* w, double* out, bool accumulate, double* acc)
{
    const int NV = I.NV;
    for (int base = ipm_rank() * blockDim.x; base < I.S; base += ipm_nranks() * blockDim.x)
    {
        const int j = base + threadIdx.x;
        if (j < I.S)
        {
            for (int c = 0; c < NV; ++c) acc[c * (int)blockDim.x + threadIdx.x] = accumulate ? out[(size_t)j * NV + c] : 0.0;
            const int node = j + I.node0;
            const int e_in = node == 0 ? I.n_points - 1 : node - 1;
            // the element's entries of w, once, into a per-thread array: indexed by the (warp-uniform) row of slot k it is a
            // coalesced local-memory access, where w[e * NCPE + row] was a 32-sector gather per slot
            double wr[IPM_NCPEMAX];
            for (int r = 0; r < I.NCPE; ++r) wr[r] = __ldg(w + (size_t)e_in * I.NCPE + r);
#pragma unroll 8
            for (int k = 0; k < I.NJA; ++k)
                acc[I.t.ja_col[k] * (int)blockDim.x + threadIdx.x] += __ldg(jac + (size_t)k * I.n_el + e_in) * wr[I.t.ja_row[k]];
            const bool has_out = I.closed || node + 1 < I.n_points;
            if (has_out)
            {
                const int e_out = node, eb = e_out - I.node0;
                for (int r = 0; r < I.NCPE; ++r) wr[r] = __ldg(w + (size_t)e_out * I.NCPE + r);
#pragma unroll 8
                for (int k = 0; k < I.NJB; ++k)
                    acc[I.t.jb_col[k] * (int)blockDim.x + threadIdx.x] += __ldg(jac + (size_t)I.NJA * I.n_el + (size_t)k * I.nB + eb) * wr[I.t.jb_row[k]];
            }
            for (int c = 0; c < NV; ++c) out[(size_t)j * NV + c] = acc[c * (int)blockDim.x + threadIdx.x];
        }
    }
}

// out[m] = J v : thread per element
__device__ void ipm_J_mul(const fl_ipm_dev& I, const double* jac, const double* v, double* out, double* acc)
{
    const int NV = I.NV, NCPE = I.NCPE;
    for (int base = ipm_rank() * blockDim.x; base < I.n_el; base += ipm_nranks() * blockDim.x)
    {
        const int e = base + threadIdx.x;
        if (e < I.n_el)
        {
            for (int r = 0; r < NCPE; ++r) acc[r * (int)blockDim.x + threadIdx.x] = 0.0;
            const int nj = (e + 1 == I.n_points) ? 0 : e + 1;          // right node
            const int sj = nj - I.node0;
            double vr[IPM_NVMAX];
            for (int c = 0; c < NV; ++c) vr[c] = __ldg(v + (size_t)sj * NV + c);
#pragma unroll 8
            for (int k = 0; k < I.NJA; ++k)
                acc[I.t.ja_row[k] * (int)blockDim.x + threadIdx.x] += __ldg(jac + (size_t)k * I.n_el + e) * vr[I.t.ja_col[k]];
            const int eb = e - I.node0;
            if (eb >= 0)
            {
                for (int c = 0; c < NV; ++c) vr[c] = __ldg(v + (size_t)eb * NV + c);
#pragma unroll 8
                for (int k = 0; k < I.NJB; ++k)
                    acc[I.t.jb_row[k] * (int)blockDim.x + threadIdx.x] += __ldg(jac + (size_t)I.NJA * I.n_el + (size_t)k * I.nB + eb) * vr[I.t.jb_col[k]];
            }
            for (int r = 0; r < NCPE; ++r) out[(size_t)e * NCPE + r] = acc[r * (int)blockDim.x + threadIdx.x];
        }
    }
}

// out[n] += H v (symmetric, lower-triangle slots + control couplings); out must be initialised
__device__ void ipm_H_mul_add(const fl_ipm_dev& I, const double* hes, const double* v, double* out, double* acc)
{
    const int NV = I.NV;
    for (int base = ipm_rank() * blockDim.x; base < I.S; base += ipm_nranks() * blockDim.x)
    {
        const int j = base + threadIdx.x;
        if (j < I.S)
        {
            for (int c = 0; c < NV; ++c) acc[c * (int)blockDim.x + threadIdx.x] = out[(size_t)j * NV + c];
            double vr[IPM_NVMAX];
            for (int c = 0; c < NV; ++c) vr[c] = __ldg(v + (size_t)j * NV + c);
#pragma unroll 8
            for (int k = 0; k < I.NH; ++k)
            {
                const int r = I.t.h_row[k], c = I.t.h_col[k];
                const double h = __ldg(hes + (size_t)k * I.S + j);
                acc[r * (int)blockDim.x + threadIdx.x] += h * vr[c];
                if (r != c) acc[c * (int)blockDim.x + threadIdx.x] += h * vr[r];
            }
            // couplings with the previous node (element e_in) and with the next node (element e_out)
            const int node = j + I.node0;
            const int e_in = node == 0 ? I.n_points - 1 : node - 1;
            const int eb_in = e_in - I.node0;
            if (eb_in >= 0 && (I.closed || node > 0))
            {
                const int jp = (j == 0) ? I.S - 1 : j - 1;
                for (int q = 0; q < I.NCPL; ++q)
                {
                    const int cp = I.t.ctrl_pos[q];
                    acc[cp * (int)blockDim.x + threadIdx.x] += hes[(size_t)I.NH * I.S + (size_t)q * I.nC + eb_in] * v[(size_t)jp * NV + cp];
                }
            }
            const bool has_out = I.closed || node + 1 < I.n_points;
            if (has_out)
            {
                const int eb_out = node - I.node0;
                const int jn = (j + 1 == I.S) ? 0 : j + 1;
                for (int q = 0; q < I.NCPL; ++q)
                {
                    const int cp = I.t.ctrl_pos[q];
                    acc[cp * (int)blockDim.x + threadIdx.x] += hes[(size_t)I.NH * I.S + (size_t)q * I.nC + eb_out] * v[(size_t)jn * NV + cp];
                }
            }
            for (int c = 0; c < NV; ++c) out[(size_t)j * NV + c] = acc[c * (int)blockDim.x + threadIdx.x];
        }
    }
}

#define IPM_LAP_PROLOGUE \
    const int slot_ = blockIdx.x / ipm_nranks(); \
    const int lap = I.lap_list ? I.lap_list[slot_] : slot_; \
    const int t0 = ipm_rank() * blockDim.x + threadIdx.x, tstride = ipm_nranks() * blockDim.x;   /* this thread's share of the lap's vectors */ \
    fl_ipm_lap& L = I.lap[lap]; \
    const size_t on = (size_t)lap * I.n, om = (size_t)lap * I.m, oi = (size_t)lap * I.ni; \
    /* bounds: one set for the batch (strides 0) or one per lap */ \
    const double* const bxl = I.xl + (size_t)lap * I.bn; const double* const bxu = I.xu + (size_t)lap * I.bn; \
    const double* const bgl = I.gl + (size_t)lap * I.bm; \
    const double* const bsl = I.sl + (size_t)lap * I.bi; const double* const bsu = I.su + (size_t)lap * I.bi; \
    __shared__ double red[IPM_BLOCK_MAX / 32]; \
    (void)on; (void)om; (void)oi; (void)red; (void)t0; (void)tstride; (void)bxl; (void)bxu; (void)bgl; (void)bsl; (void)bsu;

__device__ __forceinline__ double ipm_push(double v, double lo, double hi, const fl_ipm_options& o)
{
    const double pl = fmin(o.bound_push * fmax(1.0, fabs(lo)), o.bound_frac * (hi - lo));
    const double pu = fmin(o.bound_push * fmax(1.0, fabs(hi)), o.bound_frac * (hi - lo));
    return fmin(fmax(v, lo + pl), hi - pu);
}

// ---- start ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_init(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    for (int i = t0; i < I.n; i += tstride)
    {
        const double v = ipm_push(I.nx[on + i], bxl[i], bxu[i], I.o);
        I.x[on + i] = v; I.nx[on + i] = v;
        I.zl[on + i] = 1.0; I.zu[on + i] = 1.0;
    }
    for (int i = t0; i < I.m; i += tstride) { I.y[om + i] = 0.0; I.nlam[om + i] = 0.0; }
    if (t0 == 0)
    {
        L.mu = I.o.mu_init; L.tau = fmax(I.o.tau_min, 1.0 - L.mu); L.dw = 0.0; L.dw_last = 0.0; L.dc = I.o.delta_c_bar * pow(L.mu, I.o.kappa_c);
        L.theta_max = -1.0; L.theta_min = -1.0; L.status = IPM_RUNNING; L.ls = LS_IDLE; L.iter = 0; L.acc_count = 0; L.n_filter = 0;
        L.n_backtrack = 0; L.n_soc = 0; L.n_reg = 0; L.n_fact = 0; L.switching = 0; L.n_restart = 0; L.n_small = 0; L.dw_prev = 0.0;
        L.fail_stage = -1; L.stages_fail = 0; L.stages_ok = 0; L.fail_shift = 0; L.n_fail = 0;
        I.dw[lap] = 0.0; I.dc[lap] = L.dc; I.act_run[lap] = 1; I.act_factor[lap] = 1; I.act_soc[lap] = 0;
    }
}

// Warm start (Ipopt's warm_start_init_point; the reference passes lambda, zl, zu of a previous solution,
// optimal_laptime.hpp:546-551): x is pushed inside its bounds by `push` only, the given multipliers are kept (bound
// multipliers at least `push`), the barrier parameter starts at mu0.  y, zl, zu have been copied into I.y, I.zl, I.zu.
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_init_warm(const fl_ipm_dev I, double push, double mu0)
{
    IPM_LAP_PROLOGUE
    fl_ipm_options o = I.o;
    o.bound_push = push; o.bound_frac = push;
    for (int i = t0; i < I.n; i += tstride)
    {
        const double v = ipm_push(I.nx[on + i], bxl[i], bxu[i], o);
        I.x[on + i] = v; I.nx[on + i] = v;
        I.zl[on + i] = fmax(I.zl[on + i], push); I.zu[on + i] = fmax(I.zu[on + i], push);
    }
    for (int i = t0; i < I.m; i += tstride) I.nlam[om + i] = I.y[om + i];
    if (t0 == 0)
    {
        L.mu = mu0; L.tau = fmax(I.o.tau_min, 1.0 - L.mu); L.dw = 0.0; L.dw_last = 0.0; L.dc = I.o.delta_c_bar * pow(L.mu, I.o.kappa_c);
        L.theta_max = -1.0; L.theta_min = -1.0; L.status = IPM_RUNNING; L.ls = LS_IDLE; L.iter = 0; L.acc_count = 0; L.n_filter = 0;
        L.n_backtrack = 0; L.n_soc = 0; L.n_reg = 0; L.n_fact = 0; L.switching = 0; L.n_restart = 0; L.n_small = 0; L.dw_prev = 0.0;
        L.fail_stage = -1; L.stages_fail = 0; L.stages_ok = 0; L.fail_shift = 0; L.n_fail = 0;
        I.dw[lap] = 0.0; I.dc[lap] = L.dc; I.act_run[lap] = 1; I.act_factor[lap] = 1; I.act_soc[lap] = 0;
    }
}

// slacks of a warm start: pushed by `push`; their bound multipliers from the sign of the row's multiplier (vu - vl = y)
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_init_slacks_warm(const fl_ipm_dev I, double push)
{
    IPM_LAP_PROLOGUE
    fl_ipm_options o = I.o;
    o.bound_push = push; o.bound_frac = push;
    for (int k = t0; k < I.ni; k += tstride)
    {
        const int row = ipm_row_of_ineq(I, k);
        I.s[oi + k] = ipm_push(I.ng[om + row], bsl[k], bsu[k], o);
        const double y = I.y[om + row];
        I.vl[oi + k] = fmax(-y, 0.0) + push; I.vu[oi + k] = fmax(y, 0.0) + push;
    }
}

// after the first evaluation: slacks pushed inside their bounds, multipliers 1; right-hand side of the least-squares
// multiplier system  [I J^T; J -D][w; y] = -[grad f - zl + zu; 0]
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_init_slacks(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    for (int k = t0; k < I.ni; k += tstride)
    {
        I.s[oi + k] = ipm_push(I.ng[om + ipm_row_of_ineq(I, k)], bsl[k], bsu[k], I.o);
        I.vl[oi + k] = 1.0; I.vu[oi + k] = 1.0;
    }
    for (int i = t0; i < I.n; i += tstride) I.r1[on + i] = -(I.ngrad[on + i] - I.zl[on + i] + I.zu[on + i]);
    for (int i = t0; i < I.m; i += tstride) I.r2[om + i] = 0.0;
}

__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_take_multipliers(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    double mx = 0.0;
    bool bad = false;
    for (int i = t0; i < I.m; i += tstride) { const double v = I.dy[om + i]; if (!isfinite(v)) bad = true; mx = fmax(mx, fabs(v)); }
    mx = ipm_block_max(bad ? 1.0e300 : mx, red);
    const bool take = mx <= I.o.y_init_max;
    for (int i = t0; i < I.m; i += tstride) { const double v = take ? I.dy[om + i] : 0.0; I.y[om + i] = v; I.nlam[om + i] = v; }
}

// ---- state of the current iterate: optimality error, barrier update, right-hand sides ----------------------------------
// Needs the callbacks evaluated at (x, y): f, g, grad f, Jacobian.
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_state(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    extern __shared__ double acc[];          // [IPM_NCPEMAX][blockDim.x]
    __shared__ ipm_verdict sh;
    if (L.status != IPM_RUNNING) return;
    const fl_ipm_options& o = I.o;
    const double* jac = I.njac + (size_t)lap * I.nnz_jac;
    // gx = grad f + J^T y
    for (int i = t0; i < I.n; i += tstride) I.gx[on + i] = I.ngrad[on + i];
    ipm_sync();
    ipm_JT_mul(I, jac, I.y + om, I.gx + on, true, acc);
    ipm_sync();
    // constraint residual c = g - gl (equality rows), g - s (inequality rows)
    double th = 0.0, cmax = 0.0, ysum = 0.0;
    bool bad = false;
    for (int i = t0; i < I.m; i += tstride)
    {
        const int r = i % I.NCPE;
        const int q = I.t.ineq_of_row[r];
        const double g = I.ng[om + i];
        const double c = (q < 0) ? g - bgl[i] : g - I.s[oi + (size_t)(i / I.NCPE) * I.NINEQ + q];
        I.c[om + i] = c;
        th += fabs(c); cmax = fmax(cmax, fabs(c)); ysum += fabs(I.y[om + i]);
        if (!isfinite(g)) bad = true;
    }
    th = ipm_block_sum(th, red); cmax = ipm_block_max(cmax, red); ysum = ipm_block_sum(ysum, red);
    // dual infeasibility, complementarity products, multiplier sums
    double dmax = 0.0, zsum = 0.0, pmin = 1.0e300, pmax = -1.0e300;
    for (int i = t0; i < I.n; i += tstride)
    {
        const double zl = I.zl[on + i], zu = I.zu[on + i], x = I.x[on + i];
        const double rx = I.gx[on + i] - zl + zu;
        dmax = fmax(dmax, fabs(rx)); zsum += zl + zu;
        const double p1 = (x - bxl[i]) * zl, p2 = (bxu[i] - x) * zu;
        pmin = fmin(pmin, fmin(p1, p2)); pmax = fmax(pmax, fmax(p1, p2));
        if (!isfinite(rx)) bad = true;
    }
    for (int k = t0; k < I.ni; k += tstride)
    {
        const double vl = I.vl[oi + k], vu = I.vu[oi + k], s = I.s[oi + k];
        const double rs = -I.y[om + ipm_row_of_ineq(I, k)] - vl + vu;
        dmax = fmax(dmax, fabs(rs)); zsum += vl + vu;
        const double p1 = (s - bsl[k]) * vl, p2 = (bsu[k] - s) * vu;
        pmin = fmin(pmin, fmin(p1, p2)); pmax = fmax(pmax, fmax(p1, p2));
    }
    dmax = ipm_block_max(dmax, red); zsum = ipm_block_sum(zsum, red);
    pmin = ipm_block_min(pmin, red); pmax = ipm_block_max(pmax, red);
    const double badf = ipm_block_max(bad ? 1.0 : 0.0, red);
    if (t0 == 0)
    {
        const double nz = 2.0 * I.n + 2.0 * I.ni;
        const double sd = fmax(o.s_max, (ysum + zsum) / (I.m + nz)) / o.s_max;
        const double sc = fmax(o.s_max, zsum / nz) / o.s_max;
        auto err = [&](double mu) { return fmax(fmax(dmax / sd, cmax), fmax(fabs(pmax - mu), fabs(pmin - mu)) / sc); };
        const double f = I.nf[lap];
        L.f = f; L.cviol = cmax; L.dual_inf = dmax; L.compl_inf = fmax(fabs(pmax), fabs(pmin)); L.err0 = err(0.0); L.theta = th;
        if (badf > 0.0 || !isfinite(f)) L.status = IPM_NOT_FINITE;
        else if (L.err0 <= o.tol && cmax <= o.constr_viol_tol) L.status = IPM_SUCCESS;
        else
        {
            L.acc_count = (L.err0 <= o.acceptable_tol) ? L.acc_count + 1 : 0;
            if (L.acc_count >= o.acceptable_iter) L.status = IPM_ACCEPTABLE;
            else if (L.iter >= o.max_iter) L.status = IPM_MAX_ITER;
        }
        if (L.status == IPM_RUNNING)
        {
            // barrier parameter update (7); several reductions may follow each other; the filter is reset
            bool changed = false;
            double mu = L.mu;
            while (mu > o.tol / 10.0 && err(mu) <= o.kappa_eps * mu) { mu = fmax(o.tol / 10.0, fmin(o.kappa_mu * mu, pow(mu, o.theta_mu))); changed = true; }
            if (changed) L.n_filter = 0;
            L.mu = mu; L.tau = fmax(o.tau_min, 1.0 - mu);
            if (L.theta_max < 0.0) { L.theta_max = 1.0e4 * fmax(1.0, th); L.theta_min = 1.0e-4 * fmax(1.0, th); }
            L.dw = 0.0; L.dc = o.delta_c_bar * pow(mu, o.kappa_c);
            I.dw[lap] = 0.0; I.dc[lap] = L.dc; I.act_factor[lap] = 1;
            L.ls = LS_IDLE;
        }
        else { I.act_run[lap] = 0; I.act_factor[lap] = 0; }
        sh.a = L.mu; sh.code = L.status;
    }
    const ipm_verdict v = ipm_bcast(&sh);
    if (v.code != IPM_RUNNING) return;
    const double mu = v.a;
    // Sigma, barrier function, gradient of the barrier Lagrangian
    double lg = 0.0;
    for (int i = t0; i < I.n; i += tstride)
    {
        const double x = I.x[on + i], dl = x - bxl[i], du = bxu[i] - x;
        const double zl = I.zl[on + i], zu = I.zu[on + i];
        I.sigx[on + i] = zl / dl + zu / du;
        lg += log(dl) + log(du);
        // r1 = -(grad phi_x + J^T y) = -(gx - mu/dl + mu/du)
        I.r1[on + i] = -(I.gx[on + i] - mu / dl + mu / du);
    }
    for (int k = t0; k < I.ni; k += tstride)
    {
        const double s = I.s[oi + k], dl = s - bsl[k], du = bsu[k] - s;
        I.sigs[oi + k] = I.vl[oi + k] / dl + I.vu[oi + k] / du;
        lg += log(dl) + log(du);
        I.rsbar[oi + k] = -mu / dl + mu / du - I.y[om + ipm_row_of_ineq(I, k)];
    }
    lg = ipm_block_sum(lg, red);
    if (t0 == 0) L.phi = L.f - mu * lg;
}

// ---- right-hand side of the augmented system: r2 = -c, inequality rows -= rs_bar / (Sigma_s + dw) ------------------------
// which = 0: c of the iterate, for every running lap; which = 1: the second-order-correction residual, for laps in LS_SOC_SOLVE
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_rhs(const fl_ipm_dev I, int which)
{
    IPM_LAP_PROLOGUE
    if (L.status != IPM_RUNNING) return;
    if (which == 1 && L.ls != LS_SOC_SOLVE) return;
    const double* cv = (which == 0 ? I.c : I.csoc) + om;
    const double dw = L.dw;
    for (int i = t0; i < I.m; i += tstride)
    {
        const int q = I.t.ineq_of_row[i % I.NCPE];
        double v = -cv[i];
        if (q >= 0)
        {
            const size_t k = oi + (size_t)(i / I.NCPE) * I.NINEQ + q;
            v -= I.rsbar[k] / (I.sigs[k] + dw);
        }
        I.r2[om + i] = v;
    }
}

// ---- residual of the full augmented system for iterative refinement: e = (r1, r2) - K (tx, ty) -----------------------------
// Also decides whether the lap needs the refinement solve at all (Ipopt's residual_ratio_max): act_refine[lap] =
// |e|_inf > ratio_max (|rhs|_inf + |solution|_inf).
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_residual(const fl_ipm_dev I, const int* active, int ls_mode, double ratio_max)
{
    IPM_LAP_PROLOGUE
    extern __shared__ double acc[];          // [IPM_NCPEMAX][blockDim.x]
    if (active && !active[lap]) { if (t0 == 0) I.act_refine[lap] = 0; return; }
    double emax = 0.0, rmax = 0.0, smax = 0.0;
    const double* jac = I.njac + (size_t)lap * I.nnz_jac;
    const double* hes = I.nhess + (size_t)lap * I.nnz_hess;
    const double dw = ls_mode ? 0.0 : L.dw, dc = L.dc;
    double* e1 = I.e1 + on; double* e2 = I.e2 + om;
    const double* tx = I.tx + on; const double* ty = I.ty + om;
    // e2 = r2 - (J tx - D ty)
    ipm_J_mul(I, jac, tx, e2, acc);
    ipm_sync();
    for (int i = t0; i < I.m; i += tstride)
    {
        const int q = I.t.ineq_of_row[i % I.NCPE];
        double D = dc;
        if (q >= 0)
        {
            const size_t k = oi + (size_t)(i / I.NCPE) * I.NINEQ + q;
            D = ls_mode ? 1.0 : 1.0 / (I.sigs[k] + dw) + dc;
        }
        const double r = I.r2[om + i], e = r - (e2[i] - D * ty[i]);
        e2[i] = e;
        emax = fmax(emax, fabs(e)); rmax = fmax(rmax, fabs(r)); smax = fmax(smax, fabs(ty[i]));
    }
    // e1 = r1 - (H tx + (sigx + dw) tx + J^T ty)
    for (int i = t0; i < I.n; i += tstride) e1[i] = ls_mode ? tx[i] : (I.sigx[on + i] + dw) * tx[i];
    ipm_sync();
    if (!ls_mode) ipm_H_mul_add(I, hes, tx, e1, acc);
    ipm_sync();
    ipm_JT_mul(I, jac, ty, e1, true, acc);
    ipm_sync();
    for (int i = t0; i < I.n; i += tstride)
    {
        const double r = I.r1[on + i], e = r - e1[i];
        e1[i] = e;
        emax = fmax(emax, fabs(e)); rmax = fmax(rmax, fabs(r)); smax = fmax(smax, fabs(tx[i]));
    }
    emax = ipm_block_max(emax, red); rmax = ipm_block_max(rmax, red); smax = ipm_block_max(smax, red);
    if (t0 == 0) I.act_refine[lap] = (emax > ratio_max * (rmax + smax) || !isfinite(emax)) ? 1 : 0;
}

// (tx, ty) += (dx, dy) of the refinement solve;  or initialise
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_axpy_sol(const fl_ipm_dev I, const int* active, const double* sx, const double* sy, double* tx, double* ty, int init)
{
    IPM_LAP_PROLOGUE
    if (active && !active[lap]) return;
    for (int i = t0; i < I.n; i += tstride) tx[on + i] = (init ? 0.0 : tx[on + i]) + sx[on + i];
    for (int i = t0; i < I.m; i += tstride) ty[om + i] = (init ? 0.0 : ty[om + i]) + sy[om + i];
}

// fraction to the boundary (15): largest alpha in (0, 1] with v + alpha dv >= lo + (1 - tau)(v - lo) and the same at the upper bound
__device__ __forceinline__ double ipm_ftb(double v, double dv, double lo, double hi, double tau)
{
    if (dv < 0.0) return -tau * (v - lo) / dv;
    if (dv > 0.0) return tau * (hi - v) / dv;
    return 1.0;
}

__device__ __forceinline__ void ipm_write_trial(const fl_ipm_dev& I, int lap, const double* dx, const double* ds, double alpha)
{
    const size_t on = (size_t)lap * I.n, oi = (size_t)lap * I.ni;
    const int t0 = ipm_rank() * blockDim.x + threadIdx.x, tstride = ipm_nranks() * blockDim.x;
    for (int i = t0; i < I.n; i += tstride) I.nx[on + i] = I.x[on + i] + alpha * dx[on + i];
    for (int k = t0; k < I.ni; k += tstride) I.st[oi + k] = I.s[oi + k] + alpha * ds[oi + k];
}

// ---- search direction: slacks, step lengths, directional derivative, first trial point --------------------------------
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_direction(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    if (L.status != IPM_RUNNING) return;
    const double mu = L.mu, tau = L.tau, dw = L.dw;
    double amax = 1.0, adual = 1.0, dphi = 0.0;
    bool bad = false;
    for (int k = t0; k < I.ni; k += tstride)
    {
        const double ds = (I.dy[om + ipm_row_of_ineq(I, k)] - I.rsbar[oi + k]) / (I.sigs[oi + k] + dw);
        I.ds[oi + k] = ds;
        const double s = I.s[oi + k], dl = s - bsl[k], du = bsu[k] - s, vl = I.vl[oi + k], vu = I.vu[oi + k];
        amax = fmin(amax, ipm_ftb(s, ds, bsl[k], bsu[k], tau));
        const double dvl = mu / dl - vl - vl / dl * ds, dvu = mu / du - vu + vu / du * ds;
        if (dvl < 0.0) adual = fmin(adual, -tau * vl / dvl);
        if (dvu < 0.0) adual = fmin(adual, -tau * vu / dvu);
        dphi += (-mu / dl + mu / du) * ds;
        if (!isfinite(ds)) bad = true;
    }
    for (int i = t0; i < I.n; i += tstride)
    {
        const double dx = I.dx[on + i], x = I.x[on + i], dl = x - bxl[i], du = bxu[i] - x, zl = I.zl[on + i], zu = I.zu[on + i];
        amax = fmin(amax, ipm_ftb(x, dx, bxl[i], bxu[i], tau));
        const double dzl = mu / dl - zl - zl / dl * dx, dzu = mu / du - zu + zu / du * dx;
        if (dzl < 0.0) adual = fmin(adual, -tau * zl / dzl);
        if (dzu < 0.0) adual = fmin(adual, -tau * zu / dzu);
        dphi += (I.ngrad[on + i] - mu / dl + mu / du) * dx;
        if (!isfinite(dx)) bad = true;
    }
    amax = ipm_block_min(amax, red); adual = ipm_block_min(adual, red); dphi = ipm_block_sum(dphi, red);
    const double badf = ipm_block_max(bad ? 1.0 : 0.0, red);
    if (t0 == 0)
    {
        L.alpha_max = amax; L.alpha = amax; L.alpha_dual = adual; L.dphi = dphi; L.n_backtrack = 0; L.n_soc = 0;
        if (badf > 0.0) { L.status = IPM_NOT_FINITE; I.act_run[lap] = 0; L.ls = LS_IDLE; }
        else { L.ls = LS_TRIAL; atomicAdd(I.counters + 2, 1); }
    }
    if (badf > 0.0) return;          // every thread of the lap holds the same reduced value
    ipm_write_trial(I, lap, I.dx, I.ds, amax);
}

// step length of a second-order-corrected step and its trial point (laps in LS_SOC_SOLVE, after the solve)
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_soc_direction(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    if (L.status != IPM_RUNNING || L.ls != LS_SOC_SOLVE) return;
    const double tau = L.tau, dw = L.dw;
    double amax = 1.0;
    for (int k = t0; k < I.ni; k += tstride)
    {
        const double ds = (I.dys[om + ipm_row_of_ineq(I, k)] - I.rsbar[oi + k]) / (I.sigs[oi + k] + dw);
        I.dss[oi + k] = ds;
        amax = fmin(amax, ipm_ftb(I.s[oi + k], ds, bsl[k], bsu[k], tau));
    }
    for (int i = t0; i < I.n; i += tstride) amax = fmin(amax, ipm_ftb(I.x[on + i], I.dxs[on + i], bxl[i], bxu[i], tau));
    amax = ipm_block_min(amax, red);
    if (t0 == 0) { L.alpha_soc = amax; L.ls = LS_SOC_TRIAL; }
    ipm_write_trial(I, lap, I.dxs, I.dss, amax);
}

// Restart of the barrier problem at the current primal point (stands in for Ipopt's restoration phase, which this solver
// does not have): larger mu, empty filter, bound multipliers on the central path, equality multipliers forgotten when they
// exceed y_init_max, theta_max tied to the current violation when the iterate is nearly feasible.  (Forgetting moderate
// multipliers and the loose theta_max of a fresh start threw a nearly converged 8000-node lap back to a violation of 1.5;
// with this form 4091 instead of 4057 laps of the 4096-lap sweep converge within 250 iterations.)
// Used when the line search finds no acceptable step, or when the iteration has crawled with tiny steps for a while
// (the multipliers then typically blow up).  At most IPM_MAX_RESTARTS times per lap.
__device__ void ipm_restart(const fl_ipm_dev& I, int lap, fl_ipm_lap& L, double* red)
{
    const double* const bxl = I.xl + (size_t)lap * I.bn; const double* const bxu = I.xu + (size_t)lap * I.bn;
    const double* const bsl = I.sl + (size_t)lap * I.bi; const double* const bsu = I.su + (size_t)lap * I.bi;
    const size_t on = (size_t)lap * I.n, om = (size_t)lap * I.m, oi = (size_t)lap * I.ni;
    __shared__ double sh_mu;
    const int t0 = ipm_rank() * blockDim.x + threadIdx.x, tstride = ipm_nranks() * blockDim.x;
    // the equality multipliers are forgotten only when they have blown up (the crawl this restart is meant to end)
    double ymax = 0.0;
    for (int i = t0; i < I.m; i += tstride) ymax = fmax(ymax, fabs(I.y[om + i]));
    ymax = ipm_block_max(ymax, red);
    const bool keep_y = ymax <= I.o.y_init_max;
    if (t0 == 0)
    {
        L.n_restart += 1;
        L.mu = fmin(I.o.mu_init, fmax(100.0 * L.mu, 1.0e-4));
        L.tau = fmax(I.o.tau_min, 1.0 - L.mu);
        L.n_filter = 0; L.acc_count = 0; L.n_small = 0; L.dw_last = 0.0; L.dw_prev = 0.0;
        // a nearly feasible iterate must stay nearly feasible: the restarted problem may not buy barrier decrease with
        // more than 100 times the current constraint violation
        L.theta_max = (L.theta > 1.0e-2) ? 1.0e4 * fmax(1.0, L.theta) : fmax(100.0 * L.theta, 1.0e-3);
        L.theta_min = 1.0e-4 * fmax(1.0, L.theta);
        sh_mu = L.mu;
    }
    const double mu = ipm_bcast(&sh_mu);
    for (int i = t0; i < I.n; i += tstride)
    {
        const double x = I.x[on + i];
        I.nx[on + i] = x;
        I.zl[on + i] = mu / (x - bxl[i]); I.zu[on + i] = mu / (bxu[i] - x);
    }
    for (int k = t0; k < I.ni; k += tstride)
    {
        const double s = I.s[oi + k];
        I.vl[oi + k] = mu / (s - bsl[k]); I.vu[oi + k] = mu / (bsu[k] - s);
    }
    if (!keep_y)
        for (int i = t0; i < I.m; i += tstride) { I.y[om + i] = 0.0; I.nlam[om + i] = 0.0; }
}

// ---- filter line search: judge the trial point (f, g evaluated at it) --------------------------------------------------
__global__ void __launch_bounds__(IPM_BLOCK_MAX) k_ipm_line_search(const fl_ipm_dev I)
{
    IPM_LAP_PROLOGUE
    __shared__ ipm_verdict sh;       // code: 0 backtrack, 1 accept, 2 start / continue SOC, 3 fail;  a: step length judged;  b: next one
    if (L.status != IPM_RUNNING || (L.ls != LS_TRIAL && L.ls != LS_SOC_TRIAL)) return;
    const fl_ipm_options& o = I.o;
    const bool soc = L.ls == LS_SOC_TRIAL;
    const double mu = L.mu;
    // theta and barrier function at the trial point; c_trial into e2 (scratch)
    double th = 0.0, lg = 0.0;
    bool bad = false;
    for (int i = t0; i < I.m; i += tstride)
    {
        const int q = I.t.ineq_of_row[i % I.NCPE];
        const double g = I.ng[om + i];
        const double c = (q < 0) ? g - bgl[i] : g - I.st[oi + (size_t)(i / I.NCPE) * I.NINEQ + q];
        I.e2[om + i] = c;
        th += fabs(c);
        if (!isfinite(g)) bad = true;
    }
    for (int i = t0; i < I.n; i += tstride) { const double x = I.nx[on + i]; lg += log(x - bxl[i]) + log(bxu[i] - x); }
    for (int k = t0; k < I.ni; k += tstride) { const double s = I.st[oi + k]; lg += log(s - bsl[k]) + log(bsu[k] - s); }
    th = ipm_block_sum(th, red); lg = ipm_block_sum(lg, red);
    const double badf = ipm_block_max(bad ? 1.0 : 0.0, red);
    if (t0 == 0)
    {
        const double ft = I.nf[lap];
        const double phi_t = ft - mu * lg;
        const double theta = L.theta, phi = L.phi, dphi = L.dphi;
        const double alpha = soc ? L.alpha_soc : L.alpha;
        bool okf = badf == 0.0 && isfinite(phi_t) && th <= L.theta_max;
        for (int k = 0; okf && k < L.n_filter; ++k)
            okf = th < (1.0 - o.gamma_theta) * L.filter[2 * k] || phi_t < L.filter[2 * k + 1] - o.gamma_phi * L.filter[2 * k];
        // switching condition (19) at the step length of the ordinary step
        const double a_sw = L.alpha;
        const bool switching = dphi < 0.0 && a_sw * pow(-dphi, o.s_phi) > o.delta * pow(theta, o.s_theta) && theta <= L.theta_min;
        const double a_arm = soc ? L.alpha_max : L.alpha;
        const bool armijo = phi_t <= phi + o.eta_phi * a_arm * dphi;
        bool accepted = false;
        if (okf) accepted = switching ? armijo : (th <= (1.0 - o.gamma_theta) * theta || phi_t <= phi - o.gamma_phi * theta);
        int decision = 0;
        if (accepted)
        {
            decision = 1;
            if (!switching || !armijo)
            {
                if (L.n_filter < IPM_FILTER_MAX) { L.filter[2 * L.n_filter] = theta; L.filter[2 * L.n_filter + 1] = phi; L.n_filter += 1; }
            }
            L.switching = switching;
        }
        else if (!soc)
        {
            if (L.n_backtrack == 0 && th >= theta && o.max_soc > 0) { decision = 2; L.theta_soc_old = th; L.n_soc = 0; }
        }
        else
        {
            L.n_soc += 1;
            if (th <= o.kappa_soc * L.theta_soc_old && L.n_soc < o.max_soc) { decision = 2; L.theta_soc_old = th; }
        }
        if (decision == 0)
        {
            // shorter ordinary step
            L.alpha = (soc ? L.alpha_max : L.alpha) * o.alpha_red;
            if (soc) L.alpha = L.alpha_max * o.alpha_red;
            L.n_backtrack += 1;
            if (L.alpha < 1.0e-13 || L.n_backtrack >= o.max_backtracks) decision = 3;
        }
        sh.code = decision; sh.a = alpha; sh.b = L.alpha;
    }
    const ipm_verdict v = ipm_bcast(&sh);
    const int decision = v.code;
    if (decision == 1)
    {
        // accept: new iterate, multiplier steps, reset (16)
        const double alpha = v.a, ad = L.alpha_dual, ks = o.kappa_sigma;
        const double* dxa = soc ? I.dxs : I.dx; const double* dsa = soc ? I.dss : I.ds; const double* dya = soc ? I.dys : I.dy;
        for (int i = t0; i < I.n; i += tstride)
        {
            const double x0 = I.x[on + i], dl0 = x0 - bxl[i], du0 = bxu[i] - x0, dx = I.dx[on + i];
            double zl = I.zl[on + i], zu = I.zu[on + i];
            zl += ad * (mu / dl0 - zl - zl / dl0 * dx);
            zu += ad * (mu / du0 - zu + zu / du0 * dx);
            const double x = x0 + alpha * dxa[on + i];
            const double dl = x - bxl[i], du = bxu[i] - x;
            I.x[on + i] = x; I.nx[on + i] = x;
            I.zl[on + i] = fmax(fmin(zl, ks * mu / dl), mu / (ks * dl));
            I.zu[on + i] = fmax(fmin(zu, ks * mu / du), mu / (ks * du));
        }
        for (int k = t0; k < I.ni; k += tstride)
        {
            const double s0 = I.s[oi + k], dl0 = s0 - bsl[k], du0 = bsu[k] - s0, ds = I.ds[oi + k];
            double vl = I.vl[oi + k], vu = I.vu[oi + k];
            vl += ad * (mu / dl0 - vl - vl / dl0 * ds);
            vu += ad * (mu / du0 - vu + vu / du0 * ds);
            const double s = s0 + alpha * dsa[oi + k];
            const double dl = s - bsl[k], du = bsu[k] - s;
            I.s[oi + k] = s;
            I.vl[oi + k] = fmax(fmin(vl, ks * mu / dl), mu / (ks * dl));
            I.vu[oi + k] = fmax(fmin(vu, ks * mu / du), mu / (ks * du));
        }
        for (int i = t0; i < I.m; i += tstride) { const double y = I.y[om + i] + alpha * dya[om + i]; I.y[om + i] = y; I.nlam[om + i] = y; }
        __shared__ int sh_restart;
        if (t0 == 0)
        {
            L.ls = LS_DONE; L.iter += 1; I.act_soc[lap] = 0;
            L.n_small = (alpha < 0.05) ? L.n_small + 1 : 0;
            sh_restart = (L.n_small >= 8 && L.n_restart < IPM_MAX_RESTARTS) ? 1 : 0;
        }
        if (ipm_bcast(&sh_restart)) ipm_restart(I, lap, L, red);
    }
    else if (decision == 2)
    {
        // second-order correction: c_soc = alpha c_prev + c(trial), a solve with that residual follows
        const double alpha = v.a;
        const double* prev = soc ? I.csoc : I.c;
        for (int i = t0; i < I.m; i += tstride) I.csoc[om + i] = alpha * prev[om + i] + I.e2[om + i];
        if (t0 == 0) { L.ls = LS_SOC_SOLVE; I.act_soc[lap] = 1; atomicAdd(I.counters + 2, 1); atomicAdd(I.counters + 3, 1); }
    }
    else if (decision == 0)
    {
        ipm_write_trial(I, lap, I.dx, I.ds, v.b);
        if (t0 == 0) { L.ls = LS_TRIAL; I.act_soc[lap] = 0; atomicAdd(I.counters + 2, 1); }
    }
    else
    {
        // The line search found no acceptable step: restart the barrier problem from the current iterate, or give up.
        __shared__ int sh_can;
        if (t0 == 0) { I.act_soc[lap] = 0; sh_can = L.n_restart < IPM_MAX_RESTARTS ? 1 : 0; }
        if (ipm_bcast(&sh_can))
        {
            ipm_restart(I, lap, L, red);
            if (t0 == 0) { L.ls = LS_DONE; L.iter += 1; }
        }
        else
        {
            for (int i = t0; i < I.n; i += tstride) I.nx[on + i] = I.x[on + i];
            if (t0 == 0) { L.status = IPM_LINE_SEARCH_FAILED; L.ls = LS_IDLE; I.act_run[lap] = 0; }
        }
    }
}

// Algorithm IC between the attempts of the partitioned factorisation: the inertia is the sum of the negative pivots of
// the segment blocks and of the reduced system.  Laps with a correct inertia leave the mask, the others get a larger dw.
__global__ void k_ipm_inertia_part(const fl_ipm_dev I, const fl_kkt_part Q, const int* redneg, const int* redzero, int batch)
{
    const int lap = blockIdx.x * blockDim.x + threadIdx.x;
    if (lap >= batch || !I.act_factor[lap]) return;
    fl_ipm_lap& L = I.lap[lap];
    const fl_ipm_options& o = I.o;
    int neg = redneg[lap], zero = redzero[lap];
    for (int k = 0; k < Q.P; ++k) { neg += Q.segneg[(size_t)lap * Q.P + k]; zero += Q.segzero[(size_t)lap * Q.P + k]; }
    L.n_fact += 1;
    if (neg == I.S * I.NEQ && zero == 0)
    {
        I.act_factor[lap] = 0;
        if (L.dw > 0.0) L.dw_last = L.dw;
        return;
    }
    L.n_reg += 1;
    if (L.dw == 0.0) L.dw = (L.dw_last == 0.0) ? o.delta_w_0 : fmax(o.delta_w_min, o.kappa_w_minus * L.dw_last);
    else L.dw *= (L.dw_last == 0.0) ? o.kappa_w_plus_bar : o.kappa_w_plus;
    if (L.dw > o.delta_w_max) { L.status = IPM_REGULARISATION_FAILED; I.act_factor[lap] = 0; I.act_run[lap] = 0; return; }
    I.dw[lap] = L.dw;
    atomicAdd(I.counters + 1, 1);
}

// mask of the laps whose trial point is waiting to be evaluated; also zeroes the counters the line-search kernel that follows the
// evaluation adds to (one launch less per trial round than a memset of its own)
__global__ void k_ipm_mask_trials(const fl_ipm_dev I, int batch)
{
    const int lap = blockIdx.x * blockDim.x + threadIdx.x;
    if (lap < 8) I.counters[lap] = 0;
    if (lap >= batch) return;
    const fl_ipm_lap& L = I.lap[lap];
    I.act_factor[lap] = (L.status == IPM_RUNNING && (L.ls == LS_TRIAL || L.ls == LS_SOC_TRIAL)) ? 1 : 0;
}

// Counts the running laps and lists them.  Two launches fill the list: pass 0 appends the laps whose last factorisation needed a
// regularisation (their first attempt at dw = 0 fails again more often than not: they stay longest in the inertia-correcting
// factorisation kernel), pass 1 the others -- blocks are dispatched in index order, so the long laps start first and the kernel does not
// end on them (longest processing time first).  Within a pass the order is whatever the atomics give: every lap's computation is its own.
__global__ void k_ipm_count_running(const fl_ipm_dev I, int batch, int pass)
{
    const int lap = blockIdx.x * blockDim.x + threadIdx.x;
    if (lap >= batch) return;
    const fl_ipm_lap& L = I.lap[lap];
    if (L.status == IPM_RUNNING)
    {
        const bool costly = L.dw_prev > 0.0;
        if (pass < 0 || costly == (pass == 0))
        {
            const int k = atomicAdd(I.counters + 0, 1);
            if (I.lap_list_out) I.lap_list_out[k] = lap;
        }
    }
    else if (pass <= 0) { I.act_refine[lap] = 0; I.act_soc[lap] = 0; }      // (compact launches have no block for a finished lap that could clear its masks)
}
