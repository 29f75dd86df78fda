// Parallel-in-the-stages variant of the KKT factorisation and solves (used when few laps are active: a single lap, or the
// last laps of a batch, are latency-bound on the sequential chain of kkt_kernels.cuh).
//
// The chain is cut at P separator stages.  The interior of every segment (a, b) between consecutive separators is
// eliminated backwards by its own block, independently of the other segments, carrying the fill G_j = K[b, j] towards its
// right separator exactly as the sequential kernel carries F_j = K[0, j] towards stage 0.  What is left is a
// block-(cyclic-)tridiagonal system over the separators whose couplings have the original shape (only the x columns of
// the left separator are non-zero); k_kkt_factor / k_kkt_solve of kkt_kernels.cuh run on it through their separator
// mapping (fl_kkt_dev::sep).  The inertia is the sum of the negative pivots of all blocks, interior and separator ones
// (Haynsworth), so Algorithm IC needs the counts of both kernels: it runs in k_ipm_inertia_part, between attempts.
// oracle/kkt.py::PartitionedLDL is the numpy statement of this file.
#pragma once
#include "kkt_kernels.cuh"

struct fl_kkt_part          // per-solver buffers of the partitioned mode (device pointers)
{
    const int* sep;          // [P]
    int P;
    double* redP;            // [batch][P][NV*NV]
    double* redT;            // [batch][P][NB*NB]
    double* redL;            // [batch][P][NB*NV]
    double* segcarry;        // [batch][P][NV]
    double* segrb;           // [batch][P][NB]
    int* segneg;             // [batch][P]
    int* segzero;            // [batch][P]
};

// first / last interior stage and right separator of segment k
__device__ __forceinline__ void kkt_segment(const fl_kkt_dev& K, const fl_kkt_part& Q, int k, int& a, int& hi, bool& has_b, int& bstage, int& kb)
{
    a = Q.sep[k];
    hi = (k + 1 < Q.P ? Q.sep[k + 1] : K.S) - 1;
    has_b = (k + 1 < Q.P) || K.closed;
    bstage = (k + 1 < Q.P) ? Q.sep[k + 1] : 0;
    kb = (k + 1) % Q.P;
}

template <int NV, int NEQ>
__global__ void __launch_bounds__(KT) k_kkt_factor_seg(const fl_kkt_dev K, const fl_kkt_part Q)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB, CB = St::CB, CV = St::CV;
    extern __shared__ __align__(16) unsigned char kkt_raw[];
    kkt_smem<NV, NEQ>& sm = *reinterpret_cast<kkt_smem<NV, NEQ>*>(kkt_raw);
    const int k = blockIdx.x, lap = K.lap_list ? K.lap_list[blockIdx.y] : blockIdx.y, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (K.active && !K.active[lap]) return;
    int a, hi, bstage, kb;
    bool has_b;
    kkt_segment(K, Q, k, a, hi, has_b, bstage, kb);
    const int S = K.S;
    double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    double* gF = K.F + (size_t)lap * S * NB * NB;
    double* gL = K.Lst + (size_t)lap * S * NB * NV;
    double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    const int AQ = K.NINEQ * (NV + 1);
    double* oP = Q.redP + ((size_t)lap * Q.P + k) * NV * NV;
    double* oT = Q.redT + ((size_t)lap * Q.P + kb) * NB * NB;
    double* oL = Q.redL + ((size_t)lap * Q.P + kb) * NB * NV;
    const double dw = K.dw[lap], dc = K.dc[lap];
    const int ncpl = K.ls_mode ? 0 : K.NCPL;
    int n_neg = 0, n_zero = 0;
    double tacc[CB];                              // (row lane, own columns) of the Schur term handed to the right separator
#pragma unroll
    for (int i = 0; i < CB; ++i) tacc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.P, lane, c) = 0.0; }
#pragma unroll
    for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(sm.F, lane, c) = 0.0; }
    __syncthreads();
    if (has_b)
    {
        // G_{b-1} = K[b, b-1] (for b = 0: the corner block K[0, S-1] of a closed lap)
        St::gather_coupling(K, lap, bstage, sm.L);
#pragma unroll
        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.F, lane, c) = KM(sm.L, lane, c); }
        __syncthreads();
    }
    for (int j = hi; j > a; --j)
    {
        St::gather_diag(K, lap, j, dw, dc, sm.M, sm.aq);
        for (int q = t; q < K.NINEQ * NV; q += KT) gA[(size_t)j * AQ + q] = sm.aq[(q / NV) * KLD + (q % NV)];
        if (t < K.NINEQ) gA[(size_t)j * AQ + K.NINEQ * NV + t] = sm.aq[K.NINEQ * KLD + t];
        if (lane < NV)
        {
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.M, lane, c) -= KM(sm.P, lane, c); }
        }
        __syncthreads();
        St::gather_coupling(K, lap, j, sm.L);
        if (lane < NB)
        {
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) gL[(size_t)j * NB * NV + lane + NB * c] = KM(sm.L, lane, c); }
        }
        const double* Mi = kkt_invert<NB>(sm.M, sm.M2, n_neg, n_zero);
        double* XF = (Mi == sm.M) ? sm.M2 : sm.M;
        if (lane < NB)
        {
#pragma unroll
            for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) gD[(size_t)j * NB * NB + lane + NB * c] = KM(Mi, lane, c); }
            if (has_b)
            {
#pragma unroll
                for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) gF[(size_t)j * NB * NB + lane + NB * c] = KM(sm.F, lane, c); }
            }
            // XL = Dinv L, XF = Dinv G^T (G has NV non-zero columns)
            double acc[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) acc[i] = 0.0;
#pragma unroll
            for (int q = NV; q < NB; ++q)
            {
                const double av = KM(Mi, lane, q);
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += av * KM(sm.L, q, c); }
            }
            for (int q = 0; q < ncpl; ++q)
            {
                const int cp = K.t.ctrl_pos[q];
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c == cp) acc[i] += KM(Mi, lane, cp) * KM(sm.L, cp, cp); }
            }
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.XL, lane, c) = acc[i]; }
            if (has_b)
            {
                double ac[CB];
#pragma unroll
                for (int i = 0; i < CB; ++i) ac[i] = 0.0;
#pragma unroll 5
                for (int q = 0; q < NV; ++q)
                {
                    const double av = KM(Mi, lane, q);
#pragma unroll
                    for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) ac[i] += av * KM(sm.F, c, q); }
                }
#pragma unroll
                for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(XF, lane, c) = ac[i]; }
            }
        }
        __syncthreads();
        if (lane < NV)
        {
            double acc[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) acc[i] = 0.0;
#pragma unroll
            for (int q = NV; q < NB; ++q)
            {
                const double av = KM(sm.L, q, lane);
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += av * KM(sm.XL, q, c); }
            }
            for (int q = 0; q < ncpl; ++q)
            {
                const int cp = K.t.ctrl_pos[q];
                if (lane == cp)
                {
                    const double av = KM(sm.L, cp, cp);
#pragma unroll
                    for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += av * KM(sm.XL, cp, c); }
                }
            }
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.P, lane, c) = acc[i]; }
        }
        double fn[CV];
#pragma unroll
        for (int i = 0; i < CV; ++i) fn[i] = 0.0;
        if (has_b && lane < NB)
        {
#pragma unroll 5
            for (int q = 0; q < NV; ++q)
            {
                const double av = KM(sm.F, lane, q);
#pragma unroll
                for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) tacc[i] += av * KM(XF, q, c); }
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) fn[i] -= av * KM(sm.XL, q, c); }
            }
        }
        __syncthreads();
        if (has_b)
        {
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.F, lane, c) = fn[i]; }
        }
        __syncthreads();
    }
    // hand-over to the reduced system
    if (lane < NV)
    {
#pragma unroll
        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) oP[lane + NV * c] = KM(sm.P, lane, c); }
    }
    if (has_b && lane < NB)
    {
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) oT[lane + NB * c] = tacc[i]; }
#pragma unroll
        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) oL[lane + NB * c] = KM(sm.F, lane, c); }
    }
    if (t == 0) { Q.segneg[(size_t)lap * Q.P + k] = n_neg; Q.segzero[(size_t)lap * Q.P + k] = n_zero; }
}

// condensed right-hand side entry t (< NB) of chain stage j
template <int NV, int NEQ>
__device__ __forceinline__ double kkt_stage_rhs(const fl_kkt_dev& K, const double* gA, const double* r1, const double* r2, int j, int t)
{
    constexpr int NB = NV + NEQ;
    const int AQ = K.NINEQ * (NV + 1);
    const int e = kkt_stage<NV, NEQ>::e_in(K, j);
    double b = 0.0;
    if (t < NV)
    {
        b = r1[(size_t)j * NV + t];
        for (int q = 0; q < K.NINEQ; ++q)
            b += gA[(size_t)j * AQ + q * NV + t] * gA[(size_t)j * AQ + K.NINEQ * NV + q] * r2[(size_t)e * K.NCPE + K.t.ineq_rows[q]];
    }
    else if (t < NB) b = r2[(size_t)e * K.NCPE + K.t.eq_rows[t - NV]];
    return b;
}

// forward substitution inside the segments: y_j = Dinv_j b_j, b_{j-1} -= L_j^T y_j, b_right -= G_j y_j
template <int NV, int NEQ>
__global__ void __launch_bounds__(KT) k_kkt_solve_seg_fwd(const fl_kkt_dev K, const fl_kkt_part Q, const double* __restrict__ r1, const double* __restrict__ r2)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB, CB = St::CB;
    __shared__ double vb[KLD], vy[KLD], part[3][KW][KLD];
    const int k = blockIdx.x, lap = K.lap_list ? K.lap_list[blockIdx.y] : blockIdx.y, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (K.active && !K.active[lap]) return;
    int a, hi, bstage, kb;
    bool has_b;
    kkt_segment(K, Q, k, a, hi, has_b, bstage, kb);
    const int S = K.S;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F + (size_t)lap * S * NB * NB;
    const double* gL = K.Lst + (size_t)lap * S * NB * NV;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    r1 += (size_t)lap * K.n; r2 += (size_t)lap * K.m;
    double carry = 0.0, rb = 0.0;
    for (int j = hi; j > a; --j)
    {
        if (t < NB) vb[t] = kkt_stage_rhs<NV, NEQ>(K, gA, r1, r2, j, t) + carry;
        __syncthreads();
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int q = w + KW * i; if (lane < NB && q < NB) acc += gD[(size_t)j * NB * NB + lane + NB * q] * vb[q]; }
        part[0][w][lane] = acc;
        __syncthreads();
        if (t < NB)
        {
            double y = 0.0;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) y += part[0][ww][t];
            vy[t] = y; yw[(size_t)j * NB + t] = y;
        }
        __syncthreads();
        double c1 = 0.0, c2 = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i)
        {
            const int q = w + KW * i;
            if (q < NB)
            {
                if (lane < NV) c1 += gL[(size_t)j * NB * NV + q + NB * lane] * vy[q];
                if (has_b && lane < NB && q < NV) c2 += gF[(size_t)j * NB * NB + lane + NB * q] * vy[q];
            }
        }
        part[1][w][lane] = c1; part[2][w][lane] = c2;
        __syncthreads();
        if (t < NB)
        {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) { s1 += part[1][ww][t]; s2 += part[2][ww][t]; }
            carry = (t < NV) ? -s1 : 0.0;
            rb -= s2;
        }
        __syncthreads();
    }
    if (t < NV) Q.segcarry[((size_t)lap * Q.P + k) * NV + t] = carry;
    if (has_b && t < NB) Q.segrb[((size_t)lap * Q.P + kb) * NB + t] = rb;
}

// back substitution inside the segments, from the separator solutions zred [batch][P][NB]; scatters dx, dy of the interior stages
template <int NV, int NEQ>
__global__ void __launch_bounds__(KT) k_kkt_solve_seg_bwd(const fl_kkt_dev K, const fl_kkt_part Q, const double* __restrict__ zred,
                                                           const double* __restrict__ r2, double* __restrict__ dx, double* __restrict__ dy)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB, CB = St::CB;
    __shared__ double vb[KLD], vy[KLD], zb[KLD], part[2][KW][KLD];
    const int k = blockIdx.x, lap = K.lap_list ? K.lap_list[blockIdx.y] : blockIdx.y, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (K.active && !K.active[lap]) return;
    int a, hi, bstage, kb;
    bool has_b;
    kkt_segment(K, Q, k, a, hi, has_b, bstage, kb);
    const int S = K.S;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F + (size_t)lap * S * NB * NB;
    const double* gL = K.Lst + (size_t)lap * S * NB * NV;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    const int AQ = K.NINEQ * (NV + 1);
    r2 += (size_t)lap * K.m; dx += (size_t)lap * K.n; dy += (size_t)lap * K.m;
    if (t < NB) { vy[t] = zred[((size_t)lap * Q.P + k) * NB + t]; zb[t] = has_b ? zred[((size_t)lap * Q.P + kb) * NB + t] : 0.0; }
    __syncthreads();
    for (int j = a + 1; j <= hi; ++j)
    {
        // w = L_j z_{j-1} + G_j^T z_b
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i)
        {
            const int c = w + KW * i;
            if (lane < NB)
            {
                if (c < NV) acc += gL[(size_t)j * NB * NV + lane + NB * c] * vy[c];
                if (has_b && lane < NV && c < NB) acc += gF[(size_t)j * NB * NB + c + NB * lane] * zb[c];
            }
        }
        part[0][w][lane] = acc;
        const double yj = (t < NB) ? yw[(size_t)j * NB + t] : 0.0;
        __syncthreads();
        if (t < NB)
        {
            double s1 = 0.0;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) s1 += part[0][ww][t];
            vb[t] = s1;
        }
        __syncthreads();
        double b = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int q = w + KW * i; if (lane < NB && q < NB) b += gD[(size_t)j * NB * NB + lane + NB * q] * vb[q]; }
        part[1][w][lane] = b;
        __syncthreads();
        if (t < NB)
        {
            double z = yj;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) z -= part[1][ww][t];
            vy[t] = z; yw[(size_t)j * NB + t] = z;
        }
        __syncthreads();
    }
    for (int j = a + 1 + w; j <= hi; j += KW)
    {
        const int e = St::e_in(K, j);
        const double z = (lane < NB) ? yw[(size_t)j * NB + lane] : 0.0;
        if (lane < NV) dx[(size_t)j * NV + lane] = z;
        else if (lane < NB) dy[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]] = z;
        for (int q = 0; q < K.NINEQ; ++q)
        {
            double av = (lane < NV) ? gA[(size_t)j * AQ + q * NV + lane] * z : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) av += __shfl_xor_sync(0xffffffffu, av, o);
            if (lane == 0)
            {
                const int row = e * K.NCPE + K.t.ineq_rows[q];
                dy[row] = (av - r2[row]) * gA[(size_t)j * AQ + K.NINEQ * NV + q];
            }
        }
    }
}
