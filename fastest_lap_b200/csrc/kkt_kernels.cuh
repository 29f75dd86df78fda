// Batched block-(cyclic-)tridiagonal KKT factorisation and solves for the collocation NLP, sm_100a fp64.
//
// In the reference every interior-point iteration hands the augmented system to MUMPS inside Ipopt (third-party; the
// call site is lion::ipopt_cppad_solve, src/core/applications/optimal_laptime.hpp:546-551).  Here the chain structure
// of the collocation NLP (Hessian block diagonal per node + control couplings between neighbours,
// element rows touching two consecutive nodes) is used directly:
//
//        [ W + Sigma_x + dw I      J^T ] [dx]   [r1]
//        [ J                       -D  ] [dy] = [r2],     D = dc (equality rows),  1 / (Sigma_s + dw) + dc (inequality rows)
//
//   * the six inequality rows of an element touch its right node only: they are condensed into that node's block;
//   * unknowns are grouped per stage j = (dx of variable node j, dy of the equality rows of the element ENDING at it),
//     NB = NV + NEQ <= 32 unknowns; the matrix is symmetric block tridiagonal with one extra corner block for closed laps;
//   * stages are eliminated from the last one backwards (the order of a Riccati recursion: what stage j hands to stage
//     j-1 is the cost-to-go Hessian of node j-1), stage 0 last; on closed laps every stage also carries its fill block
//     F_j = K[0, j] towards stage 0;
//   * every pivot block is inverted in place by a Gauss-Jordan sweep with Bunch-Parlett pivoting (1x1 pivots on the
//     largest diagonal, 2x2 pivots when the diagonal is weak), which also counts the negative pivots: by Haynsworth
//     additivity their sum is the inertia Ipopt asks of its linear solver (n positive, m negative, none zero).
//
// One warp per lap: lane r owns row r of every stage matrix (shared memory, column-major with leading dimension 32 so a
// warp's access to a column is one conflict-free 256-byte row of banks; operands shared by all lanes are broadcast reads),
// products are accumulated in registers, reductions are warp shuffles.  Laps are independent: a batch fills the machine
// with one warp per lap, several laps per SM.  oracle/kkt.py is the numpy statement of the same algorithm.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct fl_kkt_tables          // device pointers, per variant
{
    const int *ja_row, *ja_col, *jb_row, *jb_col, *h_row, *h_col, *ctrl_pos;
    const int *eq_of_row;     // [NCPE] index among the equality rows of an element, or -1
    const int *ineq_of_row;   // [NCPE] index among the inequality rows, or -1
    const int *eq_rows;       // [NEQ]  row offsets of the equality rows
    const int *ineq_rows;     // [NINEQ]
};

struct fl_kkt_dev
{
    int S;                    // stages = variable nodes
    int n_el, node0, closed, nB, nC, n_points;
    int NJA, NJB, NH, NCPL, NCPE, NINEQ;
    int n, m;
    long nnz_jac, nnz_hess;
    fl_kkt_tables t;
    const double* jac;        // [batch][nnz_jac]   slot-major, see nlp_kernels.cuh
    const double* hess;       // [batch][nnz_hess]
    const double* sigx;       // [batch][n]         Sigma_x (without dw)
    const double* sigs;       // [batch][n_el * NINEQ]  Sigma_s of the inequality rows, element-major
    const double* dw;         // [batch]
    const double* dc;         // [batch]
    const int* active;        // [batch] or null: laps with active == 0 are skipped
    double* Dinv;             // [batch][S][NB*NB]  inverse pivot blocks (column-major, ld NB)
    double* F;                // [batch][S][NB*NB]  closed laps: fill blocks K[0, j] (rows: stage 0, columns: stage j)
    double* Aq;               // [batch][S][NINEQ*(NV+1)]  inequality rows of the stage's element w.r.t. its right node, and 1/D
    double* ywork;            // [batch][S][NB]      forward-substitution intermediates
    int* n_neg;               // [batch] negative pivots
    int* n_zero;              // [batch] vanishing pivots
    int ls_mode;              // 1: least-squares multiplier system (W = 0, Sigma_x = 1, D_ineq = 1)
};

constexpr int KLD = 32;                                   // leading dimension of the stage matrices in shared memory
#define KM(M, r, c) (M)[(r) + KLD * (c)]

__device__ __forceinline__ double kkt_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// arg max over the warp, ties to the lowest lane (deterministic)
__device__ __forceinline__ void kkt_warp_argmax(double& v, int& idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// In-place inverse of the symmetric NB x NB matrix M (shared memory, both triangles stored) by Gauss-Jordan sweeps with
// Bunch-Parlett pivoting.  rowbuf: 2 * KLD doubles of shared memory.  Returns the number of negative pivots through
// n_neg and of vanishing ones through n_zero (accumulated).
template <int NB>
__device__ __forceinline__ void kkt_invert(double* M, double* rowbuf, int& n_neg, int& n_zero)
{
    const int lane = threadIdx.x & 31;
    const double ALPHA = 0.6403882032022076;              // (1 + sqrt(17)) / 8
    unsigned swept = 0u;
    const unsigned all = (NB == 32) ? 0xffffffffu : ((1u << NB) - 1u);
    while (swept != all)
    {
        const bool mine = lane < NB && !((swept >> lane) & 1u);
        double dmax = mine ? fabs(KM(M, lane, lane)) : -1.0;
        int k = lane;
        kkt_warp_argmax(dmax, k);
        const double lam = kkt_warp_max((mine && lane != k) ? fabs(KM(M, lane, k)) : 0.0);
        bool two = false;
        int p = k, q = k;
        if (!(dmax >= ALPHA * lam && dmax > 0.0))
        {
            // largest off-diagonal entry of the unswept part
            double best = -1.0;
            int bc = 0;
            if (mine)
                for (int c = 0; c < NB; ++c)
                    if (c != lane && !((swept >> c) & 1u))
                    {
                        const double a = fabs(KM(M, lane, c));
                        if (a > best) { best = a; bc = c; }
                    }
            int br = lane;
            double mu0 = best;
            kkt_warp_argmax(mu0, br);
            bc = __shfl_sync(0xffffffffu, bc, br);
            if (mu0 > 0.0 && !(dmax >= ALPHA * mu0)) { two = true; p = br; q = bc; }
        }
        if (!two)
        {
            double d = KM(M, k, k);
            if (d == 0.0) { d = 1.0e-300; if (lane == 0) ++n_zero; }
            if (lane == 0 && d < 0.0) ++n_neg;
            const double pv = 1.0 / d;
            if (lane < NB) rowbuf[lane] = KM(M, k, lane) * pv;
            __syncwarp();
            if (lane < NB && lane != k)
            {
                const double f = KM(M, lane, k);
#pragma unroll 4
                for (int c = 0; c < NB; ++c) KM(M, lane, c) -= f * rowbuf[c];
                KM(M, lane, k) = -f * pv;
            }
            else if (lane == k)
            {
#pragma unroll 4
                for (int c = 0; c < NB; ++c) KM(M, k, c) = rowbuf[c];
                KM(M, k, k) = pv;
            }
            swept |= 1u << k;
        }
        else
        {
            const double a = KM(M, p, p), b = KM(M, p, q), b2 = KM(M, q, p), c2 = KM(M, q, q);
            const double det = a * c2 - b * b2;
            if (lane == 0) n_neg += (det < 0.0) ? 1 : ((a < 0.0) ? 2 : 0);
            const double i00 = c2 / det, i01 = -b / det, i10 = -b2 / det, i11 = a / det;
            if (lane < NB)
            {
                const double mp = KM(M, p, lane), mq = KM(M, q, lane);
                rowbuf[lane] = i00 * mp + i01 * mq;
                rowbuf[KLD + lane] = i10 * mp + i11 * mq;
            }
            __syncwarp();
            if (lane < NB && lane != p && lane != q)
            {
                const double fp = KM(M, lane, p), fq = KM(M, lane, q);
#pragma unroll 4
                for (int c = 0; c < NB; ++c) KM(M, lane, c) -= fp * rowbuf[c] + fq * rowbuf[KLD + c];
                KM(M, lane, p) = -(fp * i00 + fq * i10);
                KM(M, lane, q) = -(fp * i01 + fq * i11);
            }
            else if (lane == p || lane == q)
            {
                const double* rb = rowbuf + (lane == p ? 0 : KLD);
#pragma unroll 4
                for (int c = 0; c < NB; ++c) KM(M, lane, c) = rb[c];
                KM(M, lane, p) = (lane == p) ? i00 : i10;
                KM(M, lane, q) = (lane == p) ? i01 : i11;
            }
            swept |= (1u << p) | (1u << q);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NV, int NEQ>
struct kkt_stage
{
    static constexpr int NB = NV + NEQ;
    static_assert(NB <= 32, "a stage must fit one warp");

    // element that ends at stage j's node, and whether its left node is a variable node (coupling to stage j-1 exists)
    static __device__ __forceinline__ int e_in(const fl_kkt_dev& K, int j) { const int node = j + K.node0; return node == 0 ? K.n_points - 1 : node - 1; }

    // Diagonal block of stage j into M (zeroed here), inequality rows into Aq (global, per stage) -- without Schur terms.
    static __device__ void gather_diag(const fl_kkt_dev& K, int lap, int j, double* M, double* aq_s /* shared [NINEQ][KLD] + [KLD] */)
    {
        const int lane = threadIdx.x & 31;
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
        for (int c = 0; c < NB; ++c) KM(M, lane, c) = 0.0;
        for (int t = 0; t < K.NINEQ; ++t) aq_s[t * KLD + lane] = 0.0;
        __syncwarp();
        const double dw = K.dw[lap], dc = K.dc[lap];
        if (!K.ls_mode)
        {
            for (int k = lane; k < K.NH; k += 32)
            {
                const int r = K.t.h_row[k], c = K.t.h_col[k];
                const double v = hes[(size_t)k * K.S + j];
                KM(M, r, c) = v;
                KM(M, c, r) = v;
            }
            __syncwarp();
            if (lane < NV) KM(M, lane, lane) += K.sigx[(size_t)lap * K.n + (size_t)j * NV + lane] + dw;
        }
        else if (lane < NV) KM(M, lane, lane) = 1.0;
        if (lane >= NV && lane < NB) KM(M, lane, lane) = -dc;
        __syncwarp();
        for (int k = lane; k < K.NJA; k += 32)
        {
            const int r = K.t.ja_row[k], c = K.t.ja_col[k];
            const double v = jac[(size_t)k * K.n_el + e];
            const int q = K.t.eq_of_row[r];
            if (q >= 0) { KM(M, NV + q, c) = v; KM(M, c, NV + q) = v; }
            else aq_s[K.t.ineq_of_row[r] * KLD + c] = v;
        }
        __syncwarp();
        // 1 / D of the inequality rows
        if (lane < K.NINEQ)
        {
            const double ss = K.ls_mode ? 1.0 : K.sigs[(size_t)lap * K.n_el * K.NINEQ + (size_t)e * K.NINEQ + lane] + dw;
            aq_s[K.NINEQ * KLD + lane] = 1.0 / (1.0 / ss + (K.ls_mode ? 0.0 : dc));
        }
        __syncwarp();
        if (lane < NV)
            for (int t = 0; t < K.NINEQ; ++t)
            {
                const double w = aq_s[K.NINEQ * KLD + t] * aq_s[t * KLD + lane];
                if (w != 0.0)
                    for (int c = 0; c < NV; ++c) KM(M, lane, c) += w * aq_s[t * KLD + c];
            }
        __syncwarp();
    }

    // Coupling block K[stage j, stage j-1] (or, for j == 0 of a closed lap, K[0, S-1]) into L (NB x NV, zeroed here).
    static __device__ void gather_coupling(const fl_kkt_dev& K, int lap, int j, double* L)
    {
        const int lane = threadIdx.x & 31;
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
        for (int c = 0; c < NV; ++c) KM(L, lane, c) = 0.0;
        __syncwarp();
        const int eb = e - K.node0;              // index among the elements with a free left node
        if (eb < 0) return;
        for (int k = lane; k < K.NJB; k += 32)
        {
            const int q = K.t.eq_of_row[K.t.jb_row[k]];
            KM(L, NV + q, K.t.jb_col[k]) = jac[(size_t)K.NJA * K.n_el + (size_t)k * K.nB + eb];
        }
        if (!K.ls_mode && lane < K.NCPL)
        {
            const int cp = K.t.ctrl_pos[lane];
            KM(L, cp, cp) = hes[(size_t)K.NH * K.S + (size_t)lane * K.nC + eb];
        }
        __syncwarp();
    }
};

// shared memory of one warp: M, F, XF (NB x NB each, ld 32 -> 32 x NB), L, XL, P (32 x NV), aq, rowbuf, vectors
template <int NV, int NEQ>
struct kkt_smem
{
    static constexpr int NB = NV + NEQ;
    double M[KLD * NB], F[KLD * NB], XF[KLD * NB], L[KLD * NV], XL[KLD * NV], P[KLD * NV];
    double aq[8 * KLD], rowbuf[2 * KLD];
};

template <int NV, int NEQ>
__global__ void __launch_bounds__(32) k_kkt_factor(const fl_kkt_dev K)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB;
    extern __shared__ __align__(16) unsigned char kkt_raw[];
    kkt_smem<NV, NEQ>& sm = *reinterpret_cast<kkt_smem<NV, NEQ>*>(kkt_raw);
    const int lap = blockIdx.x, lane = threadIdx.x;
    if (K.active && !K.active[lap]) return;
    const int S = K.S;
    const bool cyc = K.closed && S > 2;
    double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    int n_neg = 0, n_zero = 0;
    double d0acc[NB];                          // row `lane` of the Schur terms handed to stage 0 through the fill blocks
#pragma unroll
    for (int c = 0; c < NB; ++c) d0acc[c] = 0.0;
    for (int c = 0; c < NV; ++c) KM(sm.P, lane, c) = 0.0;
    for (int c = 0; c < NB; ++c) KM(sm.F, lane, c) = 0.0;
    __syncwarp();
    if (cyc)
    {
        // F_{S-1} = K[0, S-1]: coupling of stage 0 with its left neighbour (the closing element)
        St::gather_coupling(K, lap, 0, sm.L);
        for (int c = 0; c < NV; ++c) KM(sm.F, lane, c) = KM(sm.L, lane, c);
        __syncwarp();
    }
    for (int j = S - 1; j >= 0; --j)
    {
        St::gather_diag(K, lap, j, sm.M, sm.aq);
        // keep the inequality rows and 1/D for the solves
        for (int t = 0; t < K.NINEQ; ++t)
            if (lane < NV) gA[(size_t)j * (K.NINEQ * (NV + 1)) + t * NV + lane] = sm.aq[t * KLD + lane];
        if (lane < K.NINEQ) gA[(size_t)j * (K.NINEQ * (NV + 1)) + K.NINEQ * NV + lane] = sm.aq[K.NINEQ * KLD + lane];
        // Schur complement of the later stages
        if (lane < NV)
            for (int c = 0; c < NV; ++c) KM(sm.M, lane, c) -= KM(sm.P, lane, c);
        if (j == 0 && cyc && lane < NB)
#pragma unroll
            for (int c = 0; c < NB; ++c) KM(sm.M, lane, c) -= d0acc[c];
        __syncwarp();
        const bool chain = j >= 1 && (j >= 2 || !cyc);          // stage j hands a Schur term to stage j-1 through L_j
        if (j >= 1) St::gather_coupling(K, lap, j, sm.L);
        if (cyc && j == 1)
        {
            // K[0, 1] = L_1^T joins the fill block
            if (lane < NB)
                for (int c = 0; c < NV; ++c) KM(sm.F, c, lane) += KM(sm.L, lane, c);
            __syncwarp();
        }
        kkt_invert<NB>(sm.M, sm.rowbuf, n_neg, n_zero);
        if (lane < NB)
            for (int c = 0; c < NB; ++c) gD[(size_t)j * NB * NB + lane + NB * c] = KM(sm.M, lane, c);
        if (cyc && j >= 1 && lane < NB)
            for (int c = 0; c < NB; ++c) gF[(size_t)j * NB * NB + lane + NB * c] = KM(sm.F, lane, c);
        if (j == 0) break;
        // XL = Dinv L (NB x NV), XF = Dinv F^T (NB x NB)
        if (chain && lane < NB)
        {
            double acc[NV];
#pragma unroll
            for (int c = 0; c < NV; ++c) acc[c] = 0.0;
            for (int k = 0; k < NB; ++k)
            {
                const double a = KM(sm.M, lane, k);
#pragma unroll
                for (int c = 0; c < NV; ++c) acc[c] += a * KM(sm.L, k, c);
            }
#pragma unroll
            for (int c = 0; c < NV; ++c) KM(sm.XL, lane, c) = acc[c];
        }
        if (cyc && lane < NB)
        {
            double acc[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) acc[c] = 0.0;
            for (int k = 0; k < NB; ++k)
            {
                const double a = KM(sm.M, lane, k);
#pragma unroll
                for (int c = 0; c < NB; ++c) acc[c] += a * KM(sm.F, c, k);
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) KM(sm.XF, lane, c) = acc[c];
        }
        __syncwarp();
        // P = L^T XL (what stage j-1 subtracts from its x-x block)
        if (lane < NV)
        {
            double acc[NV];
#pragma unroll
            for (int c = 0; c < NV; ++c) acc[c] = 0.0;
            if (chain)
                for (int k = 0; k < NB; ++k)
                {
                    const double a = KM(sm.L, k, lane);
#pragma unroll
                    for (int c = 0; c < NV; ++c) acc[c] += a * KM(sm.XL, k, c);
                }
#pragma unroll
            for (int c = 0; c < NV; ++c) KM(sm.P, lane, c) = acc[c];
        }
        double fn[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) fn[c] = 0.0;
        if (cyc && lane < NB)
        {
            // stage 0: d0acc += F XF ; new fill F_{j-1} = -F XL (columns: x of stage j-1)
#pragma unroll
            for (int c = 0; c < NB; ++c)
            {
                double a = 0.0;
                for (int k = 0; k < NB; ++k) a += KM(sm.F, lane, k) * KM(sm.XF, k, c);
                d0acc[c] += a;
            }
            if (chain)
                for (int k = 0; k < NB; ++k)
                {
                    const double a = KM(sm.F, lane, k);
#pragma unroll
                    for (int c = 0; c < NV; ++c) fn[c] -= a * KM(sm.XL, k, c);
                }
        }
        __syncwarp();
        if (cyc)
        {
#pragma unroll
            for (int c = 0; c < NV; ++c) KM(sm.F, lane, c) = fn[c];
            for (int c = NV; c < NB; ++c) KM(sm.F, lane, c) = 0.0;
        }
        __syncwarp();
    }
    if (lane == 0) { K.n_neg[lap] = n_neg; K.n_zero[lap] = n_zero; }
}

// Solves K z = (r1, r2) for every active lap with the factorisation above; dx [n], dy [m] in the NLP's own orders.
template <int NV, int NEQ>
__global__ void __launch_bounds__(32) k_kkt_solve(const fl_kkt_dev K, const double* __restrict__ r1, const double* __restrict__ r2,
                                                   double* __restrict__ dx, double* __restrict__ dy)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB;
    extern __shared__ __align__(16) unsigned char kkt_raw[];
    kkt_smem<NV, NEQ>& sm = *reinterpret_cast<kkt_smem<NV, NEQ>*>(kkt_raw);
    const int lap = blockIdx.x, lane = threadIdx.x;
    if (K.active && !K.active[lap]) return;
    const int S = K.S;
    const bool cyc = K.closed && S > 2;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    r1 += (size_t)lap * K.n; r2 += (size_t)lap * K.m; dx += (size_t)lap * K.n; dy += (size_t)lap * K.m;
    double* vb = sm.rowbuf;            // b_j / w (NB)
    double* vy = sm.rowbuf + KLD;      // y_j / z_{j-1}
    double* z0 = sm.aq;                // z_0 (NB)
    const int AQ = K.NINEQ * (NV + 1);

    // condensed right-hand side of stage j (before the updates from later stages)
    auto rhs = [&](int j) -> double {
        const int e = St::e_in(K, j);
        double b = 0.0;
        if (lane < NV)
        {
            b = r1[(size_t)j * NV + lane];
            for (int t = 0; t < K.NINEQ; ++t)
                b += gA[(size_t)j * AQ + t * NV + lane] * gA[(size_t)j * AQ + K.NINEQ * NV + t] * r2[(size_t)e * K.NCPE + K.t.ineq_rows[t]];
        }
        else if (lane < NB) b = r2[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]];
        return b;
    };

    double carry = 0.0;                // update of stage j-1's x rows from stage j
    double b0 = 0.0;                   // updates of stage 0 through the fill blocks
    for (int j = S - 1; j >= 1; --j)
    {
        const bool chain = j >= 2 || !cyc;
        const double b = rhs(j) + carry;
        if (lane < NB) vb[lane] = b;
        __syncwarp();
        double y = 0.0;
        if (lane < NB)
            for (int k = 0; k < NB; ++k) y += gD[(size_t)j * NB * NB + lane + NB * k] * vb[k];
        if (lane < NB) { vy[lane] = y; yw[(size_t)j * NB + lane] = y; }
        __syncwarp();
        carry = 0.0;
        if (chain)
        {
            St::gather_coupling(K, lap, j, sm.L);
            if (lane < NV)
                for (int k = 0; k < NB; ++k) carry -= KM(sm.L, k, lane) * vy[k];
        }
        if (cyc && lane < NB)
            for (int k = 0; k < NB; ++k) b0 -= gF[(size_t)j * NB * NB + lane + NB * k] * vy[k];
        __syncwarp();
    }
    {
        const double b = rhs(0) + carry + b0;
        if (lane < NB) vb[lane] = b;
        __syncwarp();
        double z = 0.0;
        if (lane < NB)
            for (int k = 0; k < NB; ++k) z += gD[lane + NB * k] * vb[k];
        if (lane < NB) { z0[lane] = z; vy[lane] = z; yw[lane] = z; }
        __syncwarp();
    }
    for (int j = 1; j < S; ++j)
    {
        const bool chain = j >= 2 || !cyc;
        // w = L_j z_{j-1} + F_j^T z_0
        double w = 0.0;
        if (chain)
        {
            St::gather_coupling(K, lap, j, sm.L);
            if (lane < NB)
                for (int c = 0; c < NV; ++c) w += KM(sm.L, lane, c) * vy[c];
        }
        if (cyc && lane < NB)
            for (int k = 0; k < NB; ++k) w += gF[(size_t)j * NB * NB + k + NB * lane] * z0[k];
        __syncwarp();
        if (lane < NB) vb[lane] = w;
        __syncwarp();
        double z = 0.0;
        if (lane < NB)
        {
            z = yw[(size_t)j * NB + lane];
            for (int k = 0; k < NB; ++k) z -= gD[(size_t)j * NB * NB + lane + NB * k] * vb[k];
        }
        __syncwarp();
        if (lane < NB) { vy[lane] = z; yw[(size_t)j * NB + lane] = z; }
        __syncwarp();
    }
    // scatter: dx, equality multipliers, then the condensed inequality multipliers dy_t = (a_t . dx_j - r2_t) / D_t
    for (int j = 0; j < S; ++j)
    {
        const int e = St::e_in(K, j);
        const double z = (lane < NB) ? yw[(size_t)j * NB + lane] : 0.0;
        if (lane < NV) dx[(size_t)j * NV + lane] = z;
        else if (lane < NB) dy[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]] = z;
        for (int t = 0; t < K.NINEQ; ++t)
        {
            double a = (lane < NV) ? gA[(size_t)j * AQ + t * NV + lane] * z : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0)
            {
                const int row = e * K.NCPE + K.t.ineq_rows[t];
                dy[row] = (a - r2[row]) * gA[(size_t)j * AQ + K.NINEQ * NV + t];
            }
        }
    }
}
