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
//   * every pivot block is inverted in place by a Gauss-Jordan sweep with Bunch-Kaufman pivoting (1x1 pivots on the
//     largest diagonal, 2x2 pivots when the diagonal is weak), which also counts the negative pivots: by Haynsworth
//     additivity their sum is the inertia Ipopt asks of its linear solver (n positive, m negative, none zero).
//
// One block of four warps per lap: thread (lane r, warp w) owns row r and the columns c = w, w+4, w+8, ... of every stage
// matrix (shared memory, column-major with leading dimension 32, so a warp's access to a column is one conflict-free
// 256-byte row of banks and operands shared by the lanes are broadcast reads); products are accumulated in registers,
// pivot searches are warp shuffles.  The inertia-correcting regularisation (Algorithm IC of Waechter & Biegler) runs
// inside the kernel: a stage whose pivot block has the wrong inertia stops the elimination at once, dw is raised and the
// lap starts over, while the other laps of the batch go on.  Laps are independent: a batch fills the machine with
// several laps per SM.  oracle/kkt.py is the numpy statement of the same algorithm.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "ipm_types.cuh"

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
    const int* lap_list;      // null, or a compact launch: block b serves lap lap_list[b] (the laps still running, fl_ipm.cu)
    double* Dinv;             // [batch][S][NB*NB]  inverse pivot blocks (column-major, ld NB)
    double* F;                // [batch][S][NB*NB]  closed laps: fill blocks K[0, j] (rows: stage 0, columns: stage j)
    double* Aq;               // [batch][S][NINEQ*(NV+1)]  inequality rows of the stage's element w.r.t. its right node, and 1/D
    double* Lst;              // [batch][S][NB*NV]   coupling blocks K[j, j-1] as gathered (dense, column-major, ld NB)
    double* ywork;            // [batch][S][NB]      forward-substitution intermediates
    int* n_neg;               // [batch] negative pivots
    int* n_zero;              // [batch] vanishing pivots
    int ls_mode;              // 1: least-squares multiplier system (W = 0, Sigma_x = 1, D_ineq = 1)
    // partitioned (parallel-in-the-stages) mode, see kkt_part_kernels.cuh: the reduced system over the separator stages
    const int* sep;           // [P] separator stages (ascending, sep[0] = 0), or null: the kernels below then run over all stages
    int P;
    const double* redP;       // [batch][P][NV*NV]  Schur term onto separator k's x-x block from its right segment
    const double* redT;       // [batch][P][NB*NB]  Schur term onto separator k from its left segment (through the fill blocks)
    const double* redL;       // [batch][P][NB*NV]  coupling K~[separator k, separator k-1] (k = 0: the corner block of closed laps)
    const double* segcarry;   // [batch][P][NV]     solves: right-hand-side update of separator k from its right segment
    const double* segrb;      // [batch][P][NB]     solves: the same from its left segment
    fl_ipm_lap* laps;         // [batch] or null.  Not null: dw is read from / written to the lap records and the inertia
                              // correction loop runs in the kernel; null: one pass with dw[], the pivot counts are reported
    fl_ipm_options o;
};

constexpr int KLD = 32;                                   // leading dimension of the stage matrices in shared memory
#ifndef FL_KT
#define FL_KT 128
#endif
constexpr int KT = FL_KT;                                   // threads per lap
constexpr int KW = KT / 32;                               // warps per lap: columns are dealt round-robin to the warps
#define KM(M, r, c) (M)[(r) + KLD * (c)]

__device__ __forceinline__ double kkt_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// arg max over the warp, ties to the lowest index (deterministic)
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

// magnitude of v as an order-preserving 27-bit key with the lane in the low bits: one REDUX gives arg max |v| over a warp
// (ties and magnitudes closer than 2^-15 relative go to the lowest lane; pivoting decisions do not need more)
__device__ __forceinline__ unsigned kkt_key(double v, int lane) { return ((unsigned)__double2hiint(fabs(v)) & ~31u) | (unsigned)(31 - lane); }
__device__ __forceinline__ double kkt_key_value(unsigned key) { return __hiloint2double((int)(key & ~31u), 0); }

// Inverse of the symmetric NB x NB matrix in A (shared memory, both triangles stored) by Gauss-Jordan sweeps with
// Bunch-Kaufman pivoting (candidate: the largest diagonal entry).  Every sweep rewrites the whole matrix, so the sweeps ping-pong between A and B (one barrier
// per sweep); every warp takes the same pivoting decision from the same data.  Thread (lane r, warp w) owns row r and
// the columns w, w + KW, ...  Returns the buffer holding the inverse; thread 0 accumulates the pivot counts.
template <int NB>
__device__ __forceinline__ double* kkt_invert(double* A, double* B, int& n_neg, int& n_zero)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int CB = (NB + KW - 1) / KW;
    const double ALPHA = 0.6403882032022076;              // (1 + sqrt(17)) / 8
    unsigned swept = 0u;
    const unsigned all = (NB == 32) ? 0xffffffffu : ((1u << NB) - 1u);
    double* M = A;
    double* N = B;
    while (swept != all)
    {
        const bool mine = lane < NB && !((swept >> lane) & 1u);
        const unsigned kd = __reduce_max_sync(0xffffffffu, mine ? kkt_key(KM(M, lane, lane), lane) : 0u);
        int p = 31 - (int)(kd & 31u), q = p;
        const double dmax = kkt_key_value(kd);
        const unsigned kl = __reduce_max_sync(0xffffffffu, (mine && lane != p) ? kkt_key(KM(M, lane, p), lane) : 0u);
        const double lam = kkt_key_value(kl);
        bool two = false;
        if (!(dmax >= ALPHA * lam && dmax > 0.0) && lam > 0.0)
        {
            // Bunch-Kaufman: r = row of the largest entry of column p, sigma = largest off-diagonal entry of column r
            const int r = 31 - (int)(kl & 31u);
            const unsigned ks = __reduce_max_sync(0xffffffffu, (mine && lane != r) ? kkt_key(KM(M, lane, r), lane) : 0u);
            const double sigma = kkt_key_value(ks);
            const double arr = fabs(KM(M, r, r));
            if (dmax * sigma >= ALPHA * lam * lam) { }                       // 1x1 pivot on p
            else if (arr >= ALPHA * sigma) { p = r; q = r; }                  // 1x1 pivot on r
            else { two = true; q = r; }                                       // 2x2 pivot on (p, r)
        }
        if (!two)
        {
            double d = KM(M, p, p);
            if (d == 0.0) { d = 1.0e-300; if (threadIdx.x == 0) ++n_zero; }
            if (threadIdx.x == 0 && d < 0.0) ++n_neg;
            const double pv = 1.0 / d;
            if (lane < NB)
            {
                const double f = KM(M, lane, p);
                if (lane != p)
                {
                    const double fpv = f * pv;
#pragma unroll
                    for (int i = 0; i < CB; ++i)
                    {
                        const int c = w + KW * i;
                        if (NB % KW == 0 || c < NB) KM(N, lane, c) = KM(M, lane, c) - fpv * KM(M, p, c);
                    }
                    if ((p & (KW - 1)) == w) KM(N, lane, p) = -fpv;
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < CB; ++i)
                    {
                        const int c = w + KW * i;
                        if (NB % KW == 0 || c < NB) KM(N, p, c) = KM(M, p, c) * pv;
                    }
                    if ((p & (KW - 1)) == w) KM(N, p, p) = pv;
                }
            }
            swept |= 1u << p;
        }
        else
        {
            const double a = KM(M, p, p), b = KM(M, p, q), b2 = KM(M, q, p), c2 = KM(M, q, q);
            const double det = a * c2 - b * b2;
            if (threadIdx.x == 0) n_neg += (det < 0.0) ? 1 : ((a < 0.0) ? 2 : 0);
            const double i00 = c2 / det, i01 = -b / det, i10 = -b2 / det, i11 = a / det;
            if (lane < NB)
            {
                const double fp = KM(M, lane, p), fq = KM(M, lane, q);
                const bool piv_row = lane == p || lane == q;
#pragma unroll
                for (int i = 0; i < CB; ++i)
                {
                    const int c = w + KW * i;
                    if (NB % KW == 0 || c < NB)
                    {
                        const double mp = KM(M, p, c), mq = KM(M, q, c);
                        const double rp = i00 * mp + i01 * mq, rq = i10 * mp + i11 * mq;
                        KM(N, lane, c) = piv_row ? (lane == p ? rp : rq) : KM(M, lane, c) - fp * rp - fq * rq;
                    }
                }
                if ((p & (KW - 1)) == w) KM(N, lane, p) = piv_row ? (lane == p ? i00 : i10) : -(fp * i00 + fq * i10);
                if ((q & (KW - 1)) == w) KM(N, lane, q) = piv_row ? (lane == p ? i01 : i11) : -(fp * i01 + fq * i11);
            }
            swept |= (1u << p) | (1u << q);
        }
        __syncthreads();
        double* tmp = M; M = N; N = tmp;
    }
    return M;
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NV, int NEQ>
struct kkt_stage
{
    static constexpr int NB = NV + NEQ;
    static constexpr int CB = (NB + KW - 1) / KW, CV = (NV + KW - 1) / KW;      // columns per thread
    static_assert(NB <= 32, "a stage must fit one warp of rows");

    // element that ends at stage j's node
    static __device__ __forceinline__ int e_in(const fl_kkt_dev& K, int j) { const int node = j + K.node0; return node == 0 ? K.n_points - 1 : node - 1; }

    // Diagonal block of stage j into M, inequality rows and 1/D into aq_s ([NINEQ][KLD] + [KLD]) -- without Schur terms.
    static __device__ void gather_diag(const fl_kkt_dev& K, int lap, int j, double dw, double dc, double* M, double* aq_s)
    {
        const int t = threadIdx.x, lane = t & 31, w = t >> 5;
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(M, lane, c) = 0.0; }
        for (int k = t; k < (K.NINEQ + 1) * KLD; k += KT) aq_s[k] = 0.0;
        __syncthreads();
        if (!K.ls_mode)
            for (int k = t; k < K.NH; k += KT)
            {
                const int r = K.t.h_row[k], c = K.t.h_col[k];
                const double v = hes[(size_t)k * K.S + j];
                KM(M, r, c) = v;
                KM(M, c, r) = v;
            }
        for (int k = t; k < K.NJA; k += KT)
        {
            const int r = K.t.ja_row[k], c = K.t.ja_col[k];
            const double v = jac[(size_t)k * K.n_el + e];
            const int q = K.t.eq_of_row[r];
            if (q >= 0) { KM(M, NV + q, c) = v; KM(M, c, NV + q) = v; }
            else aq_s[K.t.ineq_of_row[r] * KLD + c] = v;
        }
        if (t < K.NINEQ)
        {
            const double ss = K.ls_mode ? 1.0 : K.sigs[(size_t)lap * K.n_el * K.NINEQ + (size_t)e * K.NINEQ + t] + dw;
            aq_s[K.NINEQ * KLD + t] = 1.0 / (1.0 / ss + (K.ls_mode ? 0.0 : dc));
        }
        __syncthreads();
        if (t < NV) KM(M, t, t) = K.ls_mode ? 1.0 : KM(M, t, t) + K.sigx[(size_t)lap * K.n + (size_t)j * NV + t] + dw;
        else if (t < NB) KM(M, t, t) = -dc;
        __syncthreads();
        if (lane < NV)
        {
            double acc[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) acc[i] = 0.0;
            for (int q = 0; q < K.NINEQ; ++q)
            {
                const double wq = aq_s[K.NINEQ * KLD + q] * aq_s[q * KLD + lane];
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += wq * aq_s[q * KLD + c]; }
            }
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(M, lane, c) += acc[i]; }
        }
        __syncthreads();
    }

    // Coupling block K[stage j, stage j-1] (or, for j == 0 of a closed lap, K[0, S-1]) into L (NB x NV).
    static __device__ void gather_coupling(const fl_kkt_dev& K, int lap, int j, double* L)
    {
        const int t = threadIdx.x, lane = t & 31, w = t >> 5;
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
#pragma unroll
        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(L, lane, c) = 0.0; }
        __syncthreads();
        const int eb = e - K.node0;              // index among the elements with a free left node
        if (eb >= 0)
        {
            for (int k = t; k < K.NJB; k += KT)
            {
                const int q = K.t.eq_of_row[K.t.jb_row[k]];
                KM(L, NV + q, K.t.jb_col[k]) = jac[(size_t)K.NJA * K.n_el + (size_t)k * K.nB + eb];
            }
            if (!K.ls_mode && t < K.NCPL)
            {
                const int cp = K.t.ctrl_pos[t];
                KM(L, cp, cp) = hes[(size_t)K.NH * K.S + (size_t)t * K.nC + eb];
            }
        }
        __syncthreads();
    }
};

// shared memory of one lap: M, M2 (sweep buffers), F (32 x NB each), L, XL, P (32 x NV), inequality rows
template <int NV, int NEQ>
struct kkt_smem
{
    static constexpr int NB = NV + NEQ;
    double M[KLD * NB], M2[KLD * NB], F[KLD * NB], L[KLD * NV], XL[KLD * NV], P[KLD * NV];     // D^-1 F^T reuses the idle half of (M, M2)
    double aq[8 * KLD];
    double dw_next;
    int flag;
};

#ifndef FL_KMINB
#define FL_KMINB 6          // 80 registers instead of 96: six laps per SM instead of five (factor phase of the 4096-lap sweep 10.25 -> 9.98 s)
#endif
template <int NV, int NEQ>
__global__ void __launch_bounds__(KT, FL_KMINB) k_kkt_factor(const fl_kkt_dev K)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB, CB = St::CB, CV = St::CV;
    extern __shared__ __align__(16) unsigned char kkt_raw[];
    kkt_smem<NV, NEQ>& sm = *reinterpret_cast<kkt_smem<NV, NEQ>*>(kkt_raw);
    const int lap = K.lap_list ? K.lap_list[blockIdx.x] : blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (K.active && !K.active[lap]) return;
    fl_ipm_lap* Lp = K.laps ? K.laps + lap : nullptr;
    if (Lp && Lp->status != IPM_RUNNING) return;
    const int S = K.sep ? K.P : K.S;           // reduced system over the separators, or the whole chain
    const bool cyc = K.closed && S > 2;
    double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    double* gL = K.Lst + (size_t)lap * S * NB * NV;
    double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    const int AQ = K.NINEQ * (NV + 1);
    const double* rP = K.sep ? K.redP + (size_t)lap * S * NV * NV : nullptr;
    const double* rT = K.sep ? K.redT + (size_t)lap * S * NB * NB : nullptr;
    const double* rL = K.sep ? K.redL + (size_t)lap * S * NB * NV : nullptr;
    // coupling block of (reduced) stage j into sm.L
    auto coupling = [&](int j) {
        if (!K.sep) { St::gather_coupling(K, lap, j, sm.L); return; }
        if (lane < NB)
        {
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.L, lane, c) = rL[(size_t)j * NB * NV + lane + NB * c]; }
        }
        __syncthreads();
    };
    const double dc = Lp ? Lp->dc : K.dc[lap];
    double dw = Lp ? Lp->dw : K.dw[lap];
    // (tried: starting from the decayed regularisation of the previous iteration instead of 0 -- it saves failed
    //  factorisations but nearly doubles the iteration count: Algorithm IC's "try 0 first" is kept)

    for (int attempt = 0;; ++attempt)
    {
        int n_neg = 0, n_zero = 0;
        bool ok = true;
        double d0acc[CB];                          // (row lane, own columns) of the Schur terms handed to stage 0 through the fill blocks
#pragma unroll
        for (int i = 0; i < CB; ++i) d0acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.P, lane, c) = 0.0; }
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(sm.F, lane, c) = 0.0; }
        __syncthreads();
        if (cyc)
        {
            // F_{S-1} = K[0, S-1]: coupling of stage 0 with its left neighbour (the closing element)
            coupling(0);
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.F, lane, c) = KM(sm.L, lane, c); }
            __syncthreads();
        }
        for (int j = S - 1; j >= 0; --j)
        {
            St::gather_diag(K, lap, K.sep ? K.sep[j] : j, dw, dc, sm.M, sm.aq);
            if (K.sep && lane < NB)
            {
                // what the eliminated segments on both sides of this separator hand to it
#pragma unroll
                for (int i = 0; i < CB; ++i)
                {
                    const int c = w + KW * i;
                    if (c < NB)
                    {
                        double v = rT[(size_t)j * NB * NB + lane + NB * c];
                        if (lane < NV && c < NV) v += rP[(size_t)j * NV * NV + lane + NV * c];
                        KM(sm.M, lane, c) -= v;
                    }
                }
            }
            // keep the inequality rows and 1/D for the solves
            for (int k = t; k < K.NINEQ * NV; k += KT) gA[(size_t)j * AQ + k] = sm.aq[(k / NV) * KLD + (k % NV)];
            if (t < K.NINEQ) gA[(size_t)j * AQ + K.NINEQ * NV + t] = sm.aq[K.NINEQ * KLD + t];
            // Schur complement of the later stages
            if (lane < NV)
            {
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.M, lane, c) -= KM(sm.P, lane, c); }
            }
            if (j == 0 && cyc && lane < NB)
            {
#pragma unroll
                for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(sm.M, lane, c) -= d0acc[i]; }
            }
            __syncthreads();
            const bool chain = j >= 1 && (j >= 2 || !cyc);          // stage j hands a Schur term to stage j-1 through L_j
            // non-zero columns of the fill block: x of stage j, except at stage 1 where K[0, 1] = L_1^T brings the multiplier columns
            const int KF = (cyc && j == 1) ? NB : NV;
            const int ncpl = K.ls_mode ? 0 : K.NCPL;
            const int kL0 = K.sep ? 0 : NV;            // first row of L that can be non-zero (reduced couplings are dense)
            if (j >= 1)
            {
                coupling(j);
                if (lane < NB)
                {
#pragma unroll
                    for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) gL[(size_t)j * NB * NV + lane + NB * c] = KM(sm.L, lane, c); }
                }
            }
            if (cyc && j == 1)
            {
                // K[0, 1] = L_1^T joins the fill block: F[c][r] += L[r][c]
                if (lane < NB)
                {
#pragma unroll
                    for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.F, c, lane) += KM(sm.L, lane, c); }
                }
                __syncthreads();
            }
            const int neg_before = n_neg, zero_before = n_zero;
            const double* Mi = kkt_invert<NB>(sm.M, sm.M2, n_neg, n_zero);
            double* XF = (Mi == sm.M) ? sm.M2 : sm.M;           // the sweep buffer that does not hold the inverse
            if (Lp)
            {
                // every pivot block carries exactly NEQ negative pivots when the reduced Hessian is positive definite
                if (t == 0) sm.flag = (n_neg - neg_before == NEQ && n_zero == zero_before) ? 1 : 0;
                __syncthreads();
                ok = sm.flag != 0;
                __syncthreads();
                if (!ok) break;
            }
            if (lane < NB)
            {
#pragma unroll
                for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) gD[(size_t)j * NB * NB + lane + NB * c] = KM(Mi, lane, c); }
                if (cyc && j >= 1)
                {
#pragma unroll
                    for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) gF[(size_t)j * NB * NB + lane + NB * c] = KM(sm.F, lane, c); }
                }
            }
            if (j == 0) break;
            // XL = Dinv L (NB x NV), XF = Dinv F^T (NB x NB)
            if (lane < NB)
            {
                if (chain)
                {
                    double acc[CV];
#pragma unroll
                    for (int i = 0; i < CV; ++i) acc[i] = 0.0;
                    // rows of L that can be non-zero: the NEQ multiplier rows, and the diagonal coupling entries (cp, cp)
#pragma unroll 4
                    for (int k = kL0; k < NB; ++k)
                    {
                        const double a = KM(Mi, lane, k);
#pragma unroll
                        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += a * KM(sm.L, k, c); }
                    }
                    for (int q = 0; q < (K.sep ? 0 : ncpl); ++q)
                    {
                        const int cp = K.t.ctrl_pos[q];
#pragma unroll
                        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c == cp) acc[i] += KM(Mi, lane, cp) * KM(sm.L, cp, cp); }
                    }
#pragma unroll
                    for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.XL, lane, c) = acc[i]; }
                }
                if (cyc)
                {
                    double acc[CB];
#pragma unroll
                    for (int i = 0; i < CB; ++i) acc[i] = 0.0;
#pragma unroll 5
                    for (int k = 0; k < KF; ++k)
                    {
                        const double a = KM(Mi, lane, k);
#pragma unroll
                        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) acc[i] += a * KM(sm.F, c, k); }
                    }
#pragma unroll
                    for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) KM(XF, lane, c) = acc[i]; }
                }
            }
            __syncthreads();
            // P = L^T XL (what stage j-1 subtracts from its x-x block)
            if (lane < NV)
            {
                double acc[CV];
#pragma unroll
                for (int i = 0; i < CV; ++i) acc[i] = 0.0;
                if (chain)
                {
#pragma unroll 4
                    for (int k = kL0; k < NB; ++k)
                    {
                        const double a = KM(sm.L, k, lane);
#pragma unroll
                        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += a * KM(sm.XL, k, c); }
                    }
                    for (int q = 0; q < (K.sep ? 0 : ncpl); ++q)
                    {
                        const int cp = K.t.ctrl_pos[q];
                        if (lane == cp)
                        {
                            const double a = KM(sm.L, cp, cp);
#pragma unroll
                            for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) acc[i] += a * KM(sm.XL, cp, c); }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) KM(sm.P, lane, c) = acc[i]; }
            }
            double fn[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) fn[i] = 0.0;
            if (cyc && lane < NB)
            {
                // stage 0: d0acc += F XF ; new fill F_{j-1} = -F XL (columns: x of stage j-1)
#pragma unroll 5
                for (int k = 0; k < KF; ++k)
                {
                    const double a = KM(sm.F, lane, k);
#pragma unroll
                    for (int i = 0; i < CB; ++i) { const int c = w + KW * i; if (c < NB) d0acc[i] += a * KM(XF, k, c); }
                    if (chain)
                    {
#pragma unroll
                        for (int i = 0; i < CV; ++i) { const int c = w + KW * i; if (c < NV) fn[i] -= a * KM(sm.XL, k, c); }
                    }
                }
            }
            __syncthreads();
            if (cyc)
            {
#pragma unroll
                for (int i = 0; i < CB; ++i)
                {
                    const int c = w + KW * i;
                    if (c < NB) KM(sm.F, lane, c) = (c < NV && i < CV) ? fn[i < CV ? i : 0] : 0.0;
                }
            }
            __syncthreads();
        }
        if (!Lp)
        {
            if (t == 0) { K.n_neg[lap] = n_neg; K.n_zero[lap] = n_zero; }
            return;
        }
        // Algorithm IC (Waechter & Biegler, section 3.1)
        if (t == 0)
        {
            const fl_ipm_options& o = K.o;
            Lp->n_fact += 1;
            if (ok)
            {
                if (dw > 0.0) Lp->dw_last = dw;
                Lp->dw = dw;
                Lp->dw_prev = dw;
                sm.flag = 1;
            }
            else
            {
                Lp->n_reg += 1;
                if (dw == 0.0) dw = (Lp->dw_last == 0.0) ? o.delta_w_0 : fmax(o.delta_w_min, o.kappa_w_minus * Lp->dw_last);
                else dw *= (Lp->dw_last == 0.0) ? o.kappa_w_plus_bar : o.kappa_w_plus;
                if (dw > o.delta_w_max) { Lp->status = IPM_REGULARISATION_FAILED; sm.flag = 2; }
                else sm.flag = 0;
                sm.dw_next = dw;
            }
        }
        __syncthreads();
        const int fl = sm.flag;
        if (fl != 0) return;
        dw = sm.dw_next;
        __syncthreads();
    }
}

// Solves K z = (r1, r2) for every active lap with the factorisation above; dx [n], dy [m] in the NLP's own orders.
// Thread (lane r, warp w) multiplies row r of the stage's matrices with the entries k = w, w+4, ... of the vector; the
// operands of the next stage are loaded into registers while the current stage is reduced.
template <int NV, int NEQ>
__global__ void __launch_bounds__(KT) k_kkt_solve(const fl_kkt_dev K, const double* __restrict__ r1, const double* __restrict__ r2,
                                                   double* __restrict__ dx, double* __restrict__ dy)
{
    using St = kkt_stage<NV, NEQ>;
    constexpr int NB = St::NB, CB = St::CB, CV = St::CV;
    __shared__ double vb[KLD], vy[KLD], z0[KLD], part[3][KW][KLD];
    const int lap = K.lap_list ? K.lap_list[blockIdx.x] : blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (K.active && !K.active[lap]) return;
    if (K.laps && K.laps[lap].status != IPM_RUNNING) return;
    const int S = K.sep ? K.P : K.S;
    const bool cyc = K.closed && S > 2;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    const double* gL = K.Lst + (size_t)lap * S * NB * NV;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    const double* sc = K.sep ? K.segcarry + (size_t)lap * S * NV : nullptr;
    const double* sb = K.sep ? K.segrb + (size_t)lap * S * NB : nullptr;
    r1 += (size_t)lap * K.n; r2 += (size_t)lap * K.m; dx += (size_t)lap * K.n; dy += (size_t)lap * K.m;
    const int AQ = K.NINEQ * (NV + 1);

    // condensed right-hand side of stage j (before the updates from later stages); threads t < NB
    auto rhs = [&](int j) -> double {
        const int js = K.sep ? K.sep[j] : j;          // stage of the chain
        const int e = St::e_in(K, js);
        double b = 0.0;
        if (K.sep && t < NB) b = sb[(size_t)j * NB + t] + (t < NV ? sc[(size_t)j * NV + t] : 0.0);
        if (t < NV)
        {
            b += r1[(size_t)js * NV + t];
            for (int q = 0; q < K.NINEQ; ++q)
                b += gA[(size_t)j * AQ + q * NV + t] * gA[(size_t)j * AQ + K.NINEQ * NV + q] * r2[(size_t)e * K.NCPE + K.t.ineq_rows[q]];
        }
        else if (t < NB) b += r2[(size_t)e * K.NCPE + K.t.eq_rows[t - NV]];
        return b;
    };
    double dreg[CB], freg[CB], lreg[CB];
    auto load_rows = [&](const double* g, int j, double* reg) {            // reg[i] = G_j[lane][w + KW i]
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; reg[i] = (lane < NB && c < NB) ? g[(size_t)j * NB * NB + lane + NB * c] : 0.0; }
    };
    auto load_LT = [&](int j, double* reg) {                               // reg[i] = L_j[w + KW i][lane]  (lane < NV)
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int k = w + KW * i; reg[i] = (lane < NV && k < NB) ? gL[(size_t)j * NB * NV + k + NB * lane] : 0.0; }
    };

    // ---- forward: y_j = Dinv_j b_j,  b_{j-1} -= L_j^T y_j,  b_0 -= F_j y_j
    double carry = 0.0, b0 = 0.0;            // held by threads t < NV / t < NB
    if (S > 1) { load_rows(gD, S - 1, dreg); if (cyc) load_rows(gF, S - 1, freg); load_LT(S - 1, lreg); }
    for (int j = S - 1; j >= 1; --j)
    {
        const bool chain = j >= 2 || !cyc;
        if (t < NB) vb[t] = rhs(j) + carry;
        __syncthreads();
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int k = w + KW * i; if (k < NB) a += dreg[i] * vb[k]; }
        part[0][w][lane] = a;
        double dn[CB], fnx[CB], ln[CB];
        if (j >= 2) { load_rows(gD, j - 1, dn); if (cyc) load_rows(gF, j - 1, fnx); load_LT(j - 1, ln); }
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
            const int k = w + KW * i;
            if (k < NB) { if (chain) c1 += lreg[i] * vy[k]; if (cyc) c2 += freg[i] * vy[k]; }
        }
        part[1][w][lane] = c1; part[2][w][lane] = c2;
        __syncthreads();
        if (t < NB)
        {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) { s1 += part[1][ww][t]; s2 += part[2][ww][t]; }
            carry = (t < NV) ? -s1 : 0.0;
            b0 -= s2;
        }
        if (j >= 2)
        {
#pragma unroll
            for (int i = 0; i < CB; ++i) { dreg[i] = dn[i]; freg[i] = fnx[i]; lreg[i] = ln[i]; }
        }
        __syncthreads();
    }
    {
        load_rows(gD, 0, dreg);
        if (t < NB) vb[t] = rhs(0) + carry + b0;
        __syncthreads();
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int k = w + KW * i; if (k < NB) a += dreg[i] * vb[k]; }
        part[0][w][lane] = a;
        __syncthreads();
        if (t < NB)
        {
            double z = 0.0;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) z += part[0][ww][t];
            z0[t] = z; vy[t] = z; yw[t] = z;
        }
        __syncthreads();
    }
    // ---- backward: z_j = y_j - Dinv_j (L_j z_{j-1} + F_j^T z_0)
    auto load_L = [&](int j, double* reg) {                                // reg[i] = L_j[lane][w + KW i]
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int c = w + KW * i; reg[i] = (lane < NB && c < NV) ? gL[(size_t)j * NB * NV + lane + NB * c] : 0.0; }
    };
    auto load_FT = [&](int j, double* reg) {                               // reg[i] = F_j[w + KW i][lane]
#pragma unroll
        for (int i = 0; i < CB; ++i) { const int k = w + KW * i; reg[i] = (lane < NB && k < NB) ? gF[(size_t)j * NB * NB + k + NB * lane] : 0.0; }
    };
    if (S > 1) { load_rows(gD, 1, dreg); load_L(1, lreg); if (cyc) load_FT(1, freg); }
    for (int j = 1; j < S; ++j)
    {
        const bool chain = j >= 2 || !cyc;
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < CB; ++i)
        {
            const int c = w + KW * i;
            if (chain && c < NV) a += lreg[i] * vy[c];
            if (cyc && c < NB) a += freg[i] * z0[c];
        }
        part[0][w][lane] = a;
        double dn[CB], fnx[CB], ln[CB];
        if (j + 1 < S) { load_rows(gD, j + 1, dn); load_L(j + 1, ln); if (cyc) load_FT(j + 1, fnx); }
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
        for (int i = 0; i < CB; ++i) { const int k = w + KW * i; if (k < NB) b += dreg[i] * vb[k]; }
        part[1][w][lane] = b;
        __syncthreads();
        if (t < NB)
        {
            double z = yj;
#pragma unroll
            for (int ww = 0; ww < KW; ++ww) z -= part[1][ww][t];
            vy[t] = z; yw[(size_t)j * NB + t] = z;
        }
        if (j + 1 < S)
        {
#pragma unroll
            for (int i = 0; i < CB; ++i) { dreg[i] = dn[i]; freg[i] = fnx[i]; lreg[i] = ln[i]; }
        }
        __syncthreads();
    }
    // ---- scatter: dx, equality multipliers, then the condensed inequality multipliers dy_q = (a_q . dx_j - r2_q) / D_q
    for (int j = w; j < S; j += KW)
    {
        const int js = K.sep ? K.sep[j] : j;
        const int e = St::e_in(K, js);
        const double z = (lane < NB) ? yw[(size_t)j * NB + lane] : 0.0;
        if (lane < NV) dx[(size_t)js * NV + lane] = z;
        else if (lane < NB) dy[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]] = z;
        for (int q = 0; q < K.NINEQ; ++q)
        {
            double a = (lane < NV) ? gA[(size_t)j * AQ + q * NV + lane] * z : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0)
            {
                const int row = e * K.NCPE + K.t.ineq_rows[q];
                dy[row] = (a - r2[row]) * gA[(size_t)j * AQ + K.NINEQ * NV + q];
            }
        }
    }
}
