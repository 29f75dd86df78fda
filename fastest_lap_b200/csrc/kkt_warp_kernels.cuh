// Warp-per-lap schedule of the batched block-(cyclic-)tridiagonal KKT factorisation and solves (sm_100a, fp64).
//
// Same elimination as kkt_kernels.cuh (stages eliminated from the last one backwards, the fill F_j = K[0, j] of closed laps
// carried to stage 0, Bunch-Kaufman pivoting inside every pivot block, inertia from the pivot signs; the reference hands this
// system to MUMPS inside Ipopt, call site src/core/applications/optimal_laptime.hpp:546-551) -- restated for ONE WARP PER LAP
// with the stage matrix held in REGISTERS: lane r owns row r (NB <= 32 doubles, statically indexed), so the O(NB^3) sweeps are
// one DFMA per matrix entry with the pivot row arriving as broadcast 128-bit shared-memory loads, instead of the
// load-load-FMA-store round trip per entry of the block-per-lap kernel (profiles/prof_r1y: ~10 instructions per useful
// DFMA, short-scoreboard bound).  Used for batches (many laps in flight); few laps keep the parallel-in-the-stages schedule
// of kkt_part_kernels.cuh.
//
// Pivot block inverse: the Gauss-Jordan sweeps of kkt_invert, step for step (Bunch-Kaufman pivots, multipliers from the
// lane's own entry of the pivot column), on register rows.  A pivot index is the same in every lane, so "the register of
// column p" is reached through a switch that is a uniform branch (kw_get / kw_set); the diagonal entry of a lane, whose
// index differs per lane, is mirrored in a side register for the pivot search.  (A cheaper symmetric variant that reads a
// lane's multiplier from the pivot row instead was tried first and rejected: with the pivots of order delta_c that the
// dependent rows of the massless-wheel model produce, the explicit inverse has entries of order 1e9 and the rounding-level
// asymmetry of the sweeps is amplified to solve errors of 1e-5 .. 1e-2 where this form gives 1e-9.)
//
// Compact factors for the solves (closed laps): the fill block has non-zero rows only where the corner block K[0, S-1]
// has them (the NEQ multiplier rows and the control couplings of stage 0: NF <= 16 rows) and NV non-zero columns, the
// coupling blocks have NEQ non-zero rows: per stage the solves stream D^-1 (NB x NB), NEQ x NV coupling entries and an
// NV x NF fill block instead of three dense NB x NB / NB x NV blocks.
#pragma once
#include "kkt_kernels.cuh"
#include "fl_math.cuh"

#ifndef FL_KW_MINB
#define FL_KW_MINB 11         // laps per SM that the 19.7 KB of shared memory per lap allow (register cap 186)
#endif
// smallest leading dimension >= n with ld = 2 (mod 4): rows start 16-byte aligned, and a quarter warp storing 128-bit
// words to consecutive rows touches all 32 banks once
__host__ __device__ constexpr int kw_ld(int n) { return n % 4 == 2 ? n : (n % 4 == 3 ? n + 3 : (n % 4 == 0 ? n + 2 : n + 1)); }

// factor stores: written once per factorisation and read back by the solves after the whole batch has been factorised -- streaming
// (evict-first) so that they do not push the Jacobian / Hessian values of the coming stages out of the L2
#ifndef FL_KW_STREAM
#define FL_KW_STREAM 1
#endif
#if FL_KW_STREAM
#define KW_STORE(p, v) __stcs((p), (v))
#else
#define KW_STORE(p, v) (*(p) = (v))
#endif

// stages ahead of which the factorisation prefetches the values of its gathers into L2 (prefetch_stage); measured on a 4096-lap,
// 40-iteration solve: factor phase 3.19 s without, 3.33 s two stages ahead, 3.35 s four stages ahead -- off
#ifndef FL_KW_PREFETCH
#define FL_KW_PREFETCH 0
#endif

template <int NV, int NEQ>
struct kw_smem
{
    static constexpr int NB = NV + NEQ;
    static constexpr int LDV = kw_ld(NV);          // row-major NB x NV (XL), (NEQ + 1) x NV (compact coupling rows + the two control couplings)
    static constexpr int LDB = kw_ld(NB);          // row-major NB x NB (stage 1 of closed laps)
    static constexpr int LDF = 18;                 // row-major NB x 16 (D^-1 F^T)
    static constexpr int FT = 16;                  // fill block transposed: Ft[k][i], k < NV columns, i < NF compact rows
    static constexpr int ZSZ = (32 * NB > NB * LDB) ? 32 * NB : NB * LDB;
    static constexpr int OXL = 0, OLC = NB * LDV, OFT = OLC + (NEQ + 1) * LDV, WSZ0 = OFT + NV * FT;
    static constexpr int WSZ = WSZ0 > 32 * NB ? (WSZ0 > NB * LDB ? WSZ0 : NB * LDB) : (32 * NB > NB * LDB ? 32 * NB : NB * LDB);
    double Z[ZSZ];             // column-major staging matrix Z[r + 32 c]; row-major scratch between phases
    double W[WSZ];             // XL | Lc | Ft; stage 1 / stage 0 of closed laps reuse it whole
    static constexpr int PS = 17;                  // stride of the parked Schur term (odd: rows and columns both read without bank conflicts)
    static constexpr int AQSZ = 8 * 32 > NV * PS ? 8 * 32 : (NV * PS + 1) / 2 * 2;      // even: what follows stays 16-byte aligned
    double aq[AQSZ];           // inequality rows of the stage's element and 1 / D; pivot coefficients during the extraction; then the
                               // Schur term P handed to the next stage (P[lane][c] at aq[c * PS + lane])
    double d0[16 * 16];        // closed laps: Schur terms handed to stage 0 through the fill blocks, compact rows: d0[i' * 16 + lane]
    double pr[2 * 32];         // broadcast pivot rows
    // (the compact fill block holds at most 16 rows: NEQ + NCPL <= 16 is checked by the host before this schedule is chosen)
};

// ---------------------------------------------------------------------------------------------------------------------
// Register row addressed by an index that is the SAME in every lane (a pivot column): the switch is a uniform branch, every
// case touches one statically named register.
#define KW_CASES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) \
    X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)
template <int NB>
__device__ __forceinline__ double kw_get(const double (&m)[NB], const int idx)
{
    double r = 0.0;
#define KW_GET(k) case k: if constexpr (k < NB) r = m[k]; break;
    switch (idx) { KW_CASES(KW_GET) default: break; }
#undef KW_GET
    return r;
}
template <int NB>
__device__ __forceinline__ void kw_set(double (&m)[NB], const int idx, const double v)
{
#define KW_SET(k) case k: if constexpr (k < NB) m[k] = v; break;
    switch (idx) { KW_CASES(KW_SET) default: break; }
#undef KW_SET
}

// Inverse of the matrix whose row `lane` is m[0..NB) (diag = m[lane], handed over separately because a lane cannot name that
// register), by Gauss-Jordan sweeps with Bunch-Kaufman pivoting -- the algorithm of kkt_invert (kkt_kernels.cuh) step for
// step: largest diagonal entry as candidate, the two-column test, 1x1 or 2x2 pivots, multipliers taken from the lane's own
// entry of the pivot column.  The only liberty: the row of a 1x1 pivot keeps its scale (M^-1[p][:] = rho_p m_p[:], applied at
// the end), which saves scaling a whole row inside one lane.  On return m[c] = (M^-1)[lane][c]; the pivot counts are
// accumulated in n_neg / n_zero (same value in every lane).  pr: two pivot-row buffers of 32 doubles.
template <int NB>
__device__ __forceinline__ void kw_invert(double (&m)[NB], double diag, double* __restrict__ pr, const int lane, int& n_neg, int& n_zero)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NP = (NB + 1) / 2;                                  // 128-bit words per row
    const double ALPHA = 0.6403882032022076;                          // (1 + sqrt(17)) / 8
    const unsigned all = (NB == 32) ? FULL : ((1u << NB) - 1u);
    unsigned swept = 0u;
    double rho = 1.0;
    double* pr0 = pr;
    double* pr1 = pr + 32;
    auto put_row = [&](double* dst) {
#pragma unroll
        for (int i = 0; i < NB / 2; ++i) reinterpret_cast<double2*>(dst)[i] = make_double2(m[2 * i], m[2 * i + 1]);
        if (NB & 1) dst[NB - 1] = m[NB - 1];
    };
    while (swept != all)
    {
        const bool mine = lane < NB && !((swept >> lane) & 1u);
        const unsigned kd = __reduce_max_sync(FULL, mine ? kkt_key(diag, lane) : 0u);
        const int p = 31 - (int)(kd & 31u);
        const double dmax = kkt_key_value(kd);
        const double fp = kw_get<NB>(m, p);                          // M[lane][p]
        const unsigned kl = __reduce_max_sync(FULL, (mine && lane != p) ? kkt_key(fp, lane) : 0u);
        const double lam = kkt_key_value(kl);
        int piv = p, q = -1;
        bool two = false;
        double fq = 0.0;
        if (!(dmax >= ALPHA * lam && dmax > 0.0) && lam > 0.0)
        {
            // Bunch-Kaufman: r = row of the largest entry of column p, sigma = largest off-diagonal entry of column r
            const int r = 31 - (int)(kl & 31u);
            fq = kw_get<NB>(m, r);
            const unsigned ks = __reduce_max_sync(FULL, (mine && lane != r) ? kkt_key(fq, lane) : 0u);
            const double sigma = kkt_key_value(ks);
            const double arr = fabs(__shfl_sync(FULL, diag, r));
            if (dmax * sigma >= ALPHA * lam * lam) { }                  // 1x1 pivot on p
            else if (arr >= ALPHA * sigma) piv = r;                     // 1x1 pivot on r
            else { two = true; q = r; }                                 // 2x2 pivot on (p, r)
        }
        if (!two)
        {
            if (lane == piv) put_row(pr0);
            __syncwarp();
            double d = __shfl_sync(FULL, diag, piv);
            if (d == 0.0) { d = 1.0e-300; ++n_zero; }
            if (d < 0.0) ++n_neg;
            const double pv = fl_rcp(d);
            const double f = (piv == p) ? fp : fq;
            const double nf = (lane == piv) ? 0.0 : -(f * pv);
#pragma unroll
            for (int i = 0; i < NP; ++i)
            {
                const double2 gv = reinterpret_cast<const double2*>(pr0)[i];
                m[2 * i] = fma(nf, gv.x, m[2 * i]);
                if (2 * i + 1 < NB) m[2 * i + 1] = fma(nf, gv.y, m[2 * i + 1]);
            }
            diag = (lane == piv) ? 1.0 : fma(nf, (lane < NB) ? pr0[lane] : 0.0, diag);
            kw_set<NB>(m, piv, (lane == piv) ? 1.0 : nf);
            if (lane == piv) rho = pv;
            swept |= 1u << piv;
        }
        else
        {
            if (lane == p) put_row(pr0);
            if (lane == q) put_row(pr1);
            __syncwarp();
            const double a = pr0[p], b = pr0[q], b2 = pr1[p], c2 = pr1[q];
            const double det = a * c2 - b * b2;
            n_neg += (det < 0.0) ? 1 : ((a < 0.0) ? 2 : 0);
            const double idet = fl_rcp(det);
            const double i00 = c2 * idet, i01 = -(b * idet), i10 = -(b2 * idet), i11 = a * idet;
            const bool isp = lane == p, isq = lane == q;
            double ca, cb;
            if (isp) { ca = i00; cb = i01; }
            else if (isq) { ca = i10; cb = i11; }
            else { ca = -(fp * i00 + fq * i10); cb = -(fp * i01 + fq * i11); }
            if (isp || isq)
            {
#pragma unroll
                for (int c = 0; c < NB; ++c) m[c] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < NP; ++i)
            {
                const double2 g0 = reinterpret_cast<const double2*>(pr0)[i];
                const double2 g1 = reinterpret_cast<const double2*>(pr1)[i];
                m[2 * i] = fma(ca, g0.x, fma(cb, g1.x, m[2 * i]));
                if (2 * i + 1 < NB) m[2 * i + 1] = fma(ca, g0.y, fma(cb, g1.y, m[2 * i + 1]));
            }
            if (isp) diag = i00;
            else if (isq) diag = i11;
            else if (lane < NB) diag = fma(ca, pr0[lane], fma(cb, pr1[lane], diag));
            kw_set<NB>(m, p, isp ? i00 : (isq ? i10 : ca));
            kw_set<NB>(m, q, isp ? i01 : (isq ? i11 : cb));
            swept |= (1u << p) | (1u << q);
        }
        __syncwarp();
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) m[c] *= rho;
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NV, int NEQ>
struct kw_stage
{
    static constexpr int NB = NV + NEQ;
    using Sm = kw_smem<NV, NEQ>;
    static constexpr int LDV = Sm::LDV, LDB = Sm::LDB, LDF = Sm::LDF, FT = Sm::FT;
    static constexpr int NVP = (NV + 1) / 2;

    static __device__ __forceinline__ int e_in(const fl_kkt_dev& K, int j) { const int node = j + K.node0; return node == 0 ? K.n_points - 1 : node - 1; }
    // stage-0 row of compact fill row i: the NEQ multiplier rows, then the control couplings (last two variables, direct form)
    static __device__ __forceinline__ int frow(int i) { return i < NEQ ? NV + i : NV - 2 + (i - NEQ); }

    // L2 prefetch of the Hessian / Jacobian values the gathers of stage j will read (slot-major arrays: slot k of stage j sits next to
    // slot k of stage j - 1, four stages per 32-byte sector).  Issued FL_KW_PREFETCH stages ahead: the gathers then find their sectors in
    // L2 instead of DRAM (profiles/prof_r2w_kkt_factor_w: 13.9 GB read from DRAM for 2.8 GB of distinct values, L2 hit rate 25 %).
    static __device__ __forceinline__ void prefetch_stage(const fl_kkt_dev& K, int lap, int j, int lane)
    {
        if (j < 0) return;
        const int e = e_in(K, j), eb = e - K.node0;
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
        for (int k = lane; k < K.NH; k += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(hes + (size_t)k * K.S + j));
        for (int k = lane; k < K.NJA; k += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(jac + (size_t)k * K.n_el + e));
        if (eb >= 0)
            for (int k = lane; k < K.NJB; k += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(jac + (size_t)K.NJA * K.n_el + (size_t)k * K.nB + eb));
    }

    // Row `lane` of the diagonal block of stage j (without Schur terms) into m; inequality rows and 1 / D into sm.aq.
    static __device__ __forceinline__ void gather_diag(const fl_kkt_dev& K, Sm& sm, int lap, int j, double dw, double dc, int lane, double (&m)[NB])
    {
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
#pragma unroll
        for (int c = 0; c < NB; ++c) sm.Z[lane + 32 * c] = 0.0;
        for (int k = lane; k < (K.NINEQ + 1) * 32; k += 32) sm.aq[k] = 0.0;
        __syncwarp();
        if (!K.ls_mode)
            for (int k = lane; k < K.NH; k += 32)
            {
                const int r = K.t.h_row[k], c = K.t.h_col[k];
                const double v = hes[(size_t)k * K.S + j];
                sm.Z[r + 32 * c] = v;
                sm.Z[c + 32 * r] = v;
            }
        for (int k = lane; k < K.NJA; k += 32)
        {
            const int r = K.t.ja_row[k], c = K.t.ja_col[k];
            const double v = jac[(size_t)k * K.n_el + e];
            const int q = K.t.eq_of_row[r];
            if (q >= 0) { sm.Z[NV + q + 32 * c] = v; sm.Z[c + 32 * (NV + q)] = v; }
            else sm.aq[K.t.ineq_of_row[r] * 32 + c] = v;
        }
        if (lane < K.NINEQ)
        {
            const double ss = K.ls_mode ? 1.0 : K.sigs[(size_t)lap * K.n_el * K.NINEQ + (size_t)e * K.NINEQ + lane] + dw;
            sm.aq[K.NINEQ * 32 + lane] = 1.0 / (1.0 / ss + (K.ls_mode ? 0.0 : dc));
        }
        __syncwarp();
        if (lane < NV) sm.Z[lane + 32 * lane] = K.ls_mode ? 1.0 : sm.Z[lane + 32 * lane] + K.sigx[(size_t)lap * K.n + (size_t)j * NV + lane] + dw;
        else if (lane < NB) sm.Z[lane + 32 * lane] = -dc;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < NB; ++c) m[c] = sm.Z[lane + 32 * c];
        // condensed inequality rows: M_xx += A_q^T W A_q
        if (lane < NV)
            for (int q = 0; q < K.NINEQ; ++q)
            {
                const double wq = sm.aq[K.NINEQ * 32 + q] * sm.aq[q * 32 + lane];
#pragma unroll
                for (int i = 0; i < NVP; ++i)
                {
                    const double2 av = reinterpret_cast<const double2*>(sm.aq + q * 32)[i];
                    m[2 * i] = fma(wq, av.x, m[2 * i]);
                    if (2 * i + 1 < NV) m[2 * i + 1] = fma(wq, av.y, m[2 * i + 1]);
                }
            }
    }

    // Compact coupling block K[stage j, stage j-1] (j == 0 of a closed lap: K[0, S-1]) into Lc: rows q < NEQ (multiplier rows
    // of stage j) x NV columns, row NEQ = the control couplings (cp, cp).
    static __device__ __forceinline__ void gather_coupling(const fl_kkt_dev& K, double* __restrict__ Lc, int lap, int j, int lane)
    {
        const int e = e_in(K, j);
        const double* jac = K.jac + (size_t)lap * K.nnz_jac;
        const double* hes = K.hess + (size_t)lap * K.nnz_hess;
        for (int k = lane; k < (NEQ + 1) * LDV; k += 32) Lc[k] = 0.0;
        __syncwarp();
        const int eb = e - K.node0;              // index among the elements with a free left node
        if (eb >= 0)
        {
            for (int k = lane; k < K.NJB; k += 32)
            {
                const int q = K.t.eq_of_row[K.t.jb_row[k]];
                Lc[q * LDV + K.t.jb_col[k]] = jac[(size_t)K.NJA * K.n_el + (size_t)k * K.nB + eb];
            }
            if (!K.ls_mode && lane < K.NCPL) Lc[NEQ * LDV + lane] = hes[(size_t)K.NH * K.S + (size_t)lane * K.nC + eb];
        }
        __syncwarp();
    }
};

template <int NV, int NEQ>
__global__ void __launch_bounds__(32, FL_KW_MINB) k_kkt_factor_w(const fl_kkt_dev K)
{
    using St = kw_stage<NV, NEQ>;
    using Sm = kw_smem<NV, NEQ>;
    constexpr int NB = St::NB, LDV = St::LDV, LDB = St::LDB, LDF = St::LDF, FT = St::FT, NVP = St::NVP;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char kw_raw[];
    Sm& sm = *reinterpret_cast<Sm*>(kw_raw);
    double* const XL = sm.W + Sm::OXL;
    double* const Lc = sm.W + Sm::OLC;
    double* const Ft = sm.W + Sm::OFT;
    const int lap = K.lap_list ? K.lap_list[blockIdx.x] : blockIdx.x, lane = threadIdx.x;
    if (K.active && !K.active[lap]) return;
    fl_ipm_lap* Lp = K.laps ? K.laps + lap : nullptr;
    if (Lp && Lp->status != IPM_RUNNING) return;
    const int S = K.S;
    const bool cyc = K.closed && S > 2;
    const int ncpl = K.ls_mode ? 0 : K.NCPL;
    const int NF = NEQ + ncpl;
    double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    double* gL = K.Lst + (size_t)lap * S * NB * NV;
    double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    const int AQ = K.NINEQ * (NV + 1);
    const double dc = Lp ? Lp->dc : K.dc[lap];
    double dw = Lp ? Lp->dw : K.dw[lap];
    static_assert((NEQ + 1) * LDV <= NB * NV, "the compact coupling rows fit the stage's slot of Lst");
    static_assert(NV * FT <= NB * NB, "the compact fill block fits the stage's slot of F");

    for (int attempt = 0;; ++attempt)
    {
        int n_neg = 0, n_zero = 0;
        bool ok = true;
        int j_fail = -1;
        bool have_P = false;             // the previous stage left a Schur term in sm.aq
        for (int k = lane; k < 16 * 16; k += 32) sm.d0[k] = 0.0;
        if (cyc)
        {
            // F_{S-1} = K[0, S-1]: coupling of stage 0 with its left neighbour (the closing element), compact and transposed
            St::gather_coupling(K, Lc, lap, 0, lane);
            for (int k = lane; k < NV * FT; k += 32) Ft[k] = 0.0;
            __syncwarp();
            if (lane < NEQ)
            {
#pragma unroll
                for (int k = 0; k < NV; ++k) Ft[k * FT + lane] = Lc[lane * LDV + k];
            }
            else if (lane < NF) Ft[(NV - 2 + lane - NEQ) * FT + lane] = Lc[NEQ * LDV + (lane - NEQ)];
            __syncwarp();
        }
        for (int j = S - 1; j >= 0; --j)
        {
            double m[NB];
            double pP[NV];                   // row `lane` (< NV) of the Schur term of the later stages
#pragma unroll
            for (int c = 0; c < NV; ++c) pP[c] = (have_P && lane < NV) ? sm.aq[c * Sm::PS + lane] : 0.0;
            __syncwarp();
            St::gather_diag(K, sm, lap, j, dw, dc, lane, m);
#if FL_KW_PREFETCH > 0
            St::prefetch_stage(K, lap, j - FL_KW_PREFETCH, lane);
#endif
            // keep the inequality rows and 1/D for the solves
            for (int k = lane; k < K.NINEQ * NV; k += 32) gA[(size_t)j * AQ + k] = sm.aq[(k / NV) * 32 + (k % NV)];
            if (lane < K.NINEQ) gA[(size_t)j * AQ + K.NINEQ * NV + lane] = sm.aq[K.NINEQ * 32 + lane];
            // Schur complement of the later stages
            if (lane < NV)
            {
#pragma unroll
                for (int c = 0; c < NV; ++c) m[c] -= pP[c];
            }
            if (j == 0 && cyc && lane < NB)
            {
#pragma unroll
                for (int c = 0; c < NB; ++c) m[c] -= sm.W[lane + 32 * c];
            }
            // the diagonal entry of the row into its side register (through the staging matrix: no run-time register index)
            __syncwarp();
            if (lane < NB)
            {
#pragma unroll
                for (int c = 0; c < NB; ++c) sm.Z[lane + 32 * c] = m[c];
            }
            __syncwarp();
            const double diag = (lane < NB) ? sm.Z[lane + 32 * lane] : 0.0;
            __syncwarp();
            const bool chain = j >= 1 && (j >= 2 || !cyc);          // stage j hands a Schur term to stage j-1 through L_j
            if (j >= 1)
            {
                St::gather_coupling(K, Lc, lap, j, lane);
                for (int k = lane; k < (NEQ + 1) * LDV; k += 32) KW_STORE(&gL[(size_t)j * NB * NV + k], Lc[k]);
            }
            if (cyc && j == 1)
            {
                // K[0, 1] = L_1^T joins the fill block, which becomes a full NB x NB block F1[r0][c1] in the staging matrix
#pragma unroll
                for (int c = 0; c < NB; ++c) sm.Z[lane + 32 * c] = 0.0;
                __syncwarp();
                if (lane < NF)
                {
                    const int r0 = St::frow(lane);
#pragma unroll
                    for (int k = 0; k < NV; ++k) sm.Z[r0 + 32 * k] = Ft[k * FT + lane];
                }
                if (lane < NV)
                {
#pragma unroll
                    for (int q = 0; q < NEQ; ++q) sm.Z[lane + 32 * (NV + q)] = Lc[q * LDV + lane];
                }
                __syncwarp();
                if (lane < ncpl) { const int cp = NV - 2 + lane; sm.Z[cp + 32 * cp] += Lc[NEQ * LDV + lane]; }
                __syncwarp();
                if (lane < NB)
                {
#pragma unroll
                    for (int c = 0; c < NB; ++c) gF[(size_t)NB * NB + lane + NB * c] = sm.Z[lane + 32 * c];
                }
            }
            else if (cyc && j >= 2)
                for (int k = lane; k < NV * FT; k += 32) KW_STORE(&gF[(size_t)j * NB * NB + k], Ft[k]);
            const int neg_before = n_neg, zero_before = n_zero;
            kw_invert<NB>(m, diag, sm.pr, lane, n_neg, n_zero);
            if (Lp)
            {
                // every pivot block carries exactly NEQ negative pivots when the reduced Hessian is positive definite
                ok = (n_neg - neg_before == NEQ && n_zero == zero_before);
                if (!ok) { j_fail = j; break; }
            }
            if (lane < NB)
            {
#pragma unroll
                for (int c = 0; c < NB; ++c) KW_STORE(&gD[(size_t)j * NB * NB + lane + NB * c], m[c]);
            }
            if (j == 0) break;
            if (cyc && j == 1)
            {
                // XF1 = D^-1 F1^T (NB x NB), then d0full = F1 XF1 + the compact terms of the earlier stages, left in W for stage 0.
                // Both products run in two halves of the columns: this stage is visited once per lap, the accumulators of a full
                // row would set the register count of the whole kernel.
                constexpr int HP = ((NB + 1) / 2 + 1) / 2;                 // 128-bit words per half: columns [0, 2 HP) and [2 HP, 4 HP)
                __syncwarp();
#pragma unroll 1
                for (int h = 0; h < 2; ++h)
                {
                    const int c0 = h * 2 * HP;                             // first column of the half (even)
                    double acc[2 * HP];
#pragma unroll
                    for (int c = 0; c < 2 * HP; ++c) acc[c] = 0.0;
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        const double a = m[k];
#pragma unroll
                        for (int i = 0; i < HP; ++i)
                        {
                            const double2 fv = *reinterpret_cast<const double2*>(sm.Z + 32 * k + c0 + 2 * i);       // F1[c0 + 2i ..][k]
                            acc[2 * i] = fma(a, fv.x, acc[2 * i]);
                            acc[2 * i + 1] = fma(a, fv.y, acc[2 * i + 1]);
                        }
                    }
                    if (lane < NB)
                    {
#pragma unroll
                        for (int i = 0; i < HP; ++i)
                            if (c0 + 2 * i < LDB) *reinterpret_cast<double2*>(sm.W + lane * LDB + c0 + 2 * i) = make_double2(acc[2 * i], acc[2 * i + 1]);
                    }
                }
                __syncwarp();
                // W holds XF1 until both halves are done: the rows of d0full go to global scratch first -- the stage-0 slot of the
                // fill array, which the solves never read -- and come back once W is free
                double* d0g = gF;                                           // [NB][NB] column-major
#pragma unroll 1
                for (int h = 0; h < 2; ++h)
                {
                    const int c0 = h * 2 * HP;
                    double acc[2 * HP];
#pragma unroll
                    for (int c = 0; c < 2 * HP; ++c) acc[c] = 0.0;
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        const double a = (lane < NB) ? sm.Z[lane + 32 * k] : 0.0;                          // F1[lane][k]
#pragma unroll
                        for (int i = 0; i < HP; ++i)
                        {
                            const double2 xv = *reinterpret_cast<const double2*>(sm.W + k * LDB + c0 + 2 * i);       // XF1[k][c0 + 2i ..]
                            acc[2 * i] = fma(a, xv.x, acc[2 * i]);
                            acc[2 * i + 1] = fma(a, xv.y, acc[2 * i + 1]);
                        }
                    }
                    if (lane < NB)
                    {
#pragma unroll
                        for (int c = 0; c < 2 * HP; ++c)
                            if (c0 + c < NB) d0g[lane + NB * (c0 + c)] = acc[c];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NB; ++c) sm.W[lane + 32 * c] = (lane < NB) ? d0g[lane + NB * c] : 0.0;
                __syncwarp();
                if (lane < NF)
                {
                    const int r0 = St::frow(lane);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < NF) sm.W[r0 + 32 * St::frow(i)] += sm.d0[i * 16 + lane];
                }
                have_P = false;                                             // stage 1 of a closed lap hands nothing to stage 0 through L
                __syncwarp();
                continue;
            }
            // The products below are loops over the inner index with their operands in shared memory (the row of D^-1 goes to the
            // staging matrix, row-major): unrolled on the register row they were 1700 instructions of straight-line code per stage,
            // and the warps of an SM, each at another place of 110 KB of code, stalled on instruction fetch (ncu: one
            // no-instruction stall cycle per issued instruction).
            if (lane < NB)
            {
#pragma unroll
                for (int i = 0; i < NB / 2; ++i) reinterpret_cast<double2*>(sm.Z + lane * LDB)[i] = make_double2(m[2 * i], m[2 * i + 1]);
                if (NB & 1) sm.Z[lane * LDB + NB - 1] = m[NB - 1];
            }
            __syncwarp();
            const double* Dr = sm.Z + (lane < NB ? lane : 0) * LDB;          // own row of D^-1
            // XL = D^-1 L (NB x NV): the NEQ multiplier rows of L and the two control couplings
            if (chain)
            {
                double xl[2 * NVP];
#pragma unroll
                for (int c = 0; c < 2 * NVP; ++c) xl[c] = 0.0;
#pragma unroll 1
                for (int q = 0; q < NEQ; ++q)
                {
                    const double a = Dr[NV + q];
#pragma unroll
                    for (int i = 0; i < NVP; ++i)
                    {
                        const double2 lv = reinterpret_cast<const double2*>(Lc + q * LDV)[i];
                        xl[2 * i] = fma(a, lv.x, xl[2 * i]);
                        xl[2 * i + 1] = fma(a, lv.y, xl[2 * i + 1]);
                    }
                }
                if (ncpl > 0)
                {
                    xl[NV - 2] = fma(Dr[NV - 2], Lc[NEQ * LDV + 0], xl[NV - 2]);
                    xl[NV - 1] = fma(Dr[NV - 1], Lc[NEQ * LDV + 1], xl[NV - 1]);
                }
                if (lane < NB)
                {
#pragma unroll
                    for (int i = 0; i < NVP; ++i) reinterpret_cast<double2*>(XL + lane * LDV)[i] = make_double2(xl[2 * i], xl[2 * i + 1]);
                }
            }
            // XF = D^-1 F^T (NB x NF): every lane needs its own row of D^-1 only, and overwrites it with its row of XF
            if (cyc)
            {
                double xf[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) xf[i] = 0.0;
#pragma unroll 1
                for (int k = 0; k < NV; ++k)
                {
                    const double a = Dr[k];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const double2 fv = reinterpret_cast<const double2*>(Ft + k * FT)[i];
                        xf[2 * i] = fma(a, fv.x, xf[2 * i]);
                        xf[2 * i + 1] = fma(a, fv.y, xf[2 * i + 1]);
                    }
                }
                if (lane < NB)
                {
#pragma unroll
                    for (int i = 0; i < 8; ++i) reinterpret_cast<double2*>(sm.Z + lane * LDB)[i] = make_double2(xf[2 * i], xf[2 * i + 1]);
                }
            }
            __syncwarp();
            // P = L^T XL (what stage j-1 subtracts from its x-x block), parked in sm.aq
            have_P = chain;
            if (chain)
            {
                double acc[2 * NVP];
#pragma unroll
                for (int c = 0; c < 2 * NVP; ++c) acc[c] = 0.0;
#pragma unroll 1
                for (int q = 0; q < NEQ; ++q)
                {
                    const double a = (lane < NV) ? Lc[q * LDV + lane] : 0.0;
#pragma unroll
                    for (int i = 0; i < NVP; ++i)
                    {
                        const double2 xv = reinterpret_cast<const double2*>(XL + (NV + q) * LDV)[i];
                        acc[2 * i] = fma(a, xv.x, acc[2 * i]);
                        acc[2 * i + 1] = fma(a, xv.y, acc[2 * i + 1]);
                    }
                }
                if (ncpl > 0)
                {
#pragma unroll 1
                    for (int t = 0; t < 2; ++t)
                    {
                        const double a = (lane == NV - 2 + t) ? Lc[NEQ * LDV + t] : 0.0;
#pragma unroll
                        for (int i = 0; i < NVP; ++i)
                        {
                            const double2 xv = reinterpret_cast<const double2*>(XL + (NV - 2 + t) * LDV)[i];
                            acc[2 * i] = fma(a, xv.x, acc[2 * i]);
                            acc[2 * i + 1] = fma(a, xv.y, acc[2 * i + 1]);
                        }
                    }
                }
                if (lane < NV)
                {
#pragma unroll
                    for (int c = 0; c < NV; ++c) sm.aq[c * Sm::PS + lane] = acc[c];
                }
            }
            if (cyc)
            {
                // stage 0: d0 += F XF ; new fill F_{j-1} = -F XL (columns: x of stage j-1), lane = compact fill row
                double fn[2 * NVP], d0[16];
#pragma unroll
                for (int c = 0; c < 2 * NVP; ++c) fn[c] = 0.0;
#pragma unroll
                for (int c = 0; c < 16; ++c) d0[c] = (lane < 16) ? sm.d0[c * 16 + lane] : 0.0;
#pragma unroll 1
                for (int k = 0; k < NV; ++k)
                {
                    const double a = (lane < 16) ? Ft[k * FT + lane] : 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const double2 xv = reinterpret_cast<const double2*>(sm.Z + k * LDB)[i];
                        d0[2 * i] = fma(a, xv.x, d0[2 * i]);
                        d0[2 * i + 1] = fma(a, xv.y, d0[2 * i + 1]);
                    }
                    if (chain)
                    {
#pragma unroll
                        for (int i = 0; i < NVP; ++i)
                        {
                            const double2 xv = reinterpret_cast<const double2*>(XL + k * LDV)[i];
                            fn[2 * i] = fma(-a, xv.x, fn[2 * i]);
                            fn[2 * i + 1] = fma(-a, xv.y, fn[2 * i + 1]);
                        }
                    }
                }
                __syncwarp();
                if (lane < 16)
                {
#pragma unroll
                    for (int c = 0; c < NV; ++c) Ft[c * FT + lane] = fn[c];
#pragma unroll
                    for (int c = 0; c < 16; ++c) sm.d0[c * 16 + lane] = d0[c];
                }
            }
            __syncwarp();
        }
        if (!Lp)
        {
            if (lane == 0) { K.n_neg[lap] = n_neg; K.n_zero[lap] = n_zero; }
            return;
        }
        // Algorithm IC (Waechter & Biegler, section 3.1); every lane takes the same decision from the same data, lane 0 records it
        const fl_ipm_options& o = K.o;
        int fl;
        if (lane == 0)
        {
            if (ok) Lp->stages_ok += S;
            else
            {
                Lp->stages_fail += S - j_fail;
                if (Lp->n_fail > 0) { int d = j_fail - Lp->fail_stage; if (d < 0) d = -d; if (d > S - d) d = S - d; Lp->fail_shift += d; }
                Lp->fail_stage = j_fail; Lp->n_fail += 1;
            }
        }
        if (ok)
        {
            if (lane == 0)
            {
                Lp->n_fact += 1;
                if (dw > 0.0) Lp->dw_last = dw;
                Lp->dw = dw;
                Lp->dw_prev = dw;
            }
            fl = 1;
        }
        else
        {
            const double dw_last = Lp->dw_last;
            if (dw == 0.0) dw = (dw_last == 0.0) ? o.delta_w_0 : fmax(o.delta_w_min, o.kappa_w_minus * dw_last);
            else dw *= (dw_last == 0.0) ? o.kappa_w_plus_bar : o.kappa_w_plus;
            fl = (dw > o.delta_w_max) ? 2 : 0;
            if (lane == 0)
            {
                Lp->n_fact += 1;
                Lp->n_reg += 1;
                if (fl == 2) Lp->status = IPM_REGULARISATION_FAILED;
            }
        }
        if (fl != 0) return;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Solves K z = (r1, r2) for every active lap with the factors of k_kkt_factor_w; dx [n], dy [m] in the NLP's own orders.
// One warp per lap, lane r = row r of the stage; the next stage's D^-1 row is loaded while the current stage is reduced.
#ifndef FL_KWS_MINB
#define FL_KWS_MINB 1
#endif
// stages ahead of which the solve kernel asks the L2 for a stage's factor blocks (0: off).  The kernel streams 4.3 KB per stage and lap
// and waits for it (9.8 long-scoreboard stall cycles per issue, profiles/prof_r2w_kkt_solve_w): the register prefetch one stage ahead then
// finds its lines in L2 instead of DRAM.
#ifndef FL_KWS_PREFETCH
#define FL_KWS_PREFETCH 3
#endif
template <int NV, int NEQ>
__global__ void __launch_bounds__(32, FL_KWS_MINB) k_kkt_solve_w(const fl_kkt_dev K, const double* __restrict__ r1, const double* __restrict__ r2,
                                                     double* __restrict__ dx, double* __restrict__ dy)
{
    using St = kw_stage<NV, NEQ>;
    constexpr int NB = St::NB, LDV = St::LDV, FT = St::FT, NVP = St::NVP;
    constexpr int NBP = (NB + 1) / 2;
    __shared__ __align__(16) double vb[32], vy[32], z0[32], z0c[16];
    const int lap = K.lap_list ? K.lap_list[blockIdx.x] : blockIdx.x, lane = threadIdx.x;
    if (K.active && !K.active[lap]) return;
    if (K.laps && K.laps[lap].status != IPM_RUNNING) return;
    const int S = K.S;
    const bool cyc = K.closed && S > 2;
    const int ncpl = K.ls_mode ? 0 : K.NCPL;
    const int NF = NEQ + ncpl;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    const double* gL = K.Lst + (size_t)lap * S * NB * NV;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    r1 += (size_t)lap * K.n; r2 += (size_t)lap * K.m; dx += (size_t)lap * K.n; dy += (size_t)lap * K.m;
    const int AQ = K.NINEQ * (NV + 1);

    // condensed right-hand side of stage j (before the updates from later stages); lanes < NB
    auto rhs = [&](int j) -> double {
        const int e = St::e_in(K, j);
        double b = 0.0;
        if (lane < NV)
        {
            b = r1[(size_t)j * NV + lane];
            for (int q = 0; q < K.NINEQ; ++q)
                b += gA[(size_t)j * AQ + q * NV + lane] * gA[(size_t)j * AQ + K.NINEQ * NV + q] * r2[(size_t)e * K.NCPE + K.t.ineq_rows[q]];
        }
        else if (lane < NB) b = r2[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]];
        return b;
    };
    auto load_D = [&](int j, double (&reg)[NB]) {                     // reg[c] = D_j^-1[lane][c]
#pragma unroll
        for (int c = 0; c < NB; ++c) reg[c] = (lane < NB) ? gD[(size_t)j * NB * NB + lane + NB * c] : 0.0;
    };
    // y = D^-1 v with v in shared memory (broadcast reads)
    auto matvec = [&](const double (&reg)[NB], const double* v) -> double {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < NBP; ++i)
        {
            const double2 vv = reinterpret_cast<const double2*>(v)[i];
            a0 = fma(reg[2 * i], vv.x, a0);
            if (2 * i + 1 < NB) a1 = fma(reg[2 * i + 1], vv.y, a1);
        }
        return a0 + a1;
    };

    auto prefetch_stage = [&](int j) {
        if (j < 0 || j >= S) return;
        auto lines = [&](const double* base, int bytes) {
            const char* p0 = reinterpret_cast<const char*>(base);
            for (int o = lane * 128; o < bytes + 127; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + (o < bytes ? o : bytes - 1)));
        };
        lines(gD + (size_t)j * NB * NB, NB * NB * 8);
        lines(gL + (size_t)j * NB * NV, (NEQ + 1) * LDV * 8);
        if (cyc && j >= 2) lines(gF + (size_t)j * NB * NB, NV * FT * 8);
    };
    // ---- forward: y_j = Dinv_j b_j,  b_{j-1} -= L_j^T y_j,  b_0 -= F_j y_j
    double carry = 0.0;            // lanes < NV
    double b0c = 0.0;              // lanes < NF: compact rows of the update of stage 0
    double b0f = 0.0;              // lanes < NB: full update of stage 0 (stage 1 of closed laps)
    double dreg[NB];
    if (S > 1) load_D(S - 1, dreg);
    for (int j = S - 1; j >= 1; --j)
    {
        const bool chain = j >= 2 || !cyc;
#if FL_KWS_PREFETCH > 0
        prefetch_stage(j - 1 - FL_KWS_PREFETCH);
#endif
        vb[lane] = (lane < NB) ? rhs(j) + carry : 0.0;
        // operands of this stage's updates, and the next stage's D^-1 row, in flight while the product is reduced
        double lcol[NEQ + 1];
#pragma unroll
        for (int q = 0; q < NEQ; ++q) lcol[q] = (chain && lane < NV) ? gL[(size_t)j * NB * NV + q * LDV + lane] : 0.0;
        lcol[NEQ] = (chain && ncpl > 0 && lane >= NV - 2 && lane < NV) ? gL[(size_t)j * NB * NV + NEQ * LDV + (lane - (NV - 2))] : 0.0;
        double fcol[NV];
        if (cyc && j >= 2)
        {
#pragma unroll
            for (int k = 0; k < NV; ++k) fcol[k] = (lane < 16) ? gF[(size_t)j * NB * NB + k * FT + lane] : 0.0;
        }
        double dn[NB];
        if (j >= 2) load_D(j - 1, dn);
        else load_D(0, dn);
        __syncwarp();
        const double y = matvec(dreg, vb);
        __syncwarp();
        vy[lane] = y;
        if (lane < NB) yw[(size_t)j * NB + lane] = y;
        __syncwarp();
        if (chain)
        {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < NEQ; ++q) s = fma(lcol[q], vy[NV + q], s);
            s = fma(lcol[NEQ], y, s);                                   // control coupling (cp, cp): own entry of y
            carry = (lane < NV) ? -s : 0.0;
        }
        else carry = 0.0;
        if (cyc)
        {
            if (j >= 2)
            {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < NV; ++k) s = fma(fcol[k], vy[k], s);
                b0c -= s;
            }
            else
            {
                double s = 0.0;
                if (lane < NB)
                    for (int k = 0; k < NB; ++k) s = fma(gF[(size_t)NB * NB + lane + NB * k], vy[k], s);
                b0f -= s;
            }
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) dreg[c] = dn[c];
        __syncwarp();
    }
    {
        if (S == 1) load_D(0, dreg);
        // stage 0: the compact rows of the fill updates go to their rows
        vb[lane] = 0.0;
        __syncwarp();
        if (cyc && lane < NF) vb[St::frow(lane)] = b0c;
        __syncwarp();
        const double b = (lane < NB) ? rhs(0) + carry + b0f + vb[lane] : 0.0;
        __syncwarp();
        vb[lane] = b;
        __syncwarp();
        const double z = matvec(dreg, vb);
        __syncwarp();
        z0[lane] = z; vy[lane] = z;
        if (lane < NB) yw[lane] = z;
        __syncwarp();
        if (lane < 16) z0c[lane] = (lane < NF) ? z0[St::frow(lane)] : 0.0;
        __syncwarp();
    }
    // ---- backward: z_j = y_j - Dinv_j (L_j z_{j-1} + F_j^T z_0); vy holds z_{j-1}
    auto scatter = [&](int j, double z) {
        // dx, equality multipliers, then the condensed inequality multipliers dy_q = (a_q . dx_j - r2_q) / D_q (vy holds z_j)
        const int e = St::e_in(K, j);
        if (lane < NV) dx[(size_t)j * NV + lane] = z;
        else if (lane < NB) dy[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]] = z;
        if (lane < K.NINEQ)
        {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < NV; ++c) a = fma(gA[(size_t)j * AQ + lane * NV + c], vy[c], a);
            const int row = e * K.NCPE + K.t.ineq_rows[lane];
            dy[row] = (a - r2[row]) * gA[(size_t)j * AQ + K.NINEQ * NV + lane];
        }
    };
    scatter(0, vy[lane]);
    if (S > 1) load_D(1, dreg);
    for (int j = 1; j < S; ++j)
    {
        const bool chain = j >= 2 || !cyc;
#if FL_KWS_PREFETCH > 0
        prefetch_stage(j + 1 + FL_KWS_PREFETCH);
#endif
        // t = L_j z_{j-1} + F_j^T z_0
        double t = 0.0;
        if (chain)
        {
            if (lane >= NV && lane < NB)
            {
                const double* lr = gL + (size_t)j * NB * NV + (lane - NV) * LDV;
#pragma unroll
                for (int i = 0; i < NVP; ++i)
                {
                    const double2 lv = reinterpret_cast<const double2*>(lr)[i];
                    t = fma(lv.x, vy[2 * i], t);
                    if (2 * i + 1 < NV) t = fma(lv.y, vy[2 * i + 1], t);
                }
            }
            else if (ncpl > 0 && lane >= NV - 2 && lane < NV)
                t = gL[(size_t)j * NB * NV + NEQ * LDV + (lane - (NV - 2))] * vy[lane];
        }
        if (cyc)
        {
            if (j >= 2)
            {
                if (lane < NV)
                {
                    const double* fr = gF + (size_t)j * NB * NB + lane * FT;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const double2 fv = reinterpret_cast<const double2*>(fr)[i];
                        t = fma(fv.x, z0c[2 * i], t);
                        t = fma(fv.y, z0c[2 * i + 1], t);
                    }
                }
            }
            else if (lane < NB)
                for (int k = 0; k < NB; ++k) t = fma(gF[(size_t)NB * NB + k + NB * lane], z0[k], t);
        }
        double dn[NB];
        if (j + 1 < S) load_D(j + 1, dn);
        const double yj = (lane < NB) ? yw[(size_t)j * NB + lane] : 0.0;
        __syncwarp();
        vb[lane] = t;
        __syncwarp();
        const double z = yj - matvec(dreg, vb);
        __syncwarp();
        vy[lane] = z;
        __syncwarp();
        scatter(j, z);
        if (j + 1 < S)
        {
#pragma unroll
            for (int c = 0; c < NB; ++c) dreg[c] = dn[c];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same solve with the factor blocks of a stage (D^-1, compact coupling rows, compact fill) brought into a ring of shared-memory
// slots by bulk asynchronous copies (cp.async.bulk, the 1-D TMA path; completion on one mbarrier per slot), FL_KWS_SLOTS - 1 stages
// ahead of their use, instead of the register prefetch of ONE stage: the kernel waits for its own streaming loads (9.8 long-scoreboard
// stall cycles per issue, profiles/prof_r2w_kkt_solve_w), so what counts is the bytes in flight per SM.  For stage sizes whose blocks
// are multiples of 16 bytes (NB even); the arithmetic and its order are those of k_kkt_solve_w.
// slots of the ring.  Measured on a 4096-lap, 40-iteration solve (solve phase; k_kkt_solve_w with its register prefetch: 937 ms):
// 2 slots (20 KB per lap, 11 laps per SM) 643 ms, 3 slots (7 per SM) 707-717 ms, 4 slots (5 per SM) 883 ms
#ifndef FL_KWS_SLOTS
#define FL_KWS_SLOTS 2
#endif
template <int NV, int NEQ> struct kws_ring
{
    static constexpr int NB = NV + NEQ, LDV = kw_stage<NV, NEQ>::LDV, FT = kw_stage<NV, NEQ>::FT;
    static constexpr int DB = NB * NB, LB = (NEQ + 1) * LDV, FB = NV * FT, SLOT = DB + LB + FB;      // doubles
    static constexpr bool ok = (DB % 2 == 0) && (LB % 2 == 0) && (FB % 2 == 0) && ((NB * NV) % 2 == 0);
    static constexpr size_t bytes = (size_t)FL_KWS_SLOTS * SLOT * 8 + FL_KWS_SLOTS * 8 + (32 + 32 + 32 + 16) * 8;
};

template <int NV, int NEQ>
__global__ void __launch_bounds__(32) k_kkt_solve_w2(const fl_kkt_dev K, const double* __restrict__ r1, const double* __restrict__ r2,
                                                     double* __restrict__ dx, double* __restrict__ dy)
{
    using St = kw_stage<NV, NEQ>;
    using R = kws_ring<NV, NEQ>;
    constexpr int NB = St::NB, LDV = St::LDV, FT = St::FT, NVP = St::NVP, NS = FL_KWS_SLOTS;
    constexpr int NBP = (NB + 1) / 2;
    extern __shared__ __align__(128) unsigned char kws_raw[];
    double* const ring = reinterpret_cast<double*>(kws_raw);
    double* const vb = ring + (size_t)NS * R::SLOT;          // (16-byte aligned: read as double2)
    double* const vy = vb + 32;
    double* const z0 = vy + 32;
    double* const z0c = z0 + 32;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(z0c + 16);
    const int lap = K.lap_list ? K.lap_list[blockIdx.x] : blockIdx.x, lane = threadIdx.x;
    if (K.active && !K.active[lap]) return;
    if (K.laps && K.laps[lap].status != IPM_RUNNING) return;
    const int S = K.S;
    const bool cyc = K.closed && S > 2;
    const int ncpl = K.ls_mode ? 0 : K.NCPL;
    const int NF = NEQ + ncpl;
    const double* gD = K.Dinv + (size_t)lap * S * NB * NB;
    const double* gF = K.F ? K.F + (size_t)lap * S * NB * NB : nullptr;
    const double* gL = K.Lst + (size_t)lap * S * NB * NV;
    const double* gA = K.Aq + (size_t)lap * S * (K.NINEQ * (NV + 1));
    double* yw = K.ywork + (size_t)lap * S * NB;
    r1 += (size_t)lap * K.n; r2 += (size_t)lap * K.m; dx += (size_t)lap * K.n; dy += (size_t)lap * K.m;
    const int AQ = K.NINEQ * (NV + 1);
    const bool haveF = cyc && gF != nullptr;

    // ---- the ring ------------------------------------------------------------------------------------------------------
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
    if (lane == 0)
    {
        for (int k = 0; k < NS; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8 * k));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0u;                                   // bit k: parity the next wait on slot k expects
    auto issue = [&](int j, int slot) {                    // (lane 0) stage j's blocks -> slot
        if (lane != 0) return;
        const unsigned bar = bar0 + 8 * slot;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (size_t)slot * R::SLOT);
        const unsigned bytes = (unsigned)((R::DB + R::LB + (haveF ? R::FB : 0)) * 8);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the slot's last readers (generic proxy) come first
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(gD + (size_t)j * NB * NB), "r"((unsigned)(R::DB * 8)), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst + R::DB * 8), "l"(gL + (size_t)j * NB * NV), "r"((unsigned)(R::LB * 8)), "r"(bar) : "memory");
        if (haveF)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst + (R::DB + R::LB) * 8), "l"(gF + (size_t)j * NB * NB), "r"((unsigned)(R::FB * 8)), "r"(bar) : "memory");
    };
    auto wait = [&](int slot) {
        const unsigned bar = bar0 + 8 * slot, par = (phase >> slot) & 1u;
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(par) : "memory");
        phase ^= 1u << slot;
    };

    // condensed right-hand side of stage j (before the updates from later stages); lanes < NB
    auto rhs = [&](int j) -> double {
        const int e = St::e_in(K, j);
        double b = 0.0;
        if (lane < NV)
        {
            b = r1[(size_t)j * NV + lane];
            for (int q = 0; q < K.NINEQ; ++q)
                b += gA[(size_t)j * AQ + q * NV + lane] * gA[(size_t)j * AQ + K.NINEQ * NV + q] * r2[(size_t)e * K.NCPE + K.t.ineq_rows[q]];
        }
        else if (lane < NB) b = r2[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]];
        return b;
    };
    // y = D^-1 v: row `lane` of D^-1 from the slot (or from global memory: stage 0), v in shared memory (broadcast reads); the same two
    // accumulation chains as k_kkt_solve_w
    auto matvec = [&](const double* D, const double* v) -> double {
        double a0 = 0.0, a1 = 0.0;
        if (lane < NB)
        {
#pragma unroll
            for (int i = 0; i < NBP; ++i)
            {
                const double2 vv = reinterpret_cast<const double2*>(v)[i];
                a0 = fma(D[lane + NB * (2 * i)], vv.x, a0);
                if (2 * i + 1 < NB) a1 = fma(D[lane + NB * (2 * i + 1)], vv.y, a1);
            }
        }
        return a0 + a1;
    };

    // ---- forward: y_j = Dinv_j b_j,  b_{j-1} -= L_j^T y_j,  b_0 -= F_j y_j
    double carry = 0.0;            // lanes < NV
    double b0c = 0.0;              // lanes < NF: compact rows of the update of stage 0
    double b0f = 0.0;              // lanes < NB: full update of stage 0 (stage 1 of closed laps)
    for (int i = 0; i < NS && S - 1 - i >= 1; ++i) issue(S - 1 - i, i);
    for (int j = S - 1; j >= 1; --j)
    {
        const int slot = (S - 1 - j) % NS;
        const double* sD = ring + (size_t)slot * R::SLOT;
        const double* sL = sD + R::DB;
        const double* sF = sL + R::LB;
        const bool chain = j >= 2 || !cyc;
        vb[lane] = (lane < NB) ? rhs(j) + carry : 0.0;
        wait(slot);
        __syncwarp();
        const double y = matvec(sD, vb);
        __syncwarp();
        vy[lane] = y;
        if (lane < NB) yw[(size_t)j * NB + lane] = y;
        __syncwarp();
        if (chain)
        {
            double s = 0.0;
            if (lane < NV)
            {
#pragma unroll
                for (int q = 0; q < NEQ; ++q) s = fma(sL[q * LDV + lane], vy[NV + q], s);
                if (ncpl > 0 && lane >= NV - 2) s = fma(sL[NEQ * LDV + (lane - (NV - 2))], y, s);      // control coupling (cp, cp): own entry of y
            }
            carry = (lane < NV) ? -s : 0.0;
        }
        else carry = 0.0;
        if (cyc)
        {
            if (j >= 2)
            {
                double s = 0.0;
                if (lane < 16)
                {
#pragma unroll
                    for (int k = 0; k < NV; ++k) s = fma(sF[k * FT + lane], vy[k], s);
                }
                b0c -= s;
            }
            else
            {
                double s = 0.0;
                if (lane < NB)
                    for (int k = 0; k < NB; ++k) s = fma(gF[(size_t)NB * NB + lane + NB * k], vy[k], s);
                b0f -= s;
            }
        }
        __syncwarp();
        if (j - NS >= 1) issue(j - NS, slot);
    }
    {
        // stage 0: the compact rows of the fill updates go to their rows
        vb[lane] = 0.0;
        __syncwarp();
        if (cyc && lane < NF) vb[St::frow(lane)] = b0c;
        __syncwarp();
        const double b = (lane < NB) ? rhs(0) + carry + b0f + vb[lane] : 0.0;
        __syncwarp();
        vb[lane] = b;
        __syncwarp();
        const double z = matvec(gD, vb);
        __syncwarp();
        z0[lane] = z; vy[lane] = z;
        if (lane < NB) yw[lane] = z;
        __syncwarp();
        if (lane < 16) z0c[lane] = (lane < NF) ? z0[St::frow(lane)] : 0.0;
        __syncwarp();
    }
    // ---- backward: z_j = y_j - Dinv_j (L_j z_{j-1} + F_j^T z_0); vy holds z_{j-1}
    auto scatter = [&](int j, double z) {
        // dx, equality multipliers, then the condensed inequality multipliers dy_q = (a_q . dx_j - r2_q) / D_q (vy holds z_j)
        const int e = St::e_in(K, j);
        if (lane < NV) dx[(size_t)j * NV + lane] = z;
        else if (lane < NB) dy[(size_t)e * K.NCPE + K.t.eq_rows[lane - NV]] = z;
        if (lane < K.NINEQ)
        {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < NV; ++c) a = fma(gA[(size_t)j * AQ + lane * NV + c], vy[c], a);
            const int row = e * K.NCPE + K.t.ineq_rows[lane];
            dy[row] = (a - r2[row]) * gA[(size_t)j * AQ + K.NINEQ * NV + lane];
        }
    };
    scatter(0, vy[lane]);
    for (int i = 0; i < NS && 1 + i < S; ++i) issue(1 + i, i);
    for (int j = 1; j < S; ++j)
    {
        const int slot = (j - 1) % NS;
        const double* sD = ring + (size_t)slot * R::SLOT;
        const double* sL = sD + R::DB;
        const double* sF = sL + R::LB;
        const bool chain = j >= 2 || !cyc;
        const double yj = (lane < NB) ? yw[(size_t)j * NB + lane] : 0.0;
        wait(slot);
        // t = L_j z_{j-1} + F_j^T z_0
        double t = 0.0;
        if (chain)
        {
            if (lane >= NV && lane < NB)
            {
                const double* lr = sL + (lane - NV) * LDV;
#pragma unroll
                for (int i = 0; i < NVP; ++i)
                {
                    const double2 lv = reinterpret_cast<const double2*>(lr)[i];
                    t = fma(lv.x, vy[2 * i], t);
                    if (2 * i + 1 < NV) t = fma(lv.y, vy[2 * i + 1], t);
                }
            }
            else if (ncpl > 0 && lane >= NV - 2 && lane < NV)
                t = sL[NEQ * LDV + (lane - (NV - 2))] * vy[lane];
        }
        if (cyc)
        {
            if (j >= 2)
            {
                if (lane < NV)
                {
                    const double* fr = sF + lane * FT;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const double2 fv = reinterpret_cast<const double2*>(fr)[i];
                        t = fma(fv.x, z0c[2 * i], t);
                        t = fma(fv.y, z0c[2 * i + 1], t);
                    }
                }
            }
            else if (lane < NB)
                for (int k = 0; k < NB; ++k) t = fma(gF[(size_t)NB * NB + k + NB * lane], z0[k], t);
        }
        __syncwarp();
        vb[lane] = t;
        __syncwarp();
        const double z = yj - matvec(sD, vb);
        __syncwarp();
        vy[lane] = z;
        __syncwarp();
        scatter(j, z);
        __syncwarp();
        if (j + NS < S) issue(j + NS, slot);
    }
}
