// Collocation NLP callbacks on sm_100a: kernel templates over a generated node program (csrc/generated/gen_*.cuh).
//
// What the reference computes by sweeping one CppAD tape over all mesh nodes on one CPU thread
// (Optimal_laptime::FG_direct / FG_derivative::compute, src/core/applications/optimal_laptime.hpp:1226-1657, evaluated
// through lion::ipopt_cppad_solve, :546-551) is split here into three data-parallel passes, one thread per collocation
// node (or element), fp64 throughout:
//
//   k_node_values   x_j -> node values V (states, s-derivatives, algebraic rows, extra constraints, dt/ds) and the node's
//                   checkpoints C (every sin/cos/atan/tanh/exp/reciprocal/sqrt the derivative code needs)
//   k_assemble      elements: trapezoidal / algebraic / extra / control-rate rows g, and the objective f
//   k_node_derivs   x_j, lambda, obj_factor -> grad f, the node's two Jacobian blocks, its Lagrangian-Hessian block
//                   (thread = node, blockIdx.y = one of N balanced partitions of the node's derivative outputs)
//
// Layout in HBM (all fp64):
//   xfull  [problem][node][NV]      node-major, the reference's variable order (optimal_laptime.hpp:1269-1303); the fixed
//                                   first node of open problems occupies slot 0 and is filled by the host
//   tc     [chunk][NTC][32]         per-node track / mesh constants, shared by the whole batch
//   D      [problem][ND]            vehicle parameters followed by their parameter-only sub-expressions ([problem][node][ND] when
//                                   they vary along the lap)
//   V, C   [problem][chunk][NVAL|NCK][32]  structure-of-arrays in chunks of 32 nodes (one warp): slot k of a node sits at a
//                                   fixed byte offset k*256 from the node's base, so every access is base + immediate and a
//                                   warp's access to a slot is one contiguous 256-byte segment
//   g, lambda [problem][element][NCPE]   element-major, the reference's constraint order (optimal_laptime.hpp:413-483)
//   grad   [problem][var node][NV]
//   jac    [problem]{ A[NJA][element] , B[NJB][element with a free left node] }   slot-major: consecutive threads write
//   hess   [problem]{ H[NH][var node] , C[NCPL][coupled element] }               consecutive addresses (coalesced)
// The triplet (row, col) of every slot is fixed at problem creation (fl_nlp_structure); Ipopt's triplet format accepts
// any fixed order, so the order is chosen for the GPU, not for the CPU.
#pragma once
#include <cuda_runtime.h>
#include "fl_math.cuh"
#include <cstdint>
#include <cstring>

struct fl_dev
{
    int n_points, n_elements, n_var_nodes, node0, closed, batch;
    int n_pad;                   // n_points rounded up to a multiple of 32 (chunked workspace layout)
    int nB, nC;                  // number of elements that carry a B block / a control coupling
    int n, m;                    // variables / constraints per problem (user-visible sizes)
    long nnz_jac, nnz_hess;      // per problem
    const double* x;             // [batch][n_points*NV]
    const double* tc;            // [n_pad/32][NTC][32]
    const double* D;             // [batch][ND], or [batch][n_points][ND] when the parameters vary along the lap
    size_t d_prob_stride, d_node_stride;     // ND, 0   or   n_points * ND, ND
    const double* lambda;        // [batch][m]
    const double* zeros;         // >= NCPE zeros
    double obj_factor;
    double* V;                   // [batch][n_pad/32][NVAL][32]
    double* C;                   // [batch][n_pad/32][NCK][32]  checkpoints: transcendental / reciprocal / sqrt intermediates of each node
    double* g;                   // [batch][m]
    double* f;                   // [batch]
    double* f_partial;           // [batch][n_blocks_assemble]
    unsigned int* f_counter;     // [batch]
    double* grad;                // [batch][n]
    double* jac;                 // [batch][nnz_jac]
    double* hess;                // [batch][nnz_hess]
    const int* active;           // [batch] or null: laps with active == 0 are skipped (device-resident solvers)
    int steady_assemble;         // one-node problems: the values pass also writes the rows g and the objective f of its problem
};

// offset of slot 0 of `node` inside a chunked [chunk][nslot][32] array
__host__ __device__ __forceinline__ size_t fl_chunk_base(int node, int nslot) { return ((size_t)(node >> 5) * nslot << 5) + (node & 31); }

#ifndef FL_VBLOCK_N
#define FL_VBLOCK_N 128
#endif
constexpr int FL_BLOCK = FL_VBLOCK_N;        // values pass: nodes per block (x one code partition)
// derivative pass: nodes per block (x one code partition), and the element-assembly block size
#ifndef FL_DBLOCK_N
#define FL_DBLOCK_N 256
#endif
#ifndef FL_DMINB
#define FL_DMINB 1
#endif
#ifndef FL_STAGE
#define FL_STAGE 0
#endif
#ifndef FL_PREFETCH
#define FL_PREFETCH 0
#endif
#ifndef FL_PDL
#define FL_PDL 1
#endif
// FL_L1BYPASS: the derivative kernel's streaming traffic (checkpoints, variables, multipliers, track constants in; gradient, Jacobian and
// Hessian slots out) goes past the L1 (ld.global.cg / st.global.cg), which is then left to the spill frames of the block's warps --
// the loads the warps actually wait for (profiles/prof_r2q_*)
#ifndef FL_L1BYPASS
#define FL_L1BYPASS 0
#endif
constexpr int FL_DBLOCK = FL_DBLOCK_N;
constexpr int FL_ABLOCK = FL_DBLOCK;

// ---------------------------------------------------------------------------------------------------------------------
// Parameter vector of a problem handed over as a kernel argument: when every problem of the batch drives the same vehicle
// (a single lap, replicas, G-G sweeps) the generated code reads D[k] as a constant-bank operand of the arithmetic
// instruction itself -- no load, no register (CD = FL_D_CONST).  Batches with per-problem vehicles read D from global memory at an
// address that is uniform over the block (FL_D_LAP: it stays in uniform registers); parameters that vary along the lap are read per
// node (FL_D_NODE).
enum { FL_D_LAP = 0, FL_D_CONST = 1, FL_D_NODE = 2 };
template <class G> struct DVec { double d[G::ND]; };

template <class G, int CD>
struct ValuesIO
{
    const double* xn; const double* tcn; const double* Dp; const DVec<G>* Dk; double* Vn; double* Cn;
    const double* xp; const double* xq; const double* lin_; const double* lout_; double obj_;     // tape pass only
    __device__ __forceinline__ double lin(int r) const { return lin_[r]; }
    __device__ __forceinline__ double lout(int r) const { return lout_[r]; }
    __device__ __forceinline__ double obj() const { return obj_; }
    __device__ __forceinline__ double uprev(int c) const { return xp[G::ctrl_pos(c)]; }
    __device__ __forceinline__ double unext(int c) const { return xq[G::ctrl_pos(c)]; }
    __device__ __forceinline__ double x(int k) const { return xn[k]; }
    __device__ __forceinline__ double t(int k) const { return __ldg(tcn + k * 32); }
    __device__ __forceinline__ double d(int k) const { if constexpr (CD == FL_D_CONST) return Dk->d[k]; else return __ldg(Dp + k); }
    __device__ __forceinline__ void val(int k, double v) const { Vn[k * 32] = v; }
    __device__ __forceinline__ void cs(int k, double v) const { Cn[k * 32] = v; }
    __device__ __forceinline__ void fence(double) const {}
};

template <class G> __device__ __forceinline__ double assemble_element(const fl_dev& P, const int e, const int prob);

// TAPE: the values pass of a full round (Jacobian and Hessian): it also writes the shared part of the primal / adjoint
// tape, which depends on the multipliers (codegen/generate.py shared_tape), for the derivative partitions to load.
template <class G, int CD, bool TAPE>
__global__ void __launch_bounds__(FL_BLOCK) k_node_values(const __grid_constant__ fl_dev P, const __grid_constant__ DVec<G> Dk)
{
    // programmatic dependent launch: the derivative grid that follows may be scheduled while this one runs; it waits
    // (griddepcontrol.wait) for this grid's completion before it reads the values and checkpoints
    asm volatile("griddepcontrol.launch_dependents;");
    // one-node problems (steady-state NLPs, G::IS_STEADY): a thread is a PROBLEM of the batch -- 32 problems per warp instead of one
    // warp (and block) per problem with a single active lane
    const int node = G::IS_STEADY ? 0 : blockIdx.x * FL_BLOCK + threadIdx.x;
    const int part = blockIdx.y;
    const int prob = G::IS_STEADY ? blockIdx.x * FL_BLOCK + threadIdx.x : blockIdx.z;
    if (G::IS_STEADY ? prob >= P.batch : node >= P.n_points) return;
    if (P.active && !P.active[prob]) return;
    ValuesIO<G, CD> io;
    io.xn = P.x + ((size_t)prob * P.n_points + node) * G::NV;
    io.tcn = P.tc + fl_chunk_base(node, G::NTC);
    io.Dp = CD == FL_D_NODE ? P.D + (size_t)prob * P.d_prob_stride + (size_t)node * P.d_node_stride : P.D + (size_t)prob * G::ND;
    io.Dk = &Dk;
    io.Vn = P.V + (size_t)prob * G::NVAL * P.n_pad + fl_chunk_base(node, G::NVAL);
    io.Cn = P.C + (size_t)prob * G::NCK * P.n_pad + fl_chunk_base(node, G::NCK);
    if constexpr (TAPE)
    {
        const int np = P.n_points;
        const int prev = (node == 0) ? np - 1 : node - 1;
        const int next = (node + 1 == np) ? 0 : node + 1;
        const double* xb = P.x + (size_t)prob * np * G::NV;
        const double* lam = P.lambda + (size_t)prob * P.m;
        io.xp = xb + (size_t)prev * G::NV;
        io.xq = xb + (size_t)next * G::NV;
        io.lin_ = (P.closed || node > 0) ? lam + (size_t)prev * G::NCPE : P.zeros;          // element e starts at node e
        io.lout_ = (P.closed || node + 1 < np) ? lam + (size_t)node * G::NCPE : P.zeros;
        io.obj_ = P.obj_factor;
        G::tape_part(part, io);
    }
    else
        G::values_part(part, io);
    if constexpr (G::IS_STEADY != 0)
    {
        // a thread is a problem and its single element reads this node's values only: the rows and the objective in the same launch
        // (a line-search trial of a G-G batch is then one launch, not two; the thread reads back what it has just written itself)
        if (P.steady_assemble && part == 0) P.f[prob] = assemble_element<G>(P, 0, prob);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Output channels of the model at every node (get_outputs_map of the reference: tyre slips, forces, dissipation, chassis
// accelerations ...): O[problem][channel][node], thread = node.  Not on the solver's path: post-processing of a solution.
template <class G, int CD>
struct OutputsIO
{
    const double* xn; const double* tcn; const double* Dp; const DVec<G>* Dk; double* On; size_t sO;
    __device__ __forceinline__ double x(int k) const { return xn[k]; }
    __device__ __forceinline__ double t(int k) const { return __ldg(tcn + k * 32); }
    __device__ __forceinline__ double d(int k) const { if constexpr (CD == FL_D_CONST) return Dk->d[k]; else return __ldg(Dp + k); }
    __device__ __forceinline__ void out(int k, double v) const { On[k * sO] = v; }
    __device__ __forceinline__ void fence(double) const {}
};

template <class G, int CD>
__global__ void __launch_bounds__(FL_BLOCK) k_node_outputs(const __grid_constant__ fl_dev P, const __grid_constant__ DVec<G> Dk, double* __restrict__ O)
{
    const int node = blockIdx.x * FL_BLOCK + threadIdx.x;
    const int prob = blockIdx.z;
    if (node >= P.n_points) return;
    OutputsIO<G, CD> io;
    io.xn = P.x + ((size_t)prob * P.n_points + node) * G::NV;
    io.tcn = P.tc + fl_chunk_base(node, G::NTC);
    io.Dp = CD == FL_D_NODE ? P.D + (size_t)prob * P.d_prob_stride + (size_t)node * P.d_node_stride : P.D + (size_t)prob * G::ND;
    io.Dk = &Dk;
    io.sO = P.n_points;
    io.On = O + (size_t)prob * G::NOUT * P.n_points + node;
    G::outputs(io);
}

// ---------------------------------------------------------------------------------------------------------------------
// Element assembly.  V offsets: Q [0,NTD) QD [NTD,2NTD) ALG EXT DTDS CDOT.
// Runs as its own kernel (eval_f / eval_g alone) or as one extra blockIdx.y slice of the derivative kernel (a full
// callback round: one launch less, and the rows are written while the derivative partitions compute).
// rows of element e of problem prob; returns the element's term of the objective
template <class G>
__device__ __forceinline__ double assemble_element(const fl_dev& P, const int e, const int prob)
{
    constexpr int oQ = 0, oQD = G::NTD, oALG = 2 * G::NTD, oEXT = oALG + G::NALG, oDT = oEXT + G::NEXT, oCD = oDT + 1;
    const int np = P.n_points;
    double fe = 0.0;
    {
        const int i = e;
        const int j = (e + 1 == np) ? 0 : e + 1;
        const bool closing = (e + 1 == np);
        const double* D = P.D + (size_t)prob * P.d_prob_stride;          // sigma, dissipations: the same at every node
        const double sigma = __ldg(D + G::P_SIGMA), oms = 1.0 - sigma;
        const double h = __ldg(P.tc + fl_chunk_base(i, G::NTC) + (G::NTRACK + 1) * 32);           // hout of the left node
        const double* Vi = P.V + (size_t)prob * G::NVAL * P.n_pad + fl_chunk_base(i, G::NVAL);
        const double* Vj = P.V + (size_t)prob * G::NVAL * P.n_pad + fl_chunk_base(j, G::NVAL);
        const double* xi = P.x + ((size_t)prob * np + i) * G::NV;
        const double* xj = P.x + ((size_t)prob * np + j) * G::NV;
        double* g = P.g + (size_t)prob * P.m + (size_t)e * G::NCPE;
        int row = 0;
#pragma unroll
        for (int r = 0; r < G::NTD - G::NAUG; ++r, ++row)
            g[row] = Vj[(oQ + r) * 32] - Vi[(oQ + r) * 32] - h * (oms * Vi[(oQD + r) * 32] + sigma * Vj[(oQD + r) * 32]);
        if constexpr (G::NAUG > 0)
        {
            // augmented rows (codegen/nlpnode.py): the right node's value carries the free jump, the integrand weights are per-node
            // constants (the last two track slots: awin of the right node, awout of the left one)
            constexpr int oQA = oCD + G::NCR;
            const double awout = __ldg(P.tc + fl_chunk_base(i, G::NTC) + (G::NTRACK - 1) * 32);
            const double awin = __ldg(P.tc + fl_chunk_base(j, G::NTC) + (G::NTRACK - 2) * 32);
#pragma unroll
            for (int r = G::NTD - G::NAUG; r < G::NTD; ++r, ++row)
                g[row] = Vj[(oQA + r - (G::NTD - G::NAUG)) * 32] - Vi[(oQ + r) * 32] - (awout * Vi[(oQD + r) * 32] + awin * Vj[(oQD + r) * 32]);
        }
#pragma unroll
        for (int r = 0; r < G::NALG; ++r, ++row) g[row] = Vj[(oALG + r) * 32];
#pragma unroll
        for (int r = 0; r < G::NEXT; ++r, ++row) g[row] = Vj[(oEXT + r) * 32];
        fe = h * (oms * Vi[oDT * 32] + sigma * Vj[oDT * 32]);
        if (G::IS_DERIVATIVE)
        {
            // control-rate rows (optimal_laptime.hpp:1573-1583, 1632-1641) and penalty (:1553, :1611; the left rate is
            // taken at node 0 for every non-closing element, as the reference's dcontrols_dt[i-i] does)
            const double* xl = P.x + ((size_t)prob * np + (closing ? np - 1 : 0)) * G::NV;
#pragma unroll
            for (int c = 0; c < G::NCR; ++c, ++row)
            {
                g[row] = xj[G::ctrl_pos(c)] - xi[G::ctrl_pos(c)]
                       - h * (oms * Vi[(oCD + c) * 32] + sigma * Vj[(oCD + c) * 32]);
                const double dl = xl[G::dctrl_pos(c)], dr = xj[G::dctrl_pos(c)];
                fe += 0.5 * __ldg(D + G::P_DISS0 + c) * (dl * dl + dr * dr) * h;
            }
        }
        else
        {
#pragma unroll
            for (int c = 0; c < G::NFM; ++c)
            {
                const double derivative = (xj[G::ctrl_pos(c)] - xi[G::ctrl_pos(c)]) / h;     // optimal_laptime.hpp:1362-1366
                fe += __ldg(D + G::P_DISS0 + c) * (derivative * derivative) * h;
            }
        }
    }
    return fe;
}

template <class G, int BS>
__device__ __forceinline__ void assemble_block(const fl_dev& P, const int bx, const int nbx, const int prob)
{
    const int e = bx * BS + threadIdx.x;
    const double fe = e < P.n_elements ? assemble_element<G>(P, e, prob) : 0.0;
    if (P.n_elements == 1)
    {
        // one-element problems (steady-state NLPs): the element's term is the objective, no reduction
        if (threadIdx.x == 0) P.f[prob] = fe;
        return;
    }
    // objective: fixed-order tree inside the block; the last block to finish sums the block partials, again in a fixed
    // order (thread t takes partials t, t + BS, ...; then the same tree) -- the result does not depend on block timing
    __shared__ double red[BS];
    __shared__ bool last;
    red[threadIdx.x] = fe;
    __syncthreads();
#pragma unroll
    for (int s = BS / 2; s > 0; s >>= 1)
    {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    double* part = P.f_partial + (size_t)prob * nbx;
    if (threadIdx.x == 0)
    {
        part[bx] = red[0];
        __threadfence();
        const unsigned int t = atomicInc(P.f_counter + prob, nbx - 1);      // wraps back to 0 for the next launch
        last = (t == (unsigned)nbx - 1);
    }
    __syncthreads();
    if (last)
    {
        __threadfence();
        double acc = 0.0;
        for (int b = threadIdx.x; b < nbx; b += BS) acc += __ldcg(part + b);
        red[threadIdx.x] = acc;
        __syncthreads();
#pragma unroll
        for (int s = BS / 2; s > 0; s >>= 1)
        {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) P.f[prob] = red[0];
    }
}

// (same block size as the slice inside the derivative launch: both paths give the same bits for f)
template <class G>
__global__ void __launch_bounds__(FL_ABLOCK) k_assemble(const __grid_constant__ fl_dev P)
{
    if constexpr (G::IS_STEADY != 0)
    {
        // one-element problems: a thread is a problem, its element's term is the objective (no reduction)
        const int prob = blockIdx.x * FL_ABLOCK + threadIdx.x;
        if (prob >= P.batch || (P.active && !P.active[prob])) return;
        P.f[prob] = assemble_element<G>(P, 0, prob);
        return;
    }
    if (P.active && !P.active[blockIdx.y]) return;
    assemble_block<G, FL_ABLOCK>(P, blockIdx.x, gridDim.x, blockIdx.y);
}

// ---------------------------------------------------------------------------------------------------------------------
template <class G, int CD>
struct DerivIO
{
    const double* xn; const double* xp; const double* xq;     // this node, previous node, next node
    const double* tcn; const double* Dp; const DVec<G>* Dk; const double* lin_; const double* lout_; const double* Cn;
    double obj_;
    double* gradn; double* ja_; double* jb_; double* h_; double* c_;
    size_t sA, sB, sH, sC;
    bool hasB, hasC;
#if FL_L1BYPASS
    static __device__ __forceinline__ double ld(const double* p) { return __ldcg(p); }
    static __device__ __forceinline__ void st(double* p, double v) { __stcg(p, v); }
#else
    static __device__ __forceinline__ double ld(const double* p) { return *p; }
    static __device__ __forceinline__ void st(double* p, double v) { *p = v; }
#endif
    __device__ __forceinline__ double x(int k) const { return ld(xn + k); }
    __device__ __forceinline__ double t(int k) const { return FL_L1BYPASS ? ld(tcn + k * 32) : __ldg(tcn + k * 32); }
    __device__ __forceinline__ double d(int k) const { if constexpr (CD == FL_D_CONST) return Dk->d[k]; else return __ldg(Dp + k); }
    __device__ __forceinline__ double c(int k) const { return ld(Cn + k * 32); }          // global workspace, or the warp's staged copy in shared memory
    __device__ __forceinline__ double lin(int r) const { return ld(lin_ + r); }
    __device__ __forceinline__ double lout(int r) const { return ld(lout_ + r); }
    __device__ __forceinline__ double obj() const { return obj_; }
    __device__ __forceinline__ double uprev(int c) const { return ld(xp + G::ctrl_pos(c)); }
    __device__ __forceinline__ double unext(int c) const { return ld(xq + G::ctrl_pos(c)); }
    // codegen FL_RING: per-node loads as 8-byte cp.async copies into a ring of shared-memory slots (one 256-byte row per slot and
    // warp), issued groups ahead of their use and read back with ld.shared: the L2 round trip of a checkpoint load overlaps the
    // arithmetic in between without a register per load in flight
    unsigned ring_;                                                                    // shared-space address of this thread's slot 0
    __device__ __forceinline__ const double* pc(int k) const { return Cn + k * 32; }
    __device__ __forceinline__ const double* px(int k) const { return xn + k; }
    __device__ __forceinline__ const double* pt(int k) const { return tcn + k * 32; }
    __device__ __forceinline__ const double* plin(int r) const { return lin_ + r; }
    __device__ __forceinline__ const double* plout(int r) const { return lout_ + r; }
    __device__ __forceinline__ const double* puprev(int c) const { return xp + G::ctrl_pos(c); }
    __device__ __forceinline__ const double* punext(int c) const { return xq + G::ctrl_pos(c); }
    __device__ __forceinline__ void rput(int slot, const double* p) const { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(ring_ + slot * 256), "l"(p)); }
    __device__ __forceinline__ void rcommit() const { asm volatile("cp.async.commit_group;"); }
    __device__ __forceinline__ double rget(int slot) const { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ring_ + slot * 256)); return v; }
    __device__ __forceinline__ void val(int, double) const {}
    // codegen FL_SEGMENT > 0 separates code segments with fence(); measured on B200 (profiles/README.md) the fenced,
    // software-prefetched schedule is slower than ptxas' own (the spills are the reverse sweep's tape, not hoisted loads)
    __device__ __forceinline__ void fence(double) const {}
    __device__ __forceinline__ void grad(int k, double v) const { st(gradn + k, v); }
    __device__ __forceinline__ void ja(int k, double v) const { st(ja_ + k * sA, v); }
    __device__ __forceinline__ void jb(int k, double v) const { if (hasB) st(jb_ + k * sB, v); }
    __device__ __forceinline__ void h(int k, double v) const { st(h_ + k * sH, v); }
    __device__ __forceinline__ void cpl(int k, double v) const { if (hasC) st(c_ + k * sC, v); }
};

#define FL_RING_WAIT(N) asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory")

enum { FL_DERIV_JAC = 1, FL_DERIV_HESS = 2, FL_DERIV_ALL = 3 };

template <class G, int WHAT> struct deriv_parts { };
template <class G> struct deriv_parts<G, FL_DERIV_JAC>  { static constexpr int N = G::NP_JAC;  template <class IO> static __device__ __forceinline__ void run(int w, IO& io) { G::jac_part(w, io); } };
template <class G> struct deriv_parts<G, FL_DERIV_HESS> { static constexpr int N = G::NP_HESS; template <class IO> static __device__ __forceinline__ void run(int w, IO& io) { G::hess_part(w, io); } };
template <class G> struct deriv_parts<G, FL_DERIV_ALL>  { static constexpr int N = G::NP_ALL;  template <class IO> static __device__ __forceinline__ void run(int w, IO& io) { G::all_part(w, io); } };

// The derivative outputs of a node are split into N balanced code partitions (straight-line code, ~2 k fp64 operations
// each).  One block = FL_DBLOCK consecutive nodes x ONE partition (blockIdx.y): all warps of a block run the same
// instruction stream, so each instruction-cache line is fetched once per block and shared by its warps -- the code is
// executed exactly once per thread, which makes instruction fetch, not data, the first bottleneck (profiles/).
// A lap of 8000 nodes spreads over 32 x N blocks instead of 63 blocks of one thread per node.

// Staging (FL_STAGE): a warp = one 32-node chunk of the workspace, whose NCK checkpoint rows are contiguous (NCK x 256
// bytes).  Lane 0 brings them into the warp's shared-memory slice with ONE bulk asynchronous copy (cp.async.bulk, the
// 1-D TMA path, completion on an mbarrier); the straight-line code then reads its checkpoints with LDS at immediate
// offsets instead of ~70 dependent global loads per partition (the long-scoreboard stalls of profiles/prof_r1g).
template <class G> constexpr size_t fl_stage_bytes() { return FL_STAGE ? (size_t)(FL_DBLOCK / 32) * (G::NCK * 256 + 8) : 0; }
// ring of the asynchronous loads (codegen FL_RING): RING_SLOTS rows of 256 bytes per warp
template <class G> constexpr size_t fl_ring_bytes(int threads) { return (size_t)(threads / 32) * G::RING_SLOTS * 256; }

template <class G, int WHAT, bool ASSEMBLE, int CD>
__global__ void __launch_bounds__(FL_DBLOCK, FL_DMINB) k_node_derivs(const __grid_constant__ fl_dev P, const __grid_constant__ DVec<G> Dk)
{
    asm volatile("griddepcontrol.wait;" ::: "memory");       // the values pass this launch may overlap with has completed
    // (one-node problems, G::IS_STEADY: a thread is a problem of the batch, blocks of FL_SBLOCK threads)
    const int node = G::IS_STEADY ? 0 : blockIdx.x * FL_DBLOCK + threadIdx.x;   // warps are aligned with the 32-node chunks of the workspace
    const int vn = node - P.node0;                           // variable-node index
    const int part = blockIdx.y;
    const int prob = G::IS_STEADY ? blockIdx.x * blockDim.x + threadIdx.x : blockIdx.z;
    if (G::IS_STEADY && prob >= P.batch) return;
    if (P.active && !P.active[prob]) return;
    if (ASSEMBLE && part == deriv_parts<G, WHAT>::N)         // the slice that assembles g and f (block-uniform branch)
    {
        const int nbx = (P.n_elements + FL_DBLOCK - 1) / FL_DBLOCK;
        if ((int)blockIdx.x < nbx) assemble_block<G, FL_DBLOCK>(P, blockIdx.x, nbx, prob);
        return;
    }
    const int np = P.n_points;
#if FL_STAGE
    extern __shared__ __align__(128) unsigned char fl_stage_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (node - lane >= np) return;                           // the whole warp is beyond the lap
    double* sC = reinterpret_cast<double*>(fl_stage_smem) + (size_t)warp * (G::NCK * 32);
    {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(fl_stage_smem + (size_t)(FL_DBLOCK / 32) * (G::NCK * 256) + warp * 8);
        if (lane == 0)
        {
            const double* src = P.C + (size_t)prob * G::NCK * P.n_pad + fl_chunk_base(node, G::NCK);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sC);
            constexpr unsigned bytes = G::NCK * 256;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
        }
        __syncwarp();
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }
#endif
#if FL_PREFETCH
    {
        // warm L1 with the rows this warp is going to read: its chunk's checkpoints, its nodes' variables and multipliers
        const int lane_ = threadIdx.x & 31;
        const char* cb = reinterpret_cast<const char*>(P.C + (size_t)prob * G::NCK * P.n_pad + fl_chunk_base(node - lane_, G::NCK));
        for (int off = lane_ * 128; off < G::NCK * 256; off += 32 * 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(cb + off));
        const char* xb_ = reinterpret_cast<const char*>(P.x + ((size_t)prob * np + (node - lane_)) * G::NV);
        if (lane_ * 128 < 32 * G::NV * 8 && node - lane_ + 32 <= np) asm volatile("prefetch.global.L1 [%0];" :: "l"(xb_ + lane_ * 128));
    }
#endif
    if (vn < 0 || node >= np) return;
    const int prev = (node == 0) ? np - 1 : node - 1;
    const int next = (node + 1 == np) ? 0 : node + 1;
    const int e_in = prev;                                     // element e starts at node e
    const int e_out = node;
    const bool has_out = P.closed || (node + 1 < np);
    const double* xb = P.x + (size_t)prob * np * G::NV;
    const double* lam = P.lambda + (size_t)prob * P.m;
    DerivIO<G, CD> io;
    io.Dk = &Dk;
    io.xn = xb + (size_t)node * G::NV;
    io.xp = xb + (size_t)prev * G::NV;
    io.xq = xb + (size_t)next * G::NV;
    io.tcn = P.tc + fl_chunk_base(node, G::NTC);
    io.Dp = CD == FL_D_NODE ? P.D + (size_t)prob * P.d_prob_stride + (size_t)node * P.d_node_stride : P.D + (size_t)prob * G::ND;
    io.lin_ = lam + (size_t)e_in * G::NCPE;
    io.lout_ = has_out ? lam + (size_t)e_out * G::NCPE : P.zeros;
    io.obj_ = P.obj_factor;
    if constexpr (G::RING_SLOTS > 0)
    {
        static_assert(!FL_STAGE, "the staged checkpoints and the load ring both use the dynamic shared memory");
        extern __shared__ __align__(128) unsigned char fl_ring_smem[];
        io.ring_ = (unsigned)__cvta_generic_to_shared(fl_ring_smem) + (threadIdx.x >> 5) * (G::RING_SLOTS * 256) + (threadIdx.x & 31) * 8;
    }
#if FL_STAGE
    io.Cn = sC + lane;
#else
    io.Cn = P.C + (size_t)prob * G::NCK * P.n_pad + fl_chunk_base(node, G::NCK);
#endif
    io.gradn = P.grad + (size_t)prob * P.n + (size_t)vn * G::NV;
    double* jac = P.jac + (size_t)prob * P.nnz_jac;
    double* hes = P.hess + (size_t)prob * P.nnz_hess;
    io.sA = P.n_elements; io.sB = P.nB; io.sH = P.n_var_nodes; io.sC = P.nC;
    io.ja_ = jac + e_in;
    // B blocks / couplings exist for elements whose left node is free: all of them on closed laps, e >= 1 on open ones
    io.hasB = has_out && (e_out >= P.node0);
    io.jb_ = jac + (size_t)G::NJA * P.n_elements + (e_out - P.node0);
    io.h_ = hes + vn;
    io.hasC = (e_in >= P.node0);
    io.c_ = hes + (size_t)G::NH * P.n_var_nodes + (e_in - P.node0);
    deriv_parts<G, WHAT>::run(part, io);
}

// ---------------------------------------------------------------------------------------------------------------------
// What the C-ABI layer sees of a variant (one translation unit per variant fills this in).
struct fl_variant
{
    const char* name;
    int NV, NCPE, NTD, NALG, NEXT, NCR, NFM, NVAL, NCK, NTC, NTRACK, NRAW, ND, NJA, NJB, NH, NCPL, IS_DERIVATIVE, IS_3D;
    int IS_STEADY, NP_VALUES;
    int NAUG, AUG_POS, P_AUG;      // augmented state (codegen/nlpnode.py): count, node variable of `a`, first aug_* parameter slot
    int NOUT; const char* const* OUT_NAMES;                     // named output channels
    void (*launch_outputs)(const fl_dev&, const double* D0, int shared_d, double* O, cudaStream_t);
    int OPS_VALUES, OPS_JAC, OPS_HESS, OPS_ALL;
    const int *CTRL_POS, *DCTRL_POS, *JA_ROW, *JA_COL, *JB_ROW, *JB_COL, *H_ROW, *H_COL;
    int P_SIGMA, P_DISS0, P_DISS1, P_F_WHEEL_SCALE, P_R_WHEEL_SCALE;     // -1 when the model has no such slot
    void (*derive)(double* D);
    // `D0` = host copy of the parameter vector of problem 0; `shared_d` = every problem of the batch has that same vector
    int HAS_TAPE;      // a full round runs the tape pass (launch_values with tape = 1) instead of the values pass
    void (*launch_values)(const fl_dev&, const double* D0, int shared_d, int tape, cudaStream_t);
    void (*launch_assemble)(const fl_dev&, cudaStream_t);
    void (*launch_derivs)(const fl_dev&, const double* D0, int shared_d, int what, int with_assemble, cudaStream_t);
    int (*assemble_blocks)(const fl_dev&);
};

template <class G> struct has_wheel_scale { template <class T> static auto test(int) -> decltype(T::P_F_WHEEL_SCALE, char()); template <class T> static long test(...);
                                            static constexpr bool value = sizeof(test<G>(0)) == 1; };
template <class G, bool = has_wheel_scale<G>::value> struct wheel_slots { static constexpr int f = -1, r = -1; };
template <class G> struct wheel_slots<G, true> { static constexpr int f = G::P_F_WHEEL_SCALE, r = G::P_R_WHEEL_SCALE; };

template <class G>
struct variant_ops
{
    static dim3 grid(int items, int batch) { return dim3((items + FL_BLOCK - 1) / FL_BLOCK, batch, 1); }
    static void values(const fl_dev& P, const double* D0, int shared_d, int tape, cudaStream_t s)
    {
        DVec<G> Dk;
        memcpy(Dk.d, D0, sizeof(Dk.d));
        const dim3 grid(G::IS_STEADY ? (P.batch + FL_BLOCK - 1) / FL_BLOCK : (P.n_points + FL_BLOCK - 1) / FL_BLOCK, tape ? G::NP_TAPE : G::NP_VALUES,
                        G::IS_STEADY ? 1 : P.batch);
        if constexpr (G::HAS_TAPE != 0)          // (the tape kernels exist only in builds generated with FL_SHARE > 0)
        {
            if (tape)
            {
                if (shared_d) k_node_values<G, FL_D_CONST, true><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
                else if (P.d_node_stride) k_node_values<G, FL_D_NODE, true><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
                else k_node_values<G, FL_D_LAP, true><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
                return;
            }
        }
        if (shared_d) k_node_values<G, FL_D_CONST, false><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
        else if (P.d_node_stride) k_node_values<G, FL_D_NODE, false><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
        else k_node_values<G, FL_D_LAP, false><<<grid, FL_BLOCK, 0, s>>>(P, Dk);
    }
    static void outputs(const fl_dev& P, const double* D0, int shared_d, double* O, cudaStream_t s)
    {
        if constexpr (G::NOUT > 0)
        {
            DVec<G> Dk;
            memcpy(Dk.d, D0, sizeof(Dk.d));
            const dim3 grid((P.n_points + FL_BLOCK - 1) / FL_BLOCK, 1, P.batch);
            if (shared_d) k_node_outputs<G, FL_D_CONST><<<grid, FL_BLOCK, 0, s>>>(P, Dk, O);
            else if (P.d_node_stride) k_node_outputs<G, FL_D_NODE><<<grid, FL_BLOCK, 0, s>>>(P, Dk, O);
            else k_node_outputs<G, FL_D_LAP><<<grid, FL_BLOCK, 0, s>>>(P, Dk, O);
        }
    }
    static void assemble(const fl_dev& P, cudaStream_t s)
    {
        if constexpr (G::IS_STEADY != 0) k_assemble<G><<<dim3((P.batch + FL_ABLOCK - 1) / FL_ABLOCK, 1, 1), FL_ABLOCK, 0, s>>>(P);
        else k_assemble<G><<<dim3((P.n_elements + FL_ABLOCK - 1) / FL_ABLOCK, P.batch, 1), FL_ABLOCK, 0, s>>>(P);
    }
    static int assemble_blocks(const fl_dev& P) { return (P.n_elements + FL_ABLOCK - 1) / FL_ABLOCK; }
    template <int CD>
    static void derivs_cd(const fl_dev& P, const DVec<G>& Dk, int what, int with_assemble, cudaStream_t s)
    {
        // n_elements <= n_var_nodes + 1: one more block column covers the assembly slice of an open problem
        const unsigned nb = (P.n_points + FL_DBLOCK - 1) / FL_DBLOCK;
        const size_t sm = fl_stage_bytes<G>() + fl_ring_bytes<G>(FL_DBLOCK);
        static unsigned long long attr_set = 0;          // per device (and per instantiation of this function)
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((attr_set >> dev) & 1ull) && sm > 48 * 1024)
        {
            cudaFuncSetAttribute(k_node_derivs<G, FL_DERIV_JAC, false, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(k_node_derivs<G, FL_DERIV_HESS, false, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(k_node_derivs<G, FL_DERIV_ALL, false, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(k_node_derivs<G, FL_DERIV_ALL, true, CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            attr_set |= 1ull << dev;
        }
        // launched with programmatic stream serialisation: the grid may be scheduled before the values pass ahead of it in
        // the stream has drained (its blocks wait at griddepcontrol.wait), which hides the launch and ramp-up latency
        cudaLaunchConfig_t cfg = {};
        // one-node problems (steady-state NLPs): one warp per block instead of eight -- a block of 256 threads at 255 registers fills
        // an SM on its own, and a batch of 2016 problems x 8 partitions then runs in 110 waves of blocks with ONE active thread each
        // (208 us per launch in the G-G sweep, profiles/r2_gg_launches.md); the assembly slice needs the full block and is launched on its own
        const bool tiny = !FL_STAGE && P.n_points <= 32 && !with_assemble;
        constexpr int FL_SBLOCK = 64;               // one-node problems: problems per block (a thread each)
        cfg.blockDim = dim3(G::IS_STEADY ? FL_SBLOCK : (tiny ? 32 : FL_DBLOCK), 1, 1); cfg.dynamicSmemBytes = G::IS_STEADY ? fl_ring_bytes<G>(FL_SBLOCK) : (tiny ? fl_ring_bytes<G>(32) : sm); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = (FL_PDL && P.batch <= 64) ? 1 : 0;          // measured: 44.4 -> 42.2 us per 8000-node round; no gain once the batch fills the GPU
        if constexpr (G::IS_STEADY != 0)
        {
            const unsigned nbs = (P.batch + FL_SBLOCK - 1) / FL_SBLOCK;
            if (with_assemble || P.n_points != 1) return;          // (fl_eval_device never folds the assembly into one-node launches)
            if (what == FL_DERIV_JAC) { cfg.gridDim = dim3(nbs, G::NP_JAC, 1); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_JAC, false, CD>, P, Dk); }
            else if (what == FL_DERIV_HESS) { cfg.gridDim = dim3(nbs, G::NP_HESS, 1); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_HESS, false, CD>, P, Dk); }
            else { cfg.gridDim = dim3(nbs, G::NP_ALL, 1); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_ALL, false, CD>, P, Dk); }
            return;
        }
        if (what == FL_DERIV_JAC) { cfg.gridDim = dim3(nb, G::NP_JAC, P.batch); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_JAC, false, CD>, P, Dk); }
        else if (what == FL_DERIV_HESS) { cfg.gridDim = dim3(nb, G::NP_HESS, P.batch); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_HESS, false, CD>, P, Dk); }
        else if (!with_assemble) { cfg.gridDim = dim3(nb, G::NP_ALL, P.batch); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_ALL, false, CD>, P, Dk); }
        else { cfg.gridDim = dim3(nb, G::NP_ALL + 1, P.batch); cudaLaunchKernelEx(&cfg, k_node_derivs<G, FL_DERIV_ALL, true, CD>, P, Dk); }
    }
    static void derivs(const fl_dev& P, const double* D0, int shared_d, int what, int with_assemble, cudaStream_t s)
    {
        DVec<G> Dk;
        memcpy(Dk.d, D0, sizeof(Dk.d));
        if (shared_d) derivs_cd<FL_D_CONST>(P, Dk, what, with_assemble, s);
        else if (P.d_node_stride) derivs_cd<FL_D_NODE>(P, Dk, what, with_assemble, s);
        else derivs_cd<FL_D_LAP>(P, Dk, what, with_assemble, s);
    }
    static fl_variant make()
    {
        fl_variant v;
        v.name = G::name;
        v.NV = G::NV; v.NCPE = G::NCPE; v.NTD = G::NTD; v.NALG = G::NALG; v.NEXT = G::NEXT; v.NCR = G::NCR; v.NFM = G::NFM;
        v.HAS_TAPE = G::HAS_TAPE;
        v.NVAL = G::NVAL; v.NCK = G::NCK; v.NTC = G::NTC; v.NTRACK = G::NTRACK; v.NRAW = G::NRAW; v.ND = G::ND;
        v.NJA = G::NJA; v.NJB = G::NJB; v.NH = G::NH; v.NCPL = G::NCPL; v.IS_DERIVATIVE = G::IS_DERIVATIVE; v.IS_3D = G::IS_3D;
        v.NAUG = G::NAUG; v.AUG_POS = G::AUG_POS; v.P_AUG = G::P_AUG; v.IS_STEADY = G::IS_STEADY; v.NP_VALUES = G::NP_VALUES;
        v.NOUT = G::NOUT; v.OUT_NAMES = G::OUT_NAMES; v.launch_outputs = &outputs;
        v.OPS_VALUES = G::OPS_VALUES; v.OPS_JAC = G::OPS_JAC; v.OPS_HESS = G::OPS_HESS; v.OPS_ALL = G::OPS_ALL;
        v.CTRL_POS = G::CTRL_POS; v.DCTRL_POS = G::DCTRL_POS;
        v.JA_ROW = G::JA_ROW; v.JA_COL = G::JA_COL; v.JB_ROW = G::JB_ROW; v.JB_COL = G::JB_COL; v.H_ROW = G::H_ROW; v.H_COL = G::H_COL;
        v.P_SIGMA = G::P_SIGMA; v.P_DISS0 = G::P_DISS0; v.P_DISS1 = G::P_DISS1;
        v.P_F_WHEEL_SCALE = wheel_slots<G>::f; v.P_R_WHEEL_SCALE = wheel_slots<G>::r;
        v.derive = &G::derive;
        v.launch_values = &values; v.launch_assemble = &assemble; v.launch_derivs = &derivs; v.assemble_blocks = &assemble_blocks;
        return v;
    }
};
