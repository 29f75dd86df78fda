// C-ABI of the collocation-NLP callbacks (include/fastestlap_b200.h).  Host side: problem set-up (per-node track and mesh
// constants, parameter vectors, fixed sparsity, bounds) and kernel launches; no model arithmetic happens on the host.
// Reference counterparts: Optimal_laptime::compute_direct / compute_derivative (variable and constraint vectors, bounds:
// src/core/applications/optimal_laptime.hpp:326-483, 780-960) and lion::ipopt_cppad_solve's TNLP callbacks (:546-551).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "fl_internal.h"

#define FL_DECLARE_VARIANT(name) extern "C" const fl_variant* fl_variant_##name();
FL_DECLARE_VARIANT(f1_direct)
FL_DECLARE_VARIANT(f1_derivative)
FL_DECLARE_VARIANT(f1_direct_3d)
FL_DECLARE_VARIANT(f1_derivative_3d)
FL_DECLARE_VARIANT(kart_direct)
FL_DECLARE_VARIANT(kart_derivative)
FL_DECLARE_VARIANT(kart_direct_torque)
FL_DECLARE_VARIANT(kart_derivative_torque)
FL_DECLARE_VARIANT(kart_steady)
FL_DECLARE_VARIANT(f1_steady)
FL_DECLARE_VARIANT(f1_direct_aug)

namespace {

thread_local std::string g_error;

bool fail(const std::string& msg) { g_error = msg; return false; }

#define FL_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    g_error = std::string(#call) + ": " + cudaGetErrorString(e_); return 0; } } while (0)

constexpr double DEG = 3.14159265358979323846 / 180.0;
constexpr double KMH = 1.0 / 3.6;
constexpr double G0 = 9.81;
constexpr int F1_NPARAMS = 58, KART_NPARAMS = 72;
constexpr int F1_NIN = 14, KART_NIN = 13, F1_NCTRL = 4, KART_NCTRL = 2;

const fl_variant* pick_variant(int model, int form, int elevation, int torque, int aug = 0)
{
    if (aug) return (model == FL_MODEL_F1_3DOF && form == FL_FORM_DIRECT && !elevation) ? fl_variant_f1_direct_aug() : nullptr;
    if (model == FL_MODEL_F1_3DOF)
    {
        if (form == FL_FORM_STEADY) return elevation ? nullptr : fl_variant_f1_steady();
        if (form == FL_FORM_DIRECT) return elevation ? fl_variant_f1_direct_3d() : fl_variant_f1_direct();
        return elevation ? fl_variant_f1_derivative_3d() : fl_variant_f1_derivative();
    }
    if (model == FL_MODEL_KART_6DOF && !elevation)
    {
        if (form == FL_FORM_STEADY) return fl_variant_kart_steady();
        if (form == FL_FORM_DIRECT) return torque ? fl_variant_kart_direct_torque() : fl_variant_kart_direct();
        return torque ? fl_variant_kart_derivative_torque() : fl_variant_kart_derivative();
    }
    return nullptr;
}

template <class T> T* dev_alloc(size_t count)
{
    void* p = nullptr;
    if (fl_dev_malloc(&p, (count ? count : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, (count ? count : 1) * sizeof(T));
    return static_cast<T*>(p);
}

} // namespace


namespace {

// x of the user ([batch][n], free nodes only) -> xfull ([batch][n_points*NV]); a strided device copy when the first node is fixed
int scatter_x(fl_nlp* h)
{
    const fl_dev& P = h->P;
    const int NV = h->v->NV;
    if (P.node0 == 0) return 1;     // xfull aliases the user vector
    FL_CUDA(cudaMemcpy2DAsync(h->d_xfull + (size_t)P.node0 * NV, (size_t)P.n_points * NV * sizeof(double), h->d_xuser, (size_t)P.n * sizeof(double),
                              (size_t)P.n * sizeof(double), P.batch, cudaMemcpyDeviceToDevice, h->stream));
    return 1;
}

void build_structure(fl_nlp* h)
{
    const fl_variant* v = h->v;
    const fl_dev& P = h->P;
    const int np = P.n_points, NV = v->NV, NCPE = v->NCPE;
    h->jac_row.resize(P.nnz_jac); h->jac_col.resize(P.nnz_jac);
    h->hess_row.resize(P.nnz_hess); h->hess_col.resize(P.nnz_hess);
    size_t k = 0;
    for (int a = 0; a < v->NJA; ++a)
        for (int e = 0; e < P.n_elements; ++e, ++k)
        {
            const int j = (e + 1 == np) ? 0 : e + 1;
            h->jac_row[k] = e * NCPE + v->JA_ROW[a];
            h->jac_col[k] = (j - P.node0) * NV + v->JA_COL[a];
        }
    for (int b = 0; b < v->NJB; ++b)
        for (int eb = 0; eb < P.nB; ++eb, ++k)
        {
            const int e = eb + P.node0;
            h->jac_row[k] = e * NCPE + v->JB_ROW[b];
            h->jac_col[k] = (e - P.node0) * NV + v->JB_COL[b];
        }
    k = 0;
    for (int a = 0; a < v->NH; ++a)
        for (int vn = 0; vn < P.n_var_nodes; ++vn, ++k)
        {
            h->hess_row[k] = vn * NV + v->H_ROW[a];
            h->hess_col[k] = vn * NV + v->H_COL[a];
        }
    for (int c = 0; c < v->NCPL; ++c)
        for (int ec = 0; ec < P.nC; ++ec, ++k)
        {
            const int e = ec + P.node0;                       // element e joins node e -> (e+1) % np, both free
            const int j = (e + 1 == np) ? 0 : e + 1;
            const int r = (j - P.node0) * NV + v->CTRL_POS[c], q = (e - P.node0) * NV + v->CTRL_POS[c];
            h->hess_row[k] = r > q ? r : q;
            h->hess_col[k] = r > q ? q : r;
        }
}

} // namespace

void fl_set_error(const std::string& msg) { g_error = msg; }

// ---- cache of freed device blocks (see fl_internal.h) ----------------------------------------------------------------------
namespace {
struct DevCache
{
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;      // (device, bytes) -> block
    std::map<void*, std::pair<int, size_t>> live;                    // block -> (device, bytes)
    size_t cached = 0;
    size_t limit = [] { const char* e = std::getenv("FL_CACHE_BYTES"); return e ? (size_t)std::strtoull(e, nullptr, 10) : (size_t)64 << 30; }();
    // (64 GB: the handles of a 4096-lap sweep are 45 GB, of which three blocks of 7 - 13 GB; cudaMalloc of that much costs 0.2 - 1.8 s on a
    //  B200, so a caller that creates handles of the same sizes again and again -- sweeps of a service, bench.py after its warm-up -- gets
    //  them back from here; a failing cudaMalloc empties the cache and tries again)
    void release_all()
    {
        int cur = 0; cudaGetDevice(&cur);
        for (auto& kv : free_blocks) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
        free_blocks.clear(); cached = 0;
        cudaSetDevice(cur);
    }
};
DevCache& dev_cache() { static DevCache c; return c; }
}

cudaError_t fl_dev_malloc(void** p, size_t bytes)
{
    DevCache& c = dev_cache();
    int dev = 0;
    cudaGetDevice(&dev);
    bytes = (bytes + 255) / 256 * 256;
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.free_blocks.find({ dev, bytes });
    if (it != c.free_blocks.end())
    {
        *p = it->second; c.free_blocks.erase(it); c.cached -= bytes;
        c.live[*p] = { dev, bytes };
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && c.cached > 0) { cudaGetLastError(); c.release_all(); e = cudaMalloc(p, bytes); }
    if (e == cudaSuccess) c.live[*p] = { dev, bytes };
    return e;
}

void fl_dev_free(void* p)
{
    if (!p) return;
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { cudaFree(p); return; }
    const auto key = it->second;
    c.live.erase(it);
    // (a full cache refuses further blocks: a caller that moves on to a different workload -- bench.py between its legs -- empties
    // it with fl_release_cached_memory, otherwise every handle of the new workload goes through cudaMalloc / cudaFree again)
    if (key.second > c.limit / 2 || c.cached + key.second > c.limit)
    {
        int cur = 0; cudaGetDevice(&cur);
        cudaSetDevice(key.first); cudaFree(p); cudaSetDevice(cur);
        return;
    }
    c.free_blocks.insert({ key, p }); c.cached += key.second;
}

extern "C" void fl_release_cached_memory(void)
{
    DevCache& c = dev_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    c.release_all();
}

extern "C" {

const char* fl_last_error(void) { return g_error.c_str(); }

int fl_n_vehicle_params(int model) { return model == FL_MODEL_F1_3DOF ? F1_NPARAMS : (model == FL_MODEL_KART_6DOF ? KART_NPARAMS : -1); }

void fl_nlp_destroy(fl_nlp* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    {
        const std::vector<fl_ipm*> deps = h->dependents;
        for (fl_ipm* s : deps) fl_ipm_orphan(s);
    }
    double* bufs[] = { h->d_xfull, h->P.node0 ? h->d_xuser : nullptr, h->d_tc, h->d_D, h->d_lambda, h->d_zeros, h->d_V, h->d_C, h->d_g, h->d_f, h->d_fpart,
                       h->d_grad, h->d_jac, h->d_hess };
    for (double* b : bufs) if (b) fl_dev_free(b);
    if (h->d_fcount) fl_dev_free(h->d_fcount);
    if (h->d_flush) fl_dev_free(h->d_flush);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static fl_nlp* create_impl(const fl_problem_desc* d, const double* steady_spec);

fl_nlp* fl_nlp_create(const fl_problem_desc* d)
{
    if (d && d->form == FL_FORM_STEADY) { fail("steady-state problems are created with fl_steady_create"); return nullptr; }
    return create_impl(d, nullptr);
}

// Steady-state (G-G diagram) NLPs: a batch of one-node problems (see fl_steady_desc in the header).
fl_nlp* fl_steady_create(const fl_steady_desc* sd)
{
    if (!sd || !sd->vehicle_params || !sd->spec) { fail("null steady-state description"); return nullptr; }
    if (sd->model != FL_MODEL_KART_6DOF && sd->model != FL_MODEL_F1_3DOF) { fail("unknown model"); return nullptr; }
    const double s0 = 0.0, zeros6[6] = {0, 0, 0, 0, 0, 0}, w0 = 0.0;
    fl_problem_desc d{};
    d.model = sd->model; d.form = FL_FORM_STEADY; d.closed = 1; d.elevation = 0; d.kart_direct_torque = sd->model == FL_MODEL_KART_6DOF;
    d.n_points = 1; d.s = &s0; d.track_length = 1.0; d.track_nodes = zeros6; d.wl = &w0; d.wr = &w0;
    d.batch = sd->batch; d.vehicle_params = sd->vehicle_params; d.sigma = 0.5; d.dissipation[0] = d.dissipation[1] = 0.0;
    d.device = sd->device;
    return create_impl(&d, sd->spec);
}

static fl_nlp* create_impl(const fl_problem_desc* d, const double* steady_spec)
{
    if (!d) { fail("null problem description"); return nullptr; }
    const fl_variant* v = pick_variant(d->model, d->form, d->elevation, d->kart_direct_torque, d->aug_kind);
    if (!v && d->aug_kind) { fail("hypermesh / constant controls and integral constraints are offered for the F1, direct form, on flat tracks"); return nullptr; }
    if (!v) { fail(d->model == FL_MODEL_KART_6DOF && d->elevation ? "lot2016kart does not support tracks with elevation" : "unknown model / form"); return nullptr; }
    if (d->aug_kind)
    {
        if (!d->closed) { fail("hypermesh / constant controls and integral constraints need a closed lap"); return nullptr; }
        if (d->aug_kind != FL_AUG_BRAKE_BIAS && d->aug_kind != FL_AUG_INTEGRAL) { fail("unknown aug_kind"); return nullptr; }
        if (d->aug_kind == FL_AUG_BRAKE_BIAS && !d->aug_jump) { fail("aug_jump is null"); return nullptr; }
        if (!(d->aug_bounds[0] < d->aug_bounds[1])) { fail("aug_bounds must be an interval"); return nullptr; }
    }
    if (d->n_points < 2 && !steady_spec) { fail("provide at least two values of arclength"); return nullptr; }
    if (d->batch < 1) { fail("batch must be >= 1"); return nullptr; }
    if (!d->s || !d->track_nodes || !d->wl || !d->wr || (!d->vehicle_params && !d->node_params)) { fail("null array in problem description"); return nullptr; }
    if (d->node_params && steady_spec) { fail("steady-state problems have no arclength to vary the parameters with"); return nullptr; }
    for (int i = 1; i < d->n_points; ++i)
        if (!(d->s[i] > d->s[i - 1])) { fail("arclength must be strictly increasing"); return nullptr; }
    if (d->closed && !(d->track_length > d->s[d->n_points - 1])) { fail("closed meshes must end before the track length"); return nullptr; }
    if (!d->closed && (!d->q0 || !d->u0 || (d->form == FL_FORM_DERIVATIVE && !d->du0))) { fail("open problems need the inputs and controls of the first node"); return nullptr; }
    if (d->ctrl_default && d->model == FL_MODEL_F1_3DOF && !steady_spec)
    {
        // The controls that are not optimised do not travel through this field: the node programs take the boost as 0 and the brake
        // bias from the vehicle parameters (chassis/brake_bias, slot 18).  A caller that asks for anything else is told so.
        const double* p0 = d->node_params ? d->node_params : d->vehicle_params;
        if (d->ctrl_default[1] != 0.0) { fail("ctrl_default: a boost other than 0 is not evaluated by the node programs"); return nullptr; }
        if (std::fabs(d->ctrl_default[3] - p0[18]) > 1.0e-14)
        { fail("ctrl_default: the brake bias of a lap is vehicle parameter 18 (chassis/brake_bias); set it there (it may differ per lap of a batch)"); return nullptr; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { fail("no CUDA device: this library has no CPU path"); return nullptr; }
    if (d->device < 0 || d->device >= ndev) { fail("invalid CUDA device ordinal"); return nullptr; }
    if (cudaSetDevice(d->device) != cudaSuccess) { fail("cudaSetDevice failed"); return nullptr; }

    fl_nlp* h = new fl_nlp();
    h->v = v; h->desc = *d; h->device = d->device;
    h->n_in = d->model == FL_MODEL_F1_3DOF ? F1_NIN : KART_NIN;
    h->n_ctrl = d->model == FL_MODEL_F1_3DOF ? F1_NCTRL : KART_NCTRL;
    const int np = d->n_points, NV = v->NV, batch = d->batch;
    const int nvp = fl_n_vehicle_params(d->model);
    h->s.assign(d->s, d->s + np); h->wl.assign(d->wl, d->wl + np); h->wr.assign(d->wr, d->wr + np);
    if (d->node_params)
    {
        // parameters that vary along the lap: one vector per node; the per-problem vector (bounds, queries) is node 0's
        h->node_params.assign(d->node_params, d->node_params + (size_t)batch * np * nvp);
        h->params.resize((size_t)batch * nvp);
        for (int b = 0; b < batch; ++b) std::memcpy(h->params.data() + (size_t)b * nvp, h->node_params.data() + (size_t)b * np * nvp, nvp * sizeof(double));
    }
    else
        h->params.assign(d->vehicle_params, d->vehicle_params + (size_t)batch * nvp);
    h->desc.s = h->s.data(); h->desc.wl = h->wl.data(); h->desc.wr = h->wr.data(); h->desc.vehicle_params = h->params.data();
    h->desc.node_params = d->node_params ? h->node_params.data() : nullptr;
    h->desc.track_nodes = nullptr; h->desc.q0 = h->desc.u0 = h->desc.du0 = h->desc.ctrl_default = nullptr;
    if (d->aug_kind == FL_AUG_BRAKE_BIAS) { h->aug_jump.assign(d->aug_jump, d->aug_jump + np); h->desc.aug_jump = h->aug_jump.data(); }
    else h->desc.aug_jump = nullptr;

    fl_dev& P = h->P;
    P.n_points = np; P.closed = d->closed ? 1 : 0; P.batch = batch;
    P.node0 = d->closed ? 0 : 1;
    P.n_pad = (np + 31) / 32 * 32;
    P.n_elements = d->closed ? np : np - 1;
    P.n_var_nodes = np - P.node0;
    P.nB = P.n_elements - P.node0;
    P.nC = v->NCPL ? P.n_elements - P.node0 : 0;
    P.n = P.n_var_nodes * NV;
    P.m = P.n_elements * v->NCPE;
    P.nnz_jac = (long)v->NJA * P.n_elements + (long)v->NJB * P.nB;
    P.nnz_hess = (long)v->NH * P.n_var_nodes + (long)v->NCPL * P.nC;
    P.obj_factor = 1.0;

    // ---- per-node constants [NTC][np]: track (kz | kx ky kz sin/cos pitch, sin/cos roll), then hin, hout, (ihin, ihout | penw)
    std::vector<double> tc((size_t)v->NTC * P.n_pad, 0.0);
    auto TC = [&](int slot, int node) -> double& { return tc[fl_chunk_base(node, v->NTC) + (size_t)slot * 32]; };
    const double L = d->track_length;
    for (int i = 0; i < np; ++i)
    {
        const double* t = d->track_nodes + 6 * i;
        const double mu = t[1], phi = t[2], dth = t[3], dmu = t[4], dph = t[5];
        // Road_curvilinear::curvature (src/core/vehicles/road_curvilinear.h:49-60)
        const double kx = dph - std::sin(mu) * dth;
        const double ky = std::cos(phi) * dmu + std::cos(mu) * std::sin(phi) * dth;
        const double kz = -std::sin(phi) * dmu + std::cos(mu) * std::cos(phi) * dth;
        // wind (eastward, -northward, 0) along the track's tangent, normal and binormal (road_curvilinear.hpp:121-131,
        // chassis.hpp:269-274): the rotation by the car's own yaw alpha is left to the node program
        const double th = t[0], wE = steady_spec ? 0.0 : d->wind[1], wS = steady_spec ? 0.0 : -d->wind[0];
        const double cth = std::cos(th), sth = std::sin(th), cmu_ = std::cos(mu), smu_ = std::sin(mu), cph_ = std::cos(phi), sph_ = std::sin(phi);
        const double wt = cth * cmu_ * wE + sth * cmu_ * wS;
        const double wn = (cth * smu_ * sph_ - sth * cph_) * wE + (sth * smu_ * sph_ + cth * cph_) * wS;
        const double wb = (cth * smu_ * cph_ + sth * sph_) * wE + (sth * smu_ * cph_ - cth * sph_) * wS;
        int k = 0;
        if (v->IS_3D)
        {
            TC(k++, i) = kx; TC(k++, i) = ky; TC(k++, i) = kz;
            TC(k++, i) = smu_; TC(k++, i) = cmu_;
            TC(k++, i) = sph_; TC(k++, i) = cph_;
            TC(k++, i) = wt; TC(k++, i) = wn; TC(k++, i) = wb;
        }
        else if (v->NTRACK >= 3) { TC(k++, i) = kz; TC(k++, i) = wt; TC(k++, i) = wn; }
        else
            TC(k++, i) = kz;
        const double hin = i > 0 ? d->s[i] - d->s[i - 1] : (d->closed ? L - d->s[np - 1] : 0.0);
        const double hout = i + 1 < np ? d->s[i + 1] - d->s[i] : (d->closed ? L - d->s[np - 1] : 0.0);
        if (v->NAUG)
        {
            // rj, bout, awin, awout (codegen/nlpnode.py).  The closing element of the reference weights its integrands the other
            // way round: (1 - sigma) at node 0, sigma at the last node (optimal_laptime.hpp:1409)
            const bool integral = d->aug_kind == FL_AUG_INTEGRAL;
            TC(k++, i) = integral ? 0.0 : (d->aug_jump[i] != 0.0 ? 1.0 : 0.0);
            TC(k++, i) = (integral && i == 0) ? 0.0 : 1.0;
            TC(k++, i) = hin * (i == 0 ? 1.0 - d->sigma : d->sigma);
            TC(k++, i) = hout * (i + 1 == np ? d->sigma : 1.0 - d->sigma);
        }
        TC(k++, i) = hin;
        TC(k++, i) = hout;
        if (!v->IS_DERIVATIVE)
        {
            TC(k++, i) = hin > 0.0 ? 1.0 / hin : 0.0;
            TC(k++, i) = hout > 0.0 ? 1.0 / hout : 0.0;
        }
        else
            ++k;    // penw, below
    }
    if (v->IS_DERIVATIVE)
    {
        // total element length weighting 0.5 diss du_j^2: right node of its element, plus the reference's left-rate quirk
        // (dcontrols_dt[i-i], optimal_laptime.hpp:1553): node 0 for every non-closing element, node n-1 for the closing one
        const int penw = v->NTC - 1;
        for (int e = 0; e < P.n_elements; ++e)
        {
            const bool closing = e + 1 == np;
            const int j = closing ? 0 : e + 1;
            const double he = closing ? L - d->s[np - 1] : d->s[e + 1] - d->s[e];
            TC(penw, j) += he;
            TC(penw, closing ? np - 1 : 0) += he;
        }
    }

    // ---- parameter vectors [batch][ND] (or [batch][n_points][ND]): vehicle parameters, host-side slots, parameter-only sub-expressions
    const int nvec = d->node_params ? batch * np : batch;
    const double* raw = d->node_params ? h->node_params.data() : h->params.data();
    std::vector<double> D((size_t)nvec * v->ND, 0.0);
    for (int b = 0; b < nvec; ++b)
    {
        double* Db = D.data() + (size_t)b * v->ND;
        std::memcpy(Db, raw + (size_t)b * nvp, nvp * sizeof(double));
        if (v->P_F_WHEEL_SCALE >= 0)
        {
            // axle_car_3dof.hpp:294: wheel equations are scaled by 1/m when the wheels have no inertia
            const double m = Db[0], If = Db[22], Ir = Db[26];
            Db[v->P_F_WHEEL_SCALE] = If < 1.0e-10 ? 1.0 / m : 1.0;
            Db[v->P_R_WHEEL_SCALE] = Ir < 1.0e-10 ? 1.0 / m : 1.0;
        }
        Db[v->P_SIGMA] = d->sigma;
        Db[v->P_DISS0] = d->dissipation[0];
        Db[v->P_DISS1] = d->dissipation[1];
        if (steady_spec) std::memcpy(Db + v->P_DISS1 + 1, steady_spec + (size_t)b * 7, 7 * sizeof(double));    // ss_v ... ss_cay follow diss1
        if (v->NAUG)
        {
            Db[v->P_AUG] = d->aug_kind == FL_AUG_BRAKE_BIAS ? 1.0 : 0.0;
            for (int k = 0; k < 5; ++k) Db[v->P_AUG + 1 + k] = d->aug_kind == FL_AUG_INTEGRAL ? d->aug_weights[k] : 0.0;
        }
        v->derive(Db);
    }
    P.d_prob_stride = d->node_params ? (size_t)np * v->ND : (size_t)v->ND;
    P.d_node_stride = d->node_params ? (size_t)v->ND : 0;

    h->D0.assign(D.begin(), D.begin() + v->ND);
    for (int b = 1; b < nvec && h->shared_d; ++b)
        if (std::memcmp(D.data() + (size_t)b * v->ND, D.data(), v->ND * sizeof(double)) != 0) h->shared_d = 0;

    // ---- x with the fixed first node of open problems
    std::vector<double> xfull((size_t)batch * np * NV, 0.0);
    if (!d->closed)
    {
        const int time_idx = h->n_in - 3;
        std::vector<double> x0(NV, 0.0);
        int k = 0;
        for (int j = 0; j < h->n_in; ++j) if (j != time_idx) x0[k++] = d->q0[j];
        for (int c = 0; c < v->NFM; ++c) x0[v->CTRL_POS[c]] = d->u0[d->model == FL_MODEL_F1_3DOF ? 2 * c : c];
        if (v->IS_DERIVATIVE) for (int c = 0; c < v->NFM; ++c) x0[v->DCTRL_POS[c]] = d->du0[c];
        for (int b = 0; b < batch; ++b) std::memcpy(xfull.data() + (size_t)b * np * NV, x0.data(), NV * sizeof(double));
    }

    const int nblk = v->assemble_blocks(P);
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess
           && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    h->d_xfull = dev_alloc<double>((size_t)batch * np * NV);
    h->d_xuser = P.node0 ? dev_alloc<double>((size_t)batch * P.n) : h->d_xfull;
    h->d_tc = dev_alloc<double>(tc.size());
    h->d_D = dev_alloc<double>(D.size());
    h->d_lambda = dev_alloc<double>((size_t)batch * P.m);
    h->d_zeros = dev_alloc<double>(64);
    h->d_V = dev_alloc<double>((size_t)batch * v->NVAL * P.n_pad);
    h->d_C = dev_alloc<double>((size_t)batch * v->NCK * P.n_pad);
    h->d_g = dev_alloc<double>((size_t)batch * P.m);
    h->d_f = dev_alloc<double>(batch);
    h->d_fpart = dev_alloc<double>((size_t)batch * nblk);
    h->d_fcount = dev_alloc<unsigned int>(batch);
    h->d_grad = dev_alloc<double>((size_t)batch * P.n);
    h->d_jac = dev_alloc<double>((size_t)batch * P.nnz_jac);
    h->d_hess = dev_alloc<double>((size_t)batch * P.nnz_hess);
    ok = ok && h->d_xfull && h->d_xuser && h->d_tc && h->d_D && h->d_lambda && h->d_zeros && h->d_V && h->d_C && h->d_g && h->d_f && h->d_fpart
            && h->d_fcount && h->d_grad && h->d_jac && h->d_hess;
    ok = ok && cudaMemcpy(h->d_tc, tc.data(), tc.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess
            && cudaMemcpy(h->d_D, D.data(), D.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess
            && cudaMemcpy(h->d_xfull, xfull.data(), xfull.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok)
    {
        const std::string msg = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        fl_nlp_destroy(h);
        fail(msg);
        return nullptr;
    }
    P.x = h->d_xfull; P.tc = h->d_tc; P.D = h->d_D; P.lambda = h->d_lambda; P.zeros = h->d_zeros;
    P.V = h->d_V; P.C = h->d_C; P.g = h->d_g; P.f = h->d_f; P.f_partial = h->d_fpart; P.f_counter = h->d_fcount;
    P.grad = h->d_grad; P.jac = h->d_jac; P.hess = h->d_hess;
    build_structure(h);
    return h;
}

void fl_nlp_sizes(const fl_nlp* h, int* n, int* m, int* nnz_jac, int* nnz_hess)
{
    if (n) *n = h->P.n;
    if (m) *m = h->P.m;
    if (nnz_jac) *nnz_jac = (int)h->P.nnz_jac;
    if (nnz_hess) *nnz_hess = (int)h->P.nnz_hess;
}

int fl_nlp_batch(const fl_nlp* h) { return h->P.batch; }
const char* fl_nlp_variant(const fl_nlp* h) { return h->v->name; }
void fl_nlp_ops_per_node(const fl_nlp* h, int ops[4]) { ops[0] = h->v->OPS_VALUES; ops[1] = h->v->OPS_JAC; ops[2] = h->v->OPS_HESS; ops[3] = h->v->OPS_ALL; }
long fl_kernel_launches(const fl_nlp* h) { return h->launches; }
int fl_nlp_n_outputs(const fl_nlp* h) { return h->v->NOUT; }

const char* fl_nlp_output_name(const fl_nlp* h, int k) { return (k >= 0 && k < h->v->NOUT) ? h->v->OUT_NAMES[k] : nullptr; }

int fl_eval_outputs(fl_nlp* h, const double* x, double* out)
{
    FL_CUDA(cudaSetDevice(h->device));
    const fl_dev& P = h->P;
    if (!x || !out) return fail("null argument") ? 1 : 0;
    if (h->v->NOUT == 0) return fail("this problem has no output channels") ? 1 : 0;
    const size_t count = (size_t)P.batch * h->v->NOUT * P.n_points;
    double* d_out = nullptr;
    FL_CUDA(fl_dev_malloc((void**)&d_out, count * sizeof(double)));
    int ok = 1;
    if (cudaMemcpyAsync(h->d_xuser, x, (size_t)P.batch * P.n * sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess || !scatter_x(h)) ok = 0;
    if (ok)
    {
        h->values_valid = h->derivs_jac_valid = false;
        h->v->launch_outputs(P, h->D0.data(), h->shared_d, d_out, h->stream);
        h->launches += 1;
        if (cudaMemcpyAsync(out, d_out, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) ok = 0;
    }
    if (!ok) g_error = cudaGetErrorString(cudaGetLastError());
    fl_dev_free(d_out);
    return ok;
}

void fl_nlp_workspace_slots(const fl_nlp* h, int* n_values, int* n_checkpoints)
{
    if (n_values) *n_values = h->v->NVAL;
    if (n_checkpoints) *n_checkpoints = h->v->NCK;
}

int fl_nlp_structure(const fl_nlp* h, int* jac_row, int* jac_col, int* hess_row, int* hess_col)
{
    if (jac_row) std::memcpy(jac_row, h->jac_row.data(), h->jac_row.size() * sizeof(int));
    if (jac_col) std::memcpy(jac_col, h->jac_col.data(), h->jac_col.size() * sizeof(int));
    if (hess_row) std::memcpy(hess_row, h->hess_row.data(), h->hess_row.size() * sizeof(int));
    if (hess_col) std::memcpy(hess_col, h->hess_col.data(), h->hess_col.size() * sizeof(int));
    return 1;
}

int fl_problem_bounds(const fl_problem_desc* d, double* x_l, double* x_u, double* g_l, double* g_u)
{
    if (!d || !d->wl || !d->wr || (!d->vehicle_params && !d->node_params)) return fail("null problem description") ? 1 : 0;
    const fl_variant* v = pick_variant(d->model, d->form, d->elevation, d->kart_direct_torque, d->aug_kind);
    if (!v) return fail("unknown model / form") ? 1 : 0;
    const int node0 = d->closed ? 0 : 1, n_points = d->n_points;
    const int n_var_nodes = n_points - node0, n_elements = d->closed ? n_points : n_points - 1;
    const int NV = v->NV, n_in = d->model == FL_MODEL_F1_3DOF ? F1_NIN : KART_NIN;
    const double big = std::numeric_limits<double>::max();
    std::vector<double> lb(NV), ub(NV);
    const double* par = d->vehicle_params ? d->vehicle_params : d->node_params;        // bounds that depend on parameters use problem 0
    if (d->form == FL_FORM_STEADY && d->model == FL_MODEL_F1_3DOF)
    {
        // limebeer2014f1.h:57-86: steady_state_variable_bounds, then ax, ay in g (the reference widens them until they are
        // inactive, steady_state.hpp:200-226); constraints: 11 equations, lambda / lambda_max in [-1, 1.2]
        const double l[13] = { -1.15, -1.15, -1.15, -1.15, 0.0, -2.0, -2.0, -2.0, -2.0, 0.0, -1.0, -8.0, -8.0 };
        const double u[13] = { 0.25, 0.25, 1.0, 1.0, 1.5, 0.1, 0.1, 0.1, 0.1, 1.5, 1.0, 8.0, 8.0 };
        for (int k = 0; k < 13; ++k) { if (x_l) x_l[k] = l[k]; if (x_u) x_u[k] = u[k]; }
        for (int r = 0; r < 15; ++r) { if (g_l) g_l[r] = r < 11 ? 0.0 : -1.0; if (g_u) g_u[r] = r < 11 ? 0.0 : 1.2; }
        return 1;
    }
    if (d->form == FL_FORM_STEADY)
    {
        // lot2016kart.h:81-121: steady_state_variable_bounds, then ax, ay (solve_max_lat_acc uses +-2 / +-20, steady_state.hpp:146-147;
        // the G-G loop narrows ax per speed, callers may pass their own); constraints: 6 equations, kappa +-0.11, lambda +-0.09
        const double l[8] = { -0.5, 1.0e-4, -10.0 * DEG, -10.0 * DEG, -10.0 * DEG, -10.0 * DEG, -20.0, -20.0 };
        const double u[8] = { 0.5, 0.04, 10.0 * DEG, 10.0 * DEG, 10.0 * DEG, 10.0 * DEG, 20.0, 20.0 };
        for (int k = 0; k < 8; ++k) { if (x_l) x_l[k] = l[k]; if (x_u) x_u[k] = u[k]; }
        for (int r = 0; r < 12; ++r)
        {
            const double b = r < 6 ? 0.0 : (r < 8 ? 0.11 : 0.09);
            if (g_l) g_l[r] = -b;
            if (g_u) g_u[r] = b;
        }
        return 1;
    }
    if (d->model == FL_MODEL_F1_3DOF)
    {
        // axle_car_3dof.hpp:366-397, chassis.hpp:400-425 (50..380 km/h, |v| <= 50 km/h, |omega| <= 10), chassis_car_3dof.hpp:306-352,
        // road_curvilinear.hpp:63-82
        const double l[13] = { -1, -1, -1, -1, 50.0 * KMH, -50.0 * KMH, -10.0, -3.0, -3.0, -3.0, -3.0, 0.0, -45.0 * DEG };
        const double u[13] = { 1, 1, 1, 1, 380.0 * KMH, 50.0 * KMH, 10.0, -0.01, -0.01, -0.01, -0.01, 0.0, 45.0 * DEG };
        for (int k = 0; k < 13; ++k) { lb[k] = l[k]; ub[k] = u[k]; }
        lb[v->CTRL_POS[0]] = -20.0 * DEG; ub[v->CTRL_POS[0]] = 20.0 * DEG;      // steering
        lb[v->CTRL_POS[1]] = -1.0; ub[v->CTRL_POS[1]] = 1.0;                    // throttle
        if (v->IS_DERIVATIVE) { lb[15] = -20.0 * DEG; ub[15] = 20.0 * DEG; lb[16] = -10.0; ub[16] = 10.0; }   // limebeer2014f1.h:227-230, by control id
        if (v->NAUG)
        {
            // a: the control's own bounds / a wide box around the accumulated integral (node 0 holds the whole integral: its bounds,
            // below); q: a jump as large as the control's range / inert
            const double lo = d->aug_bounds[0], up = d->aug_bounds[1];
            const double wide = 10.0 * std::fmax(1.0, std::fmax(std::fabs(lo), std::fabs(up)));
            const bool integral = d->aug_kind == FL_AUG_INTEGRAL;
            lb[v->AUG_POS] = integral ? -wide : lo; ub[v->AUG_POS] = integral ? wide : up;
            lb[v->AUG_POS + 1] = integral ? -1.0 : -(up - lo); ub[v->AUG_POS + 1] = integral ? 1.0 : up - lo;
        }
    }
    else
    {
        // axle_car_6dof.hpp:311-335, chassis.hpp:400-425 (10..200 km/h), chassis_car_6dof.hpp:320-365, road_curvilinear.hpp:63-82
        const double l[12] = { 10.0 * KMH / 0.139, 10.0 * KMH, -10.0 * KMH, -10.0, 1.0e-5, -30.0 * DEG, -30.0 * DEG, -10.0, -10.0, -10.0, 0.0, -45.0 * DEG };
        const double u[12] = { 200.0 * KMH / 0.139, 200.0 * KMH, 10.0 * KMH, 10.0, 0.139, 30.0 * DEG, 30.0 * DEG, 10.0, 10.0, 10.0, 0.0, 45.0 * DEG };
        for (int k = 0; k < 12; ++k) { lb[k] = l[k]; ub[k] = u[k]; }
        lb[12] = -20.0 * DEG; ub[12] = 20.0 * DEG;
        lb[13] = -1.0; ub[13] = 1.0;
        if (v->IS_DERIVATIVE)
        {
            const double r = d->kart_direct_torque ? 4000.0 : 10.0;      // lot2016kart.h:186-192
            lb[14] = -20.0 * DEG; ub[14] = 20.0 * DEG; lb[15] = -r; ub[15] = r;
        }
    }
    const int n_pos = n_in - 3;      // position of the lateral displacement within the node variables
    for (int vn = 0; vn < n_var_nodes; ++vn)
    {
        const int node = vn + node0;
        for (int k = 0; k < NV; ++k)
        {
            if (x_l) x_l[(size_t)vn * NV + k] = (k == n_pos) ? -d->wl[node] : lb[k];
            if (x_u) x_u[(size_t)vn * NV + k] = (k == n_pos) ? d->wr[node] : ub[k];
        }
        if (v->NAUG && d->aug_kind == FL_AUG_INTEGRAL && node == 0)
        {
            if (x_l) x_l[(size_t)vn * NV + v->AUG_POS] = d->aug_bounds[0];
            if (x_u) x_u[(size_t)vn * NV + v->AUG_POS] = d->aug_bounds[1];
        }
    }
    (void)big;
    // constraints: equalities, then the 6 extra rows at the element's right node, then (derivative form) the control-rate rows
    const int n_eq = v->NTD + v->NALG;
    double el[6], eu[6];
    for (int e = 0; e < n_elements; ++e)
    {
        const int j = (e + 1 == n_points) ? 0 : e + 1;
        if (d->model == FL_MODEL_F1_3DOF)
        {
            // limebeer2014f1.h:232-262: +-max(lambda_max(0), lambda_max(m g0)) per tyre, track limits for n +- t/2 cos(alpha)
            const double* pj = d->node_params ? d->node_params + (size_t)j * fl_n_vehicle_params(d->model) : par;      // (problem 0)
            const double m = pj[0];
            for (int t = 0; t < 4; ++t)
            {
                const double* tp = pj + (t < 2 ? 32 : 45);
                const double l1 = tp[9] * DEG, l2 = tp[10] * DEG, Fz1 = tp[1], Fz2 = tp[2];
                const double a = (0.0 - Fz1) * (l2 - l1) / (Fz2 - Fz1) + l1, b = (m * G0 - Fz1) * (l2 - l1) / (Fz2 - Fz1) + l1;
                eu[t] = a > b ? a : b; el[t] = -eu[t];
            }
            el[4] = el[5] = -d->wl[j]; eu[4] = eu[5] = d->wr[j];
        }
        else
            for (int t = 0; t < 6; ++t) { el[t] = -0.11; eu[t] = 0.11; }      // lot2016kart.h:194-198
        for (int r = 0; r < v->NCPE; ++r)
        {
            const bool ext = r >= n_eq && r < n_eq + 6;
            if (g_l) g_l[(size_t)e * v->NCPE + r] = ext ? el[r - n_eq] : 0.0;
            if (g_u) g_u[(size_t)e * v->NCPE + r] = ext ? eu[r - n_eq] : 0.0;
        }
    }
    return 1;
}

int fl_nlp_bounds(const fl_nlp* h, double* x_l, double* x_u, double* g_l, double* g_u)
{
    return fl_problem_bounds(&h->desc, x_l, x_u, g_l, g_u);
}

// ---------------------------------------------------------------------------------------------------------------------
double* fl_dev_x(fl_nlp* h) { return h->d_xuser; }
double* fl_dev_lambda(fl_nlp* h) { return h->d_lambda; }
const double* fl_dev_f(fl_nlp* h) { return h->d_f; }
const double* fl_dev_g(fl_nlp* h) { return h->d_g; }
const double* fl_dev_grad(fl_nlp* h) { return h->d_grad; }
const double* fl_dev_jac(fl_nlp* h) { return h->d_jac; }
const double* fl_dev_hess(fl_nlp* h) { return h->d_hess; }
void* fl_nlp_stream(fl_nlp* h) { return (void*)h->stream; }

int fl_nlp_sync(fl_nlp* h)
{
    FL_CUDA(cudaSetDevice(h->device));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

int fl_eval_device(fl_nlp* h, int what, double obj_factor) { return fl_eval_device_masked(h, what, obj_factor, nullptr); }

} // extern "C"

// internal: the same for the laps whose mask entry is non-zero (device pointer, [batch]); used by the interior-point driver
int fl_eval_device_masked(fl_nlp* h, int what, double obj_factor, const int* d_active)
{
    FL_CUDA(cudaSetDevice(h->device));
    if (!scatter_x(h)) return 0;
    fl_dev P = h->P;
    P.obj_factor = obj_factor;
    P.active = d_active;
    // the values pass also writes the checkpoints every derivative pass loads: it always runs (x may have been rewritten
    // on the device since the last call)
    const int dw = ((what & FL_EVAL_JAC) ? FL_DERIV_JAC : 0) | ((what & FL_EVAL_HESS) ? FL_DERIV_HESS : 0);
    // one-node problems: the values launch writes g and f of its problems as well (nlp_kernels.cuh, k_node_values)
    P.steady_assemble = (h->v->IS_STEADY && h->v->NP_VALUES == 1 && (what & FL_EVAL_FG) && !(what & FL_EVAL_NO_VALUES)) ? 1 : 0;
    if (!(what & FL_EVAL_NO_VALUES))
    {
        h->v->launch_values(P, h->D0.data(), h->shared_d, dw == FL_DERIV_ALL, h->stream);
        h->launches += 1;
    }
    // a full round (f, g, grad f, Jacobian, Hessian) assembles the rows inside the derivative launch
    const int fold = (what & FL_EVAL_FG) && dw == FL_DERIV_ALL && P.n_points > 32;          // (one-node problems: see derivs_cd)
    if ((what & FL_EVAL_FG) && !fold && !P.steady_assemble)
    {
        h->v->launch_assemble(P, h->stream);
        h->launches += 1;
    }
    if (dw) { h->v->launch_derivs(P, h->D0.data(), h->shared_d, dw, fold, h->stream); h->launches += 1; }
    FL_CUDA(cudaGetLastError());
    return 1;
}

extern "C" {

double fl_time_device(fl_nlp* h, int what, double obj_factor, int reps)
{
    if (cudaSetDevice(h->device) != cudaSuccess || reps < 1) return -1.0;
    if (cudaEventRecord(h->ev0, h->stream) != cudaSuccess) return -1.0;
    for (int r = 0; r < reps; ++r) if (!fl_eval_device(h, what, obj_factor)) return -1.0;
    if (cudaEventRecord(h->ev1, h->stream) != cudaSuccess) return -1.0;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) { g_error = cudaGetErrorString(cudaGetLastError()); return -1.0; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    return (double)ms / reps;
}

// `steps` evaluations issued back to back on ONE stream, step k on instance k mod count (a ring of independent instances of
// the same problem defeats the L2 between steps), bracketed by one pair of events: device time of the sequence in ms.
// The host only has to keep the queue filled, so the figure does not contain its launch latency.
double fl_time_device_ring(fl_nlp** hs, int count, int what, double obj_factor, int steps)
{
    if (!hs || count < 1 || steps < 1) return -1.0;
    fl_nlp* h0 = hs[0];
    if (cudaSetDevice(h0->device) != cudaSuccess) return -1.0;
    for (int k = 0; k < count; ++k)
    {
        if (hs[k]->device != h0->device) { g_error = "fl_time_device_ring: instances on different devices"; return -1.0; }
        if (cudaStreamSynchronize(hs[k]->stream) != cudaSuccess) return -1.0;
    }
    std::vector<cudaStream_t> own(count);
    for (int k = 0; k < count; ++k) { own[k] = hs[k]->stream; hs[k]->stream = h0->stream; }
    bool ok = cudaEventRecord(h0->ev0, h0->stream) == cudaSuccess;
    for (int k = 0; ok && k < steps; ++k) ok = fl_eval_device(hs[k % count], what, obj_factor) != 0;
    ok = ok && cudaEventRecord(h0->ev1, h0->stream) == cudaSuccess && cudaEventSynchronize(h0->ev1) == cudaSuccess;
    for (int k = 0; k < count; ++k) hs[k]->stream = own[k];
    if (!ok) { g_error = cudaGetErrorString(cudaGetLastError()); return -1.0; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h0->ev0, h0->ev1);
    return (double)ms;
}

int fl_flush_l2(fl_nlp* h)
{
    FL_CUDA(cudaSetDevice(h->device));
    if (!h->d_flush)
    {
        int l2 = 0;
        FL_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
        h->flush_bytes = 2 * (size_t)(l2 > 0 ? l2 : (128 << 20));
        FL_CUDA(fl_dev_malloc((void**)&h->d_flush, h->flush_bytes));
    }
    FL_CUDA(cudaMemsetAsync(h->d_flush, 0, h->flush_bytes, h->stream));
    return 1;
}

} // extern "C"

__global__ void k_math_eval(int fn, const double* x, double* y, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r = 0.0;
    switch (fn)
    {
    case 0: r = fl_rcp(v); break;
    case 1: r = fl_sqrt(v); break;
    case 2: r = fl_sin(v); break;
    case 3: r = fl_cos(v); break;
    case 4: r = fl_atan(v); break;
    case 5: r = fl_tanh(v); break;
    case 6: r = fl_exp(v); break;
    default: break;
    }
    y[i] = r;
}

extern "C" int fl_math_eval(int device, int fn, const double* x, double* y, long n)
{
    if (n <= 0 || !x || !y || fn < 0 || fn > 6) return fail("fl_math_eval: bad arguments") ? 1 : 0;
    FL_CUDA(cudaSetDevice(device));
    double *dx = nullptr, *dy = nullptr;
    FL_CUDA(cudaMalloc(&dx, (size_t)n * sizeof(double)));
    if (cudaMalloc(&dy, (size_t)n * sizeof(double)) != cudaSuccess) { cudaFree(dx); return fail("fl_math_eval: out of device memory") ? 1 : 0; }
    bool ok = cudaMemcpy(dx, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) { k_math_eval<<<(unsigned)((n + 255) / 256), 256>>>(fn, dx, dy, n); ok = cudaGetLastError() == cudaSuccess; }
    ok = ok && cudaMemcpy(y, dy, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(dx); cudaFree(dy);
    if (!ok) return fail(cudaGetErrorString(cudaGetLastError())) ? 1 : 0;
    return 1;
}

// FP64 pipe peak: every thread runs 8 independent chains of dependent DFMAs; 148 x 8 blocks of 256 threads
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i)
    {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" {

// Measured FP64 (non-tensor) peak of the handle's device in DFMA/s (the roofline denominator SURVEY.md 8d asks for).
double fl_measure_fp64_peak(fl_nlp* h)
{
    if (cudaSetDevice(h->device) != cudaSuccess) return -1.0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = sms * 8, iters = 1 << 14;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)) != cudaSuccess) return -1.0;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep)
    {
        cudaEventRecord(h->ev0, h->stream);
        k_dfma_peak<<<blocks, 256, 0, h->stream>>>(out, iters, 0.999999, 1.0e-6);
        cudaEventRecord(h->ev1, h->stream);
        if (cudaEventSynchronize(h->ev1) != cudaSuccess) { cudaFree(out); return -1.0; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        const double rate = (double)blocks * 256 * 8 * iters / (ms * 1.0e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaFree(out);
    return best;
}

void* fl_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_error = cudaGetErrorString(cudaGetLastError()); return nullptr; }
    return p;
}

void fl_host_free(void* p) { if (p) cudaFreeHost(p); }

int fl_eval_all(fl_nlp* h, const double* x, const double* lambda, double obj_factor,
                double* f, double* g, double* grad_f, double* jac_values, double* hess_values)
{
    FL_CUDA(cudaSetDevice(h->device));
    const fl_dev& P = h->P;
    const size_t B = P.batch;
    if (!x) return fail("x is null") ? 1 : 0;
    FL_CUDA(cudaMemcpyAsync(h->d_xuser, x, B * P.n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (hess_values)
    {
        if (!lambda) return fail("lambda is null") ? 1 : 0;
        FL_CUDA(cudaMemcpyAsync(h->d_lambda, lambda, B * P.m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    const int what = ((f || g) ? FL_EVAL_FG : 0) | ((grad_f || jac_values) ? FL_EVAL_JAC : 0) | (hess_values ? FL_EVAL_HESS : 0);
    if (!fl_eval_device(h, what, obj_factor)) return 0;
    if (f) FL_CUDA(cudaMemcpyAsync(f, h->d_f, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (g) FL_CUDA(cudaMemcpyAsync(g, h->d_g, B * P.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (grad_f) FL_CUDA(cudaMemcpyAsync(grad_f, h->d_grad, B * P.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (jac_values) FL_CUDA(cudaMemcpyAsync(jac_values, h->d_jac, B * P.nnz_jac * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (hess_values) FL_CUDA(cudaMemcpyAsync(hess_values, h->d_hess, B * P.nnz_hess * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    h->values_valid = (what & FL_EVAL_FG) != 0;
    h->derivs_jac_valid = (what & FL_EVAL_JAC) != 0;
    return 1;
}

// ---- Ipopt C-interface callbacks: problem 0 of the batch ----------------------------------------------------------------
static int upload_x(fl_nlp* h, int n, const double* x, int new_x)
{
    if (n != h->P.n) return fail("eval: wrong number of variables") ? 1 : 0;
    if (new_x)
    {
        FL_CUDA(cudaSetDevice(h->device));
        FL_CUDA(cudaMemcpyAsync(h->d_xuser, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->values_valid = h->derivs_jac_valid = false;
    }
    return 1;
}

static int ensure_values(fl_nlp* h)
{
    if (h->values_valid) return 1;
    if (!fl_eval_device(h, FL_EVAL_FG, 1.0)) return 0;
    h->values_valid = true;
    return 1;
}

bool fl_eval_f(int n, double* x, bool new_x, double* obj_value, void* ud)
{
    fl_nlp* h = (fl_nlp*)ud;
    if (!upload_x(h, n, x, new_x) || !ensure_values(h)) return 0;
    FL_CUDA(cudaMemcpyAsync(obj_value, h->d_f, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

bool fl_eval_g(int n, double* x, bool new_x, int m, double* g, void* ud)
{
    fl_nlp* h = (fl_nlp*)ud;
    if (m != h->P.m) return fail("eval_g: wrong number of constraints") ? 1 : 0;
    if (!upload_x(h, n, x, new_x) || !ensure_values(h)) return 0;
    FL_CUDA(cudaMemcpyAsync(g, h->d_g, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

static int ensure_jac(fl_nlp* h)
{
    if (h->derivs_jac_valid) return 1;
    if (!fl_eval_device(h, FL_EVAL_JAC, 1.0)) return 0;
    h->derivs_jac_valid = true;
    return 1;
}

bool fl_eval_grad_f(int n, double* x, bool new_x, double* grad_f, void* ud)
{
    fl_nlp* h = (fl_nlp*)ud;
    if (!upload_x(h, n, x, new_x) || !ensure_jac(h)) return 0;
    FL_CUDA(cudaMemcpyAsync(grad_f, h->d_grad, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

bool fl_eval_jac_g(int n, double* x, bool new_x, int m, int nele_jac, int* iRow, int* jCol, double* values, void* ud)
{
    fl_nlp* h = (fl_nlp*)ud;
    if (nele_jac != (int)h->P.nnz_jac || m != h->P.m) return fail("eval_jac_g: wrong sizes") ? 1 : 0;
    if (!values) return fl_nlp_structure(h, iRow, jCol, nullptr, nullptr);
    if (!upload_x(h, n, x, new_x) || !ensure_jac(h)) return 0;
    FL_CUDA(cudaMemcpyAsync(values, h->d_jac, (size_t)nele_jac * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

bool fl_eval_h(int n, double* x, bool new_x, double obj_factor, int m, double* lambda, bool new_lambda,
              int nele_hess, int* iRow, int* jCol, double* values, void* ud)
{
    fl_nlp* h = (fl_nlp*)ud;
    (void)new_lambda;
    if (nele_hess != (int)h->P.nnz_hess || m != h->P.m) return fail("eval_h: wrong sizes") ? 1 : 0;
    if (!values) return fl_nlp_structure(h, nullptr, nullptr, iRow, jCol);
    if (!upload_x(h, n, x, new_x)) return 0;
    FL_CUDA(cudaMemcpyAsync(h->d_lambda, lambda, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (!fl_eval_device(h, FL_EVAL_HESS, obj_factor)) return 0;
    FL_CUDA(cudaMemcpyAsync(values, h->d_hess, (size_t)nele_hess * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));
    return 1;
}

} // extern "C"
