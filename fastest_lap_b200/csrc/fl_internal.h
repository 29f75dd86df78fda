// Internals shared by the translation units of libfastestlap_b200.so (not part of the C ABI).
#pragma once
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/fastestlap_b200.h"
#include "nlp_kernels.cuh"

void fl_set_error(const std::string& msg);

// Device memory of the handles goes through a small per-device cache of freed blocks (exact size match): a G-G diagram creates and
// destroys a dozen handles of the same few sizes, and cudaMalloc / cudaFree are synchronising calls whose cost on a shared box varied
// from 1 ms to half a second per handle (profiles/README.md).  Blocks return to the driver when the cache holds more than
// FL_CACHE_BYTES (default 64 GB) or when an allocation fails (cache released, allocation retried).  fl_release_cached_memory() empties it.
cudaError_t fl_dev_malloc(void** p, size_t bytes);
void fl_dev_free(void* p);
extern "C" void fl_release_cached_memory(void);
struct fl_nlp;
struct fl_ipm;
// a solver outlives its problem handle only as an empty shell: fl_nlp_destroy releases the device memory of the solvers
// still bound to the handle and detaches them (a garbage collector may finalise the two in either order)
void fl_ipm_orphan(fl_ipm* s);
int fl_eval_device_masked(fl_nlp* h, int what, double obj_factor, const int* d_active);

struct fl_nlp
{
    const fl_variant* v = nullptr;
    fl_problem_desc desc{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    fl_dev P{};
    // owned device buffers
    double *d_xfull = nullptr, *d_xuser = nullptr, *d_tc = nullptr, *d_D = nullptr, *d_lambda = nullptr, *d_zeros = nullptr;
    double *d_C = nullptr, *d_V = nullptr, *d_g = nullptr, *d_f = nullptr, *d_fpart = nullptr, *d_grad = nullptr, *d_jac = nullptr, *d_hess = nullptr;
    unsigned int* d_fcount = nullptr;
    std::vector<double> s, wl, wr, params, node_params, aug_jump, D0;      // D0: parameter vector (with derived slots) of problem 0
    int shared_d = 1;                               // every problem of the batch has the parameter vector D0
    std::vector<int> jac_row, jac_col, hess_row, hess_col;
    bool values_valid = false, derivs_jac_valid = false;
    long launches = 0;
    int n_in = 0, n_ctrl = 0;
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    std::vector<fl_ipm*> dependents;                // interior-point solvers bound to this handle
};
