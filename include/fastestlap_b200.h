/* fastestlap_b200.h -- C ABI of the B200-native collocation-NLP callbacks of fastest-lap.
 *
 * Drop-in boundary.  In the reference the collocation NLP reaches Ipopt through a C++ functor
 *     FG::operator()(ADvector& fg, const ADvector& x)        src/core/applications/optimal_laptime.h:281-287, 329-335
 * handed to lion::ipopt_cppad_solve                           src/core/applications/optimal_laptime.hpp:546-551
 * which implements Ipopt::TNLP (eval_f, eval_grad_f, eval_g, eval_jac_g, eval_h) by sweeping a CppAD tape.
 * This library replaces tape + sweeps: the functions below have the signatures of Ipopt's C interface
 * (IpStdCInterface.h: Eval_F_CB, Eval_Grad_F_CB, Eval_G_CB, Eval_Jac_G_CB, Eval_H_CB; `user_data` = fl_nlp*), so the
 * handle can be given to CreateIpoptProblem as it is, or driven by any other interior-point loop.
 *
 * Conventions (Ipopt's): indices are 0-based (C style); eval_jac_g / eval_h called with values == NULL return the
 * sparsity structure in iRow/jCol; the Hessian is the lower triangle of obj_factor * d2f + sum_i lambda_i d2g_i;
 * duplicated (row, col) pairs are to be summed.  Callbacks return 1 on success and 0 on failure (fl_last_error()).
 * No C++ exception crosses this boundary.  Host pointers everywhere unless a function says "device".
 * A handle is not re-entrant (as the reference's functor, optimal_laptime.h:222-243); different handles are independent.
 * There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef FASTESTLAP_B200_H
#define FASTESTLAP_B200_H

#include <stddef.h>

#ifndef __cplusplus
#include <stdbool.h>
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fl_nlp fl_nlp;

enum { FL_MODEL_F1_3DOF = 0, FL_MODEL_KART_6DOF = 1 };   /* limebeer2014f1 (vehicles/limebeer2014f1.h), lot2016kart (vehicles/lot2016kart.h) */
enum { FL_FORM_DIRECT = 0, FL_FORM_DERIVATIVE = 1,      /* Optimal_laptime::compute_direct / compute_derivative */
       FL_FORM_STEADY = 2 };                             /* steady-state NLPs (fl_steady_create) */

/* Number of entries of the vehicle parameter vector of a model (order: fastest_lap_b200/vehicle.py F1_PARAM_PATHS /
 * KART_PARAM_PATHS, the XML paths of the reference's DECLARE_PARAMS tables). */
int fl_n_vehicle_params(int model);

/* Problem description; replaces the constructor arguments of Optimal_laptime<Dynamic_model_t>
 * (src/core/applications/optimal_laptime.h:98-120) once vehicle and track have been evaluated at the mesh. */
typedef struct fl_problem_desc
{
    int model;                  /* FL_MODEL_*                                                                          */
    int form;                   /* FL_FORM_*                                                                           */
    int closed;                 /* 1: closed circuit (n_elements = n_points), 0: open section (first node fixed)       */
    int elevation;              /* 1: track with pitch/roll/z (F1 only; all 13 equations become differential)          */
    int kart_direct_torque;     /* kart: rear-axle control is a torque (enable_direct_torque) instead of throttle      */
    int n_points;
    const double* s;            /* [n_points] arclength mesh, strictly increasing                                      */
    double track_length;
    const double* track_nodes;  /* [n_points][6]: yaw, pitch, roll, yaw', pitch', roll' at the mesh nodes              */
    const double* wl;           /* [n_points] left track limit  (road.lateral-displacement >= -wl)                    */
    const double* wr;           /* [n_points] right track limit                                                        */
    int batch;                  /* >= 1 independent problems sharing mesh and layout                                   */
    const double* vehicle_params; /* [batch][fl_n_vehicle_params(model)]                                               */
    double sigma;               /* trapezoidal weight, 0.5 in the reference (optimal_laptime.h:33-58)                 */
    double dissipation[2];      /* of the two full-mesh controls: steering, throttle|torque                            */
    const double* q0;           /* open problems: inputs of node 0   [n_inputs]  (14 F1, 13 kart)                      */
    const double* u0;           /* open problems: controls of node 0 [n_controls] (4 F1, 2 kart)                       */
    const double* du0;          /* open problems, derivative form: control rates of node 0 [2]                         */
    const double* ctrl_default; /* NULL, or the value of every control [n_controls].  Not a way to set the controls that are not
                                   optimised: the node programs take boost = 0 and the brake bias from vehicle parameter 18
                                   (chassis/brake_bias); creation fails if this field asks for anything else               */
    int device;                 /* CUDA device ordinal                                                                 */
    const double* node_params;  /* NULL, or [batch][n_points][fl_n_vehicle_params(model)]: the vehicle parameters AT EVERY NODE --
                                   parameters that vary with the arclength (Model_parameter::operator()(s),
                                   src/core/foundation/model_parameter.h:51-77, evaluated per node at
                                   src/core/vehicles/dynamic_model_car.hpp:62-63).  vehicle_params is then ignored.       */
    double wind[2];             /* northward, eastward wind velocity [m/s] (vehicle/chassis/aerodynamics/wind_velocity/...,
                                   src/core/chassis/chassis.h:276-277; enters the aerodynamic velocity, chassis.hpp:264-288) */
    /* One quantity that couples the whole lap in the reference -- a control optimised as a constant or on a hypermesh
       (Control_variables::control_array_at_s, src/core/applications/detail/optimal_laptime_control_variables.h:190-222) or a restricted
       integral quantity (limebeer2014f1.h:284-306, optimal_laptime.hpp:1316-1317, 1352-1353, 1409-1416) -- restated node-wise so the
       NLP keeps its block structure: every node gains two variables (a, q: the value at the node / the integral up to it, and a free
       jump) and every element one trapezoid row; see DESIGN.md.  F1, direct form, flat closed laps. */
    int aug_kind;               /* FL_AUG_NONE, FL_AUG_BRAKE_BIAS (control constant over stretches), FL_AUG_INTEGRAL              */
    const double* aug_jump;     /* FL_AUG_BRAKE_BIAS: [n_points], 1 where the node starts a new stretch of the hypermesh, else 0   */
    double aug_bounds[2];       /* bounds of the control / of the integral quantity                                                */
    double aug_weights[5];      /* FL_AUG_INTEGRAL: weights of engine-energy, tire-fl/-fr/-rl/-rr-energy in the integrand         */
} fl_problem_desc;
enum { FL_AUG_NONE = 0, FL_AUG_BRAKE_BIAS = 1, FL_AUG_INTEGRAL = 2 };

fl_nlp* fl_nlp_create(const fl_problem_desc* desc);       /* NULL on failure */

/* Steady-state (G-G diagram) NLPs: a batch of independent 8-variable problems
 *     x = [w_axle, z, phi, mu, psi, delta, ax, ay]           (layout of Max_lat_acc, src/core/applications/steady_state.hpp:899)
 *     g = lot2016kart::steady_state_equations(x, ax, ay, v)   (src/core/vehicles/lot2016kart.h:124-182): 6 equations + 6 tyre-slip rows
 *     f = cax * ax + cay * ay
 * replacing the functors Solve / Max_lat_acc / Max_lon_acc / Min_lon_acc of src/core/applications/steady_state.h:99-284.
 * spec [batch][7] = (v, ax, ay, free_ax, free_ay, cax, cay): with free_ax == 0 the equations use the given ax and the
 * variable x[6] is inert (likewise ay); Solve: free 0/0, c 0/0; Max_lat_acc: free 1/1, cay -1; Max_lon_acc: free 1/0,
 * cax -1; Min_lon_acc: free 1/0, cax +1.  The handle is an fl_nlp with n = 8, m = 12: every callback, fl_eval_all and the
 * interior-point solver work on it as on a lap (it is a one-node lap to them). */
typedef struct fl_steady_desc
{
    int model;                    /* FL_MODEL_KART_6DOF, or FL_MODEL_F1_3DOF: 13 variables [kappa x4, psi/10deg, Fz x4, delta/10deg,
                                     throttle, ax, ay] (ax, ay in g), 15 rows: limebeer2014f1::steady_state_equations,
                                     src/core/vehicles/limebeer2014f1.h:103-178 */
    int batch;
    const double* vehicle_params; /* [batch][fl_n_vehicle_params(model)]                 */
    const double* spec;           /* [batch][7]                                          */
    int device;
} fl_steady_desc;
fl_nlp* fl_steady_create(const fl_steady_desc* desc);
void    fl_nlp_destroy(fl_nlp* nlp);
const char* fl_last_error(void);

/* get_nlp_info: sizes of ONE problem of the batch */
void fl_nlp_sizes(const fl_nlp* nlp, int* n, int* m, int* nnz_jac, int* nnz_hess);
int  fl_nlp_batch(const fl_nlp* nlp);
const char* fl_nlp_variant(const fl_nlp* nlp);
/* get_bounds_info (optimal_laptime.hpp:326-483): model default bounds, track limits on the lateral displacement */
int fl_nlp_bounds(const fl_nlp* nlp, double* x_l, double* x_u, double* g_l, double* g_u);
/* the same from the description alone (host-only: no device is touched); only model, form, closed, elevation,
 * kart_direct_torque, n_points, wl, wr and vehicle_params (problem 0) are read */
int fl_problem_bounds(const fl_problem_desc* desc, double* x_l, double* x_u, double* g_l, double* g_u);
/* fixed sparsity (triplets), same for every problem of the batch */
int fl_nlp_structure(const fl_nlp* nlp, int* jac_row, int* jac_col, int* hess_row, int* hess_col);
/* per-node slots of the device workspace written by the values pass: node values handed to the element assembly, and
 * checkpointed intermediates (sin/cos/atan/tanh/exp/reciprocal/sqrt) the derivative pass loads */
void fl_nlp_workspace_slots(const fl_nlp* nlp, int* n_values, int* n_checkpoints);

/* Device memory of destroyed handles is kept in a per-device cache of blocks (exact-size reuse by the next handles; at most FL_CACHE_BYTES,
 * default 64 GB, blocks above half of that are returned at once; released and retried when an allocation fails).  This empties it. */
void fl_release_cached_memory(void);

/* Output channels of the vehicle model along the lap -- what the reference evaluates with get_outputs_map() after a solve
 * (src/main/c/fastestlapc.cpp:1643-1690: tyre slips, forces, dissipation, omega, contact-point velocities: src/core/tire/tire.h:170-196;
 * chassis accelerations, resultant force, aerodynamic forces: src/core/chassis/chassis.h:282-308, chassis.hpp:228-244) -- at x [batch][n]
 * (host): out [batch][fl_nlp_n_outputs][n_points] (host), channel k named fl_nlp_output_name(nlp, k). */
int fl_nlp_n_outputs(const fl_nlp* nlp);
const char* fl_nlp_output_name(const fl_nlp* nlp, int k);
int fl_eval_outputs(fl_nlp* nlp, const double* x, double* out);
/* fp64 operations per node of each pass (values, jac, hess, all): the algorithmic work the roofline is quoted on */
void fl_nlp_ops_per_node(const fl_nlp* nlp, int ops[4]);

/* Ipopt C-interface callbacks (problem 0 of the batch); user_data = fl_nlp*.  The signatures are those of Ipopt 3.14's
 * IpStdCInterface.h (Eval_F_CB, Eval_Grad_F_CB, Eval_G_CB, Eval_Jac_G_CB, Eval_H_CB: ipindex = int, ipnumber = double, C99 bool;
 * x and lambda are not const there), so the five functions can be handed to CreateIpoptProblem as they are --
 * tests/ipopt_callbacks_check.c compiles exactly that assignment against the restated typedefs. */
bool fl_eval_f     (int n, double* x, bool new_x, double* obj_value, void* user_data);
bool fl_eval_grad_f(int n, double* x, bool new_x, double* grad_f, void* user_data);
bool fl_eval_g     (int n, double* x, bool new_x, int m, double* g, void* user_data);
bool fl_eval_jac_g (int n, double* x, bool new_x, int m, int nele_jac, int* iRow, int* jCol, double* values, void* user_data);
bool fl_eval_h     (int n, double* x, bool new_x, double obj_factor, int m, double* lambda, bool new_lambda,
                    int nele_hess, int* iRow, int* jCol, double* values, void* user_data);

/* Whole batch, one call: every output may be NULL.  Arrays are [batch][...] contiguous. */
int fl_eval_all(fl_nlp* nlp, const double* x, const double* lambda, double obj_factor,
                double* f, double* g, double* grad_f, double* jac_values, double* hess_values);

/* Device-resident path (what an on-device interior-point loop uses): pointers are device pointers on the handle's
 * device and stay owned by the handle; inputs are read from fl_dev_x / fl_dev_lambda.  `what` is a bit mask. */
enum { FL_EVAL_FG = 1, FL_EVAL_JAC = 2, FL_EVAL_HESS = 4,
       FL_EVAL_NO_VALUES = 8 };  /* measurement only: skip the values pass and run the derivative kernel alone on the checkpoints the previous call left */
double* fl_dev_x(fl_nlp* nlp);          /* [batch][n] */
double* fl_dev_lambda(fl_nlp* nlp);     /* [batch][m] */
const double* fl_dev_f(fl_nlp* nlp);    /* [batch]    */
const double* fl_dev_g(fl_nlp* nlp);    /* [batch][m] */
const double* fl_dev_grad(fl_nlp* nlp); /* [batch][n] */
const double* fl_dev_jac(fl_nlp* nlp);  /* [batch][nnz_jac]  */
const double* fl_dev_hess(fl_nlp* nlp); /* [batch][nnz_hess] */
int  fl_eval_device(fl_nlp* nlp, int what, double obj_factor);   /* enqueue on the handle's stream, no synchronisation */
int  fl_nlp_sync(fl_nlp* nlp);
void* fl_nlp_stream(fl_nlp* nlp);                                /* cudaStream_t */
/* Times `reps` back-to-back enqueues of `what` with CUDA events on the handle's stream; returns ms per enqueue (< 0 on error). */
double fl_time_device(fl_nlp* nlp, int what, double obj_factor, int reps);
/* `steps` evaluations back to back on one stream, step k on instance k mod count (a ring of independent instances of the same
 * problem, so that no step finds its data in L2), one pair of events around the sequence: device ms, -1 on error. */
double fl_time_device_ring(fl_nlp** nlps, int count, int what, double obj_factor, int steps);
long fl_kernel_launches(const fl_nlp* nlp);                      /* kernels launched by this handle so far */
/* Writes a scratch buffer larger than the L2 cache on the handle's stream (benchmark hygiene between timed iterations). */
int  fl_flush_l2(fl_nlp* nlp);
/* FP64 (non-tensor) pipe peak of the handle's device, measured with a DFMA micro-kernel: fused multiply-adds per second */
double fl_measure_fp64_peak(fl_nlp* nlp);
/* The node programs' branch-free elementary functions (csrc/fl_math.cuh) evaluated on the device, for the accuracy tests:
 * fn 0 reciprocal, 1 sqrt, 2 sin, 3 cos, 4 atan, 5 tanh, 6 exp; x and y are host arrays of n doubles. */
int fl_math_eval(int device, int fn, const double* x, double* y, long n);
/* Page-locked host memory for the host-buffer entry points (pageable memory works too, at lower PCIe throughput). */
void* fl_host_alloc(size_t bytes);
void  fl_host_free(void* p);

/* ---- interior-point solver -------------------------------------------------------------------------------------------
 * Device-resident replacement of what Ipopt + MUMPS do behind lion::ipopt_cppad_solve for this NLP
 * (src/core/applications/optimal_laptime.hpp:528-556: options, solve, result): a primal-dual interior-point method with
 * filter line search (Waechter & Biegler 2006, the algorithm Ipopt implements) whose augmented system is factorised as a
 * batched block-(cyclic-)tridiagonal LDL^T, one warp per lap.  Every lap of the handle's batch is solved independently
 * and concurrently.  Defaults = Ipopt's defaults with the reference's tolerances (tol, constr_viol_tol 1e-10,
 * acceptable_tol 1e-8, max_iter 3000). */
typedef struct fl_ipm fl_ipm;
typedef struct fl_ipm_opts
{
    double tol, constr_viol_tol, acceptable_tol;
    int acceptable_iter, max_iter;
    double mu_init, bound_push, bound_frac, bound_relax_factor, kappa_eps, kappa_mu, theta_mu, tau_min, kappa_sigma, s_max;
    double gamma_theta, gamma_phi, delta, s_theta, s_phi, eta_phi, kappa_soc, alpha_red;
    int max_soc, max_backtracks;
    double delta_w_min, delta_w_0, delta_w_max, kappa_w_minus, kappa_w_plus, kappa_w_plus_bar, delta_c_bar, kappa_c, y_init_max;
    int refine_steps;
} fl_ipm_opts;
void fl_ipm_default_options(fl_ipm_opts* opts);
/* status of a lap (ipopt_cppad_result::status analogue) */
enum { FL_IPM_RUNNING = 0, FL_IPM_SUCCESS = 1, FL_IPM_ACCEPTABLE = 2, FL_IPM_MAX_ITER = -1, FL_IPM_LINE_SEARCH_FAILED = -2,
       FL_IPM_REGULARISATION_FAILED = -3, FL_IPM_NOT_FINITE = -4 };
/* opts == NULL: defaults.  Bounds x_l, x_u [n], g_l, g_u [m] (shared by the batch) == NULL: fl_nlp_bounds.  The nlp handle must
 * outlive the solver and must not be used concurrently. */
fl_ipm* fl_ipm_create(fl_nlp* nlp, const fl_ipm_opts* opts, const double* x_l, const double* x_u, const double* g_l, const double* g_u);
/* The same with one set of bounds per problem of the batch when bounds_per_lap != 0: x_l, x_u [batch][n], g_l, g_u [batch][m]
 * (G-G sweeps: the reference narrows the bounds of the longitudinal acceleration per speed, steady_state.hpp:733-735, 787-789). */
fl_ipm* fl_ipm_create_ex(fl_nlp* nlp, const fl_ipm_opts* opts, const double* x_l, const double* x_u, const double* g_l, const double* g_u, int bounds_per_lap);
void    fl_ipm_destroy(fl_ipm* ipm);
/* x0: [batch][n] starting points (host).  Returns 1 when the iteration ran (per-lap outcomes are in fl_ipm_results). */
int fl_ipm_solve(fl_ipm* ipm, const double* x0, int verbose);
/* Warm start from a previous solution (what the reference passes to ipopt_cppad_solve when warm_start is set,
 * optimal_laptime.hpp:550-551): lambda0 [batch][m], zl0, zu0 [batch][n]; push = Ipopt's warm_start_bound_push (1e-3);
 * mu_init = starting barrier parameter. */
int fl_ipm_solve_warm(fl_ipm* ipm, const double* x0, const double* lambda0, const double* zl0, const double* zu0, double push, double mu_init, int verbose);
/* any output may be NULL: x [batch][n], y [batch][m], zl, zu [batch][n], obj [batch], status [batch], iters [batch],
 * errors [batch][4] = (scaled NLP error, constraint violation, dual infeasibility, barrier parameter) */
int fl_ipm_results(fl_ipm* ipm, double* x, double* y, double* zl, double* zu, double* obj, int* status, int* iters, double* errors);
/* Barrier terms at the points the solver holds (after a solve: the returned ones), with its own relaxed bounds, slacks and slack
 * multipliers: sigx [batch][n], sigs [batch][n_elements * 6], dc [batch] (constraint regularisation of the last iteration); any may be
 * NULL.  Together with fl_kkt_factor_solve: linear systems with the KKT matrix of the solution, i.e. the parameter sensitivities
 * dx/dp = -K^-1 d(grad L, c)/dp of the reference's Sensitivity_analysis (optimal_laptime.hpp:564-588). */
int fl_ipm_sigma(fl_ipm* ipm, double* sigx, double* sigs, double* dc);
/* [0] solver kernel launches, [1] factorisation launches, [2] solve launches, [3] full callback rounds, [4] f/g-only rounds, [5] outer rounds */
void fl_ipm_stats(const fl_ipm* ipm, long stats[6], double* device_ms);
/* The factorisation has two schedules: one block per lap walking the whole chain (throughput: many laps), or the chain cut
 * at ~sqrt(n_stages) separators whose segments are factorised in parallel (latency: few laps).  policy 0 = automatic (the
 * second one when at most max_laps laps are still active), 1 = always the second, -1 = never.  Returns the number of separators. */
int fl_ipm_set_partitioned(fl_ipm* ipm, int policy, int max_laps);
/* Mapping of the first schedule: mode 1 = one warp per lap with the stage matrix in registers (batches), -1 = one block of
 * four warps per lap, 0 = automatic (warp per lap for batches of 64 laps or more).  Returns 1 if warp per lap is in use. */
int fl_ipm_set_warp_kkt(fl_ipm* ipm, int mode);
/* One factorisation + solve of  [W + diag(sigx) + dw, J^T; J, -D][dx; dy] = [r1; r2]  at (x, lambda) for every lap of the batch
 * (D = dc on equality rows, 1/(sigs + dw) + dc on inequality rows; sigs is [batch][n_elements * 6], element-major);
 * ls_mode = 1 solves the least-squares multiplier system instead (W = 0, sigx = 1, D = 1 on inequality rows).
 * n_neg / n_zero: negative and vanishing pivots of the condensed matrix (expected: n_stages * n_equality_rows, 0). */
int fl_kkt_factor_solve(fl_ipm* ipm, const double* x, const double* lambda, const double* sigx, const double* sigs, const double* dw, const double* dc,
                        const double* r1, const double* r2, int refine_steps, int ls_mode, double* dx, double* dy, int* n_neg, int* n_zero);

#ifdef __cplusplus
}
#endif
#endif
