// Types shared by the KKT and interior-point kernels.
#pragma once

constexpr int IPM_FILTER_MAX = 256;

enum { IPM_RUNNING = 0, IPM_SUCCESS = 1, IPM_ACCEPTABLE = 2, IPM_MAX_ITER = -1, IPM_LINE_SEARCH_FAILED = -2, IPM_REGULARISATION_FAILED = -3,
       IPM_NOT_FINITE = -4 };
enum { LS_IDLE = 0, LS_TRIAL = 1, LS_SOC_SOLVE = 2, LS_SOC_TRIAL = 3, LS_DONE = 4 };

struct fl_ipm_options        // layout of fl_ipm_opts (include/fastestlap_b200.h)
{
    double tol, constr_viol_tol, acceptable_tol;
    int acceptable_iter, max_iter;
    double mu_init, bound_push, bound_frac, bound_relax_factor, kappa_eps, kappa_mu, theta_mu, tau_min, kappa_sigma, s_max;
    double gamma_theta, gamma_phi, delta, s_theta, s_phi, eta_phi, kappa_soc, alpha_red;
    int max_soc, max_backtracks;
    double delta_w_min, delta_w_0, delta_w_max, kappa_w_minus, kappa_w_plus, kappa_w_plus_bar, delta_c_bar, kappa_c, y_init_max;
    int refine_steps;
};

struct fl_ipm_lap            // state of one lap's iteration
{
    double mu, tau, theta, phi, dphi, alpha, alpha_max, alpha_dual, dw, dw_last, dc, theta_max, theta_min;
    double f, cviol, dual_inf, compl_inf, err0, theta_soc_old, alpha_soc;
    double dw_prev;          // regularisation the previous iteration's factorisation ended with
    int status, ls, iter, acc_count, n_filter, n_backtrack, n_soc, n_reg, n_fact, switching, n_restart, n_small, pad_;
    // diagnosis of the inertia-correcting factorisation (warp-per-lap schedule): stage at which the last failing attempt stopped, stages
    // walked by failing / successful attempts, and how far (cyclically) consecutive failures were apart
    int fail_stage, pad2_;
    long long stages_fail, stages_ok, fail_shift, n_fail;
    double filter[2 * IPM_FILTER_MAX];
};
