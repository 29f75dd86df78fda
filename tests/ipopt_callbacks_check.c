/* TEST INFRASTRUCTURE (compile only, plain C): the five callbacks of include/fastestlap_b200.h ARE Ipopt's Eval_*_CB types.  A
 * mismatch in constness, in bool / int or in the argument order is a compile error here (-Werror=incompatible-pointer-types). */
#include "IpStdCInterface_3_14.h"
#include "../include/fastestlap_b200.h"

IpoptProblem hand_the_callbacks_to_ipopt(int n, double* xl, double* xu, int m, double* gl, double* gu, int nnz_jac, int nnz_hess)
{
    Eval_F_CB f = fl_eval_f;
    Eval_Grad_F_CB gf = fl_eval_grad_f;
    Eval_G_CB g = fl_eval_g;
    Eval_Jac_G_CB jg = fl_eval_jac_g;
    Eval_H_CB h = fl_eval_h;
    return CreateIpoptProblem(n, xl, xu, m, gl, gu, nnz_jac, nnz_hess, 0, f, g, gf, jg, h);
}
