/* TEST INFRASTRUCTURE.  The callback typedefs of Ipopt 3.14's C interface (IpStdCInterface.h of coin-or/Ipopt, releases/3.14.x:
 * the version the reference links, quickstart.rst:77), restated here because Ipopt is not in this image: ipindex and ipnumber
 * are int and double in the default build, Bool became C99 bool in 3.14, and x / lambda are not const.  Only the typedefs the
 * five callbacks must match, and the prototype of CreateIpoptProblem that takes them. */
#ifndef IPSTDCINTERFACE_3_14_MIRROR_H
#define IPSTDCINTERFACE_3_14_MIRROR_H
#include <stdbool.h>
typedef int ipindex;
typedef double ipnumber;
typedef void* UserDataPtr;
typedef bool (*Eval_F_CB)(ipindex n, ipnumber* x, bool new_x, ipnumber* obj_value, UserDataPtr user_data);
typedef bool (*Eval_Grad_F_CB)(ipindex n, ipnumber* x, bool new_x, ipnumber* grad_f, UserDataPtr user_data);
typedef bool (*Eval_G_CB)(ipindex n, ipnumber* x, bool new_x, ipindex m, ipnumber* g, UserDataPtr user_data);
typedef bool (*Eval_Jac_G_CB)(ipindex n, ipnumber* x, bool new_x, ipindex m, ipindex nele_jac, ipindex* iRow, ipindex* jCol, ipnumber* values, UserDataPtr user_data);
typedef bool (*Eval_H_CB)(ipindex n, ipnumber* x, bool new_x, ipnumber obj_factor, ipindex m, ipnumber* lambda, bool new_lambda, ipindex nele_hess,
                          ipindex* iRow, ipindex* jCol, ipnumber* values, UserDataPtr user_data);
struct IpoptProblemInfo;
typedef struct IpoptProblemInfo* IpoptProblem;
IpoptProblem CreateIpoptProblem(ipindex n, ipnumber* x_L, ipnumber* x_U, ipindex m, ipnumber* g_L, ipnumber* g_U, ipindex nele_jac, ipindex nele_hess,
                                ipindex index_style, Eval_F_CB eval_f, Eval_G_CB eval_g, Eval_Grad_F_CB eval_grad_f, Eval_Jac_G_CB eval_jac_g, Eval_H_CB eval_h);
#endif
