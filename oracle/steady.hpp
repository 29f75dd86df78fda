// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// Steady-state (G-G diagram) NLP of the kart, restating
//   lot2016kart::steady_state_equations            /root/reference/src/core/vehicles/lot2016kart.h:124-182
//   Steady_state::Solve / Max_lat_acc / Max_lon_acc / Min_lon_acc functors   src/core/applications/steady_state.h:99-284
//   their constraint evaluations                    src/core/applications/steady_state.hpp:880-930
// in one 8-variable form  x = [w_axle, z, phi, mu, psi, delta, ax, ay]  (the layout of Max_lat_acc, steady_state.hpp:899):
// a problem fixes ax and/or ay by switching the corresponding variable off (free_ax / free_ay = 0: the given value is used
// and the variable does not enter any equation), and the objective is cax * ax + cay * ay
// (Solve: 0; Max_lat_acc: -ay; Max_lon_acc: -ax; Min_lon_acc: +ax).
#ifndef ORACLE_STEADY_HPP
#define ORACLE_STEADY_HPP

#include "jet.hpp"
#include "kart_model.hpp"

namespace oracle {

constexpr int KSS_NX = 8, KSS_NC = 12;

struct SteadySpec { double v, ax, ay, free_ax, free_ay, cax, cay; };

// constraints c[12] (6 equalities, 6 tyre-slip inequalities) and objective
template <class T>
inline void kart_steady(const T x[KSS_NX], const double* P, const SteadySpec& sp, T c[KSS_NC], T& f)
{
    const double v = sp.v;
    const T ax = x[6] * sp.free_ax + (1.0 - sp.free_ax) * sp.ax;
    const T ay = x[7] * sp.free_ay + (1.0 - sp.free_ay) * sp.ay;
    const T& psi = x[4];
    const T cps = cos(psi), sps = sin(psi);
    T q[KART_NIN];
    q[KART_IOMEGA_AXLE] = (x[0] + 1.0) * (v / 0.139);          // lot2016kart.h:139 (tyre radius hard-coded in the reference)
    q[KART_IU] = cps * v;
    q[KART_IV] = -(sps * v);
    q[KART_IOMEGA] = ay / v;
    q[KART_IZ] = x[1]; q[KART_IPHI] = x[2]; q[KART_IMU] = x[3];
    q[KART_IDZ] = T(0.0); q[KART_IDPHI] = T(0.0); q[KART_IDMU] = T(0.0);
    q[KART_ITIME] = T(0.0); q[KART_IN] = T(0.0); q[KART_IALPHA] = x[4];      // cartesian road: x, y, psi
    T u[KART_NCTRL] = { x[5], T(0.0) };
    TrackNode tk{};
    KartOut<T> o;
    kart_car<T>(q, u, P, tk, /*direct_torque*/ true, /*curvilinear*/ false, o);
    const T& du = o.dstates[KART_IU];
    const T& dv = o.dstates[KART_IV];
    c[0] = o.dstates[KART_IDZ];
    c[1] = o.dstates[KART_IOMEGA];
    c[2] = o.dstates[KART_IDPHI];
    c[3] = o.dstates[KART_IDMU];
    c[4] = du * sps + dv * cps;
    c[5] = ax - du * cps + dv * sps;
    c[6] = o.kappa[2]; c[7] = o.kappa[3];
    c[8] = o.lambda[2]; c[9] = o.lambda[3]; c[10] = o.lambda[0]; c[11] = o.lambda[1];
    f = x[6] * sp.cax + x[7] * sp.cay;
}

} // namespace oracle
#endif
