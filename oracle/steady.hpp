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
#include "f1_model.hpp"

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

// ---- F1: limebeer2014f1::steady_state_equations (/root/reference/src/core/vehicles/limebeer2014f1.h:103-178)
// x = [kappa_fl, kappa_fr, kappa_rl, kappa_rr, psi / 10 deg, Fz_fl, Fz_fr, Fz_rl, Fz_rr (in m g), delta / 10 deg, throttle, ax, ay]
// with ax, ay in g (acceleration_units = g0, limebeer2014f1.h:51); 11 equations + 4 rows lambda_i / lambda_max_i in [-1, 1.2]
constexpr int FSS_NX = 13, FSS_NC = 15;

template <class T>
inline void f1_steady(const T x[FSS_NX], const double* P, const SteadySpec& sp, T c[FSS_NC], T& f)
{
    const double g0 = 9.81, v = sp.v, m = P[F1_MASS];
    const double max_yaw = 10.0 * DEG, max_steering = 10.0 * DEG;
    const T ax = x[11] * sp.free_ax + (1.0 - sp.free_ax) * sp.ax;
    const T ay = x[12] * sp.free_ay + (1.0 - sp.free_ay) * sp.ay;
    const T psi = x[4] * max_yaw;
    const T cps = cos(psi), sps = sin(psi);
    T q[F1_NIN];
    for (int i = 0; i < 4; ++i) { q[F1_IKFL + i] = x[i]; q[F1_IFZFL + i] = x[5 + i]; }
    q[F1_IU] = cps * v;
    q[F1_IV] = -(sps * v);
    q[F1_IOMEGA] = ay * (g0 / v);
    q[F1_ITIME] = T(0.0); q[F1_IN] = T(0.0); q[F1_IALPHA] = T(0.0);        // flat cartesian road: the car does not see psi
    T u[F1_NCTRL] = { x[9] * max_steering, T(0.0), x[10], T(P[F1_BRAKE_BIAS]) };
    TrackNode tk{};
    F1Out<T> o;
    f1_car<T>(q, u, P, tk, o);
    // time derivatives: the curvilinear wrapper multiplied by dt/ds = 1 / u on a straight road
    const T inv = 1.0 / o.dtds;
    const T du = o.dstates[F1_IU] * inv, dv = o.dstates[F1_IV] * inv;
    c[0] = o.dstates[7] * inv;
    c[1] = o.dstates[8] * inv;
    c[2] = o.dstates[9] * inv;
    c[3] = o.dstates[10] * inv;
    c[4] = (du * sps + dv * cps) / g0;
    c[5] = ax + (dv * sps - du * cps) / g0;
    c[6] = o.dstates[F1_IOMEGA] * inv / g0;
    for (int i = 0; i < 4; ++i) c[7 + i] = o.dL[i] / (g0 * m);
    const SimpleTire tf = load_tire(P + F1_FT_R0), tr = load_tire(P + F1_RT_R0);
    for (int i = 0; i < 4; ++i)
    {
        // maximum_lambda of the load the model is fed with, -Fz m g (limebeer2014f1.h:123-126), not of the smoothed load
        const T Fz = -(x[5 + i] * (m * g0));
        c[11 + i] = o.lambda[i] / tire_lambda_max(i < 2 ? tf : tr, Fz);
    }
    f = x[11] * sp.cax + x[12] * sp.cay;
}

} // namespace oracle
#endif
