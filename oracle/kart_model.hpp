// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// CPU restatement of the reference's roberto-lot-2016-kart 6DOF car (flat road), templated on the scalar.
// The lion::Frame tree (road -> chassis -> axle -> tyre) is unrolled by hand; every frame below the road frame is
// only translated along z (plus the steering rotation of the front tyres), see
//   chassis/chassis_car_6dof.hpp:91-154, chassis/chassis_car_6dof.h:152-169,
//   chassis/axle_car_6dof.hpp:109-181, chassis/axle_car_6dof.h:253-263, tire/tire.hpp:56-75.
// Paths are relative to /root/reference/src/core.
#ifndef ORACLE_KART_MODEL_HPP
#define ORACLE_KART_MODEL_HPP

#include "f1_model.hpp"

namespace oracle {

enum KartParam
{
    K_MASS = 0, K_IXX, K_IYY, K_IZZ, K_IXZ, K_RHO, K_CD, K_CL, K_AREA,
    K_COM_X, K_COM_Y, K_COM_Z, K_FRONT_AXLE_X, K_FRONT_AXLE_Z, K_REAR_AXLE_X, K_REAR_AXLE_Z,
    K_F_TRACK, K_F_KCHASSIS, K_F_KANTIROLL, K_F_BETA_L, K_F_BETA_R,
    K_R_TRACK, K_R_KCHASSIS, K_R_KANTIROLL, K_R_INERTIA, K_R_SMOOTH_THROTTLE, K_R_BRAKE_TMAX, K_R_ENGINE_POWER,
    K_FT_BASE,                      // 22 tyre parameters (see KartTireParam)
    K_RT_BASE = K_FT_BASE + 22,
    K_NPARAMS = K_RT_BASE + 22
};
enum KartTireParam
{
    KT_R0 = 0, KT_KT, KT_CT, KT_FZ0, KT_LAMBDAFZ0, KT_FZ_MAX_REF2,
    KT_PCX1, KT_PDX1, KT_PEX1, KT_PKX1, KT_PKX2, KT_PKX3, KT_RBX1, KT_RCX1,
    KT_PCY1, KT_PDY1, KT_PEY1, KT_PKY1, KT_PKY2, KT_PKY4, KT_RBY1, KT_RCY1
};

constexpr int KART_NIN = 13;    // omega_axle, u, v, Omega, z, phi, mu, dz, dphi, dmu, time, n, alpha (test/vehicles/car_road_curvilinear.cpp:46-76)
constexpr int KART_NCTRL = 2;   // delta, throttle (or torque)
constexpr int KART_NEXTRA = 6;  // vehicles/lot2016kart.h:75
enum { KART_IOMEGA_AXLE = 0, KART_IU, KART_IV, KART_IOMEGA, KART_IZ, KART_IPHI, KART_IMU, KART_IDZ, KART_IDPHI, KART_IDMU, KART_ITIME, KART_IN, KART_IALPHA };

template <class T>
struct KartOut
{
    T states[KART_NIN];
    T dstates[KART_NIN];    // multiplied by dt/ds when the road is curvilinear
    T dtds;
    T extra[KART_NEXTRA];   // vehicles/lot2016kart.h:200-211
    T kappa[4], lambda[4], N[4], Fx_t[4], Fy_t[4], w[4];   // order: fl, fr, rl, rr
    T force[3], torque[3];
    T channels[45];         // output channels, same order as F1Out::channels
};

// Pacejka_standard_model: tire/tire_pacejka.hpp:123-179
template <class T>
inline void kart_tire_forces(const double* tp, const T& kappa, const T& lambda, const T& Fz, T& Fx, T& Fy)
{
    const double Fz0p = tp[KT_LAMBDAFZ0] * tp[KT_FZ0];                 // tire_pacejka.hpp:125
    T Fx0(0.0);
    if (tp[KT_PDX1] != 0.0)                                            // tire_pacejka.hpp:132-133
    {
        const T dfz = (Fz - Fz0p) / Fz0p;
        const T Bx = (tp[KT_PKX1] + tp[KT_PKX2] * dfz) * exp(tp[KT_PKX3] * dfz) / (tp[KT_PCX1] * tp[KT_PDX1]);
        const T Bk = Bx * kappa;
        Fx0 = tp[KT_PDX1] * Fz * sin(tp[KT_PCX1] * atan(Bk - tp[KT_PEX1] * (Bk - atan(Bk))));
    }
    // tire_pacejka.hpp:153-154: the abs(Fz) > eps branch (the tape of the reference is recorded with Fz > 0)
    const T By = (tp[KT_PKY1] * Fz0p * sin(tp[KT_PKY4] * atan(Fz / (Fz0p * tp[KT_PKY2])))) / (tp[KT_PCY1] * tp[KT_PDY1] * Fz);
    const T Bl = By * lambda;
    const T Fy0 = tp[KT_PDY1] * Fz * sin(tp[KT_PCY1] * atan(Bl - tp[KT_PEY1] * (Bl - atan(Bl))));
    const T Gx = cos(tp[KT_RCX1] * atan(tp[KT_RBX1] * lambda));
    const T Gy = cos(tp[KT_RCY1] * atan(tp[KT_RBY1] * kappa));
    Fx = Gx * Fx0;
    Fy = Gy * Fy0;
}

// (states, dstates_dt) = car(q, u, s) for lot2016kart, vehicles/dynamic_model_car.hpp:52-108.
// curvilinear = true: road_curvilinear (q[10..12] = time, n, alpha, flat track with curvature tk.dtheta);
// curvilinear = false: road_cartesian (q[10..12] = x, y, psi), dt/ds = 1 (vehicles/road_cartesian.hpp:7-14).
template <class T>
inline void kart_car(const T q[KART_NIN], const T u[KART_NCTRL], const double* P, const TrackNode& tk,
                     bool direct_torque, bool curvilinear, KartOut<T>& out)
{
    const double m = P[K_MASS];
    const T& omega_axle = q[KART_IOMEGA_AXLE];
    const T& vx = q[KART_IU];
    const T& vy = q[KART_IV];
    const T& Om = q[KART_IOMEGA];
    const T& z = q[KART_IZ];
    const T& phi = q[KART_IPHI];
    const T& mu = q[KART_IMU];
    const T& dz = q[KART_IDZ];
    const T& dphi = q[KART_IDPHI];
    const T& dmu = q[KART_IDMU];
    const T& delta = u[0];
    const T& throttle = u[1];

    // --- road
    T dtds(1.0), dr1, dr2;      // derivatives of the last two road states
    T wind_x(0.0), wind_y(0.0); // wind velocity in body axes (chassis.hpp:269-274), curvilinear road only
    if (curvilinear)
    {
        const T& n = q[KART_IN];
        const T& alpha = q[KART_IALPHA];
        const double kz = tk.dtheta;
        const T ca = cos(alpha), sa = sin(alpha);
        double wt, wn, wb;
        tk.wind_in_track_axes(wt, wn, wb);
        wind_x = ca * wt + sa * wn; wind_y = ca * wn - sa * wt;
        dtds = (1.0 - n * kz) / (vx * ca - vy * sa);       // road_curvilinear.hpp:9
        dr1 = vx * sa + vy * ca;                           // :12
        dr2 = Om - kz / dtds;                              // :15
    }
    else
    {
        const T& psi = q[KART_IALPHA];
        dr1 = vx * sin(psi) + vy * cos(psi);
        dr2 = Om;
    }

    // --- axle torque: axle_car_6dof.hpp:186-206
    T T_ax;
    if (direct_torque) T_ax = throttle;
    else
    {
        const T thr = smooth_pos(throttle, P[K_R_SMOOTH_THROTTLE]);
        const T brk = smooth_pos(-throttle, P[K_R_SMOOTH_THROTTLE]);
        T_ax = thr * (P[K_R_ENGINE_POWER] * 1.0e3) / omega_axle;            // engine.hpp:41-45
        T_ax = T_ax - smooth_sign(omega_axle, 1.0) * (brk * P[K_R_BRAKE_TMAX]);
    }

    const double* tpar[2] = { P + K_FT_BASE, P + K_RT_BASE };
    const double xa[2] = { P[K_FRONT_AXLE_X], P[K_REAR_AXLE_X] };
    const double za[2] = { P[K_FRONT_AXLE_Z], P[K_REAR_AXLE_Z] };
    const double track[2] = { P[K_F_TRACK], P[K_R_TRACK] };
    const double kch[2] = { P[K_F_KCHASSIS], P[K_R_KCHASSIS] };
    const double kar[2] = { P[K_F_KANTIROLL], P[K_R_KANTIROLL] };
    const T cd = cos(delta), sd = sin(delta);

    T F[3] = { T(0.0), T(0.0), T(0.0) }, Tq[3] = { T(0.0), T(0.0), T(0.0) };
    T sum_long_torque(0.0);

    for (int a = 0; a < 2; ++a)
    {
        const double* tp = tpar[a];
        const double R0 = tp[KT_R0], kt = tp[KT_KT];
        // axle frame origin relative to road frame: chassis origin (com + (0,0,z)) + axle position (x_a, 0, z_a - mu x_a)
        const T axle_z = P[K_COM_Z] + z + za[a] - mu * xa[a];
        const T daxle_z = dz - dmu * xa[a];
        // axle_car_6dof.hpp:127-157
        const T displ_sym = axle_z + R0;
        const T ddispl_sym = daxle_z;
        T displ_asym = 0.5 * track[a] * phi;
        const T ddispl_asym = 0.5 * track[a] * dphi;
        if (a == 0) displ_asym = displ_asym + P[K_F_BETA_R] * delta;
        const double sym_stiffness = 1.0 / (kch[a] + kt);
        const double asym_stiffness = 2.0 * kar[a] * kt * sym_stiffness / (2.0 * kar[a] + kch[a] + kt);
        const T wl = kch[a] * sym_stiffness * (displ_sym - displ_asym) - displ_asym * asym_stiffness;
        const T wr = kch[a] * sym_stiffness * (displ_sym + displ_asym) + displ_asym * asym_stiffness;
        const T dwl = kch[a] * sym_stiffness * (ddispl_sym - ddispl_asym) - asym_stiffness * ddispl_asym;
        const T dwr = kch[a] * sym_stiffness * (ddispl_sym + ddispl_asym) + asym_stiffness * ddispl_asym;
        const T s_l = wl - displ_sym + displ_asym, s_r = wr - displ_sym - displ_asym;
        const T ds_l = dwl - ddispl_sym + ddispl_asym, ds_r = dwr - ddispl_sym - ddispl_asym;

        const double y_t[2] = { -0.5 * track[a], 0.5 * track[a] };
        const double beta[2] = { a == 0 ? P[K_F_BETA_L] : 0.0, a == 0 ? P[K_F_BETA_R] : 0.0 };
        const T s_t[2] = { s_l, s_r }, ds_t[2] = { ds_l, ds_r };

        for (int side = 0; side < 2; ++side)
        {
            const int i = 2 * a + side;
            // tyre frame origin in axle frame: axle_car_6dof.h:253-263
            const T tz = s_t[side] + phi * y_t[side] + (a == 0 ? beta[side] * delta : T(0.0));
            const T dtz = ds_t[side] + dphi * y_t[side];
            // tyre deformation: z of the point (0,0,R0), and its velocity (tire.hpp:62-64), flat road at z = 0
            const T w = axle_z + tz + R0;
            const T dw = daxle_z + dtz;
            out.w[i] = w;
            // contact point (0,0,R0-w) is on the ground: velocity in road axes = (u - Om y, v + Om x)
            const T vcx = vx - Om * y_t[side];
            const T vcy = vy + Om * xa[a];
            T vtx, vty;
            if (a == 0) { vtx = cd * vcx + sd * vcy; vty = cd * vcy - sd * vcx; }
            else        { vtx = vcx; vty = vcy; }
            T omega_t = omega_axle;
            if (a == 0) omega_t = vtx / R0;                                   // "only lateral" tyres: tire.hpp:69-70
            out.kappa[i] = (omega_t * R0 - vtx) / vtx;                        // tire.h:154
            out.lambda[i] = -vty / vtx;                                       // tire.h:157
            const T N = smooth_pos(kt * w + tp[KT_CT] * dw, tp[KT_FZ_MAX_REF2]);    // tire_pacejka.hpp:100
            out.N[i] = N;
            T Fx, Fy;
            kart_tire_forces(tp, out.kappa[i], out.lambda[i], N, Fx, Fy);
            out.Fx_t[i] = Fx; out.Fy_t[i] = Fy;
            {
                T* ch = out.channels + 9 * i;          // tire.h:170-196
                ch[0] = out.lambda[i]; ch[1] = out.kappa[i]; ch[2] = Fx * (vtx - omega_t * R0) + Fy * vty; ch[3] = Fx; ch[4] = Fy; ch[5] = -N;
                ch[6] = omega_t; ch[7] = vtx; ch[8] = vty;
            }
            if (a == 1) sum_long_torque = sum_long_torque + (-Fx * (R0 - w));  // tire.h:107
            // force in road axes
            T Fr[3];
            if (a == 0) { Fr[0] = cd * Fx - sd * Fy; Fr[1] = sd * Fx + cd * Fy; }
            else        { Fr[0] = Fx; Fr[1] = Fy; }
            Fr[2] = -N;
            // torque about the road frame origin: wheel-centre torque cross((0,0,R0-w),F) + lever arm (x_a, y_t, axle_z + tz)
            // = cross((x_a, y_t, axle_z + tz + R0 - w), F) and axle_z + tz + R0 - w = 0
            const T rz = axle_z + tz + (R0 - w);
            F[0] = F[0] + Fr[0]; F[1] = F[1] + Fr[1]; F[2] = F[2] + Fr[2];
            Tq[0] = Tq[0] + (y_t[side] * Fr[2] - rz * Fr[1]);
            Tq[1] = Tq[1] + (rz * Fr[0] - xa[a] * Fr[2]);
            Tq[2] = Tq[2] + (xa[a] * Fr[1] - y_t[side] * Fr[0]);
        }
    }

    const T domega_axle = (T_ax + sum_long_torque) / P[K_R_INERTIA];          // axle_car_6dof.hpp:160-164

    // gravity and aerodynamics: chassis_car_6dof.hpp:127-136, chassis.hpp:264-288
    F[2] = F[2] + m * g0;
    const T av[3] = { -vx + wind_x, -vy + wind_y, T(0.0) };
    const T avn = sqrt(av[0] * av[0] + av[1] * av[1]);
    const double qd = 0.5 * P[K_RHO] * P[K_CD] * P[K_AREA];
    const T Fa[3] = { qd * av[0] * avn, qd * av[1] * avn, 0.5 * P[K_RHO] * P[K_CL] * P[K_AREA] * av[0] * av[0] };
    for (int c = 0; c < 3; ++c) F[c] = F[c] + Fa[c];
    const T cx = T(P[K_COM_X]), cy = T(P[K_COM_Y]), cz = P[K_COM_Z] + z;      // chassis frame origin
    Tq[0] = Tq[0] + (cy * Fa[2] - cz * Fa[1]);
    Tq[1] = Tq[1] + (cz * Fa[0] - cx * Fa[2]);
    Tq[2] = Tq[2] + (cx * Fa[1] - cy * Fa[0]);
    for (int c = 0; c < 3; ++c) { out.force[c] = F[c]; out.torque[c] = Tq[c]; }

    // Newton: chassis_car_6dof.hpp:139-146, 157-164  (omega = (0,0,Om), velocity = (u,v,0) on a flat road)
    const T du = Om * vy + F[0] / m;
    const T dv = -(Om * vx) + F[1] / m;
    const T d2z = F[2] / m;

    // Euler: chassis_car_6dof.hpp:149-153, 167-200
    const double Ixx = P[K_IXX], Iyy = P[K_IYY], Izz = P[K_IZZ], Ixz = P[K_IXZ];
    const double h = -P[K_COM_Z];
    const T lhs0 = m * (dv * h + (h - z) * Om * vx) - (Iyy - Izz + Ixx) * Om * dmu - (Iyy - Izz) * Om * Om * phi;
    const T lhs1 = m * ((z - h) * du + h * Om * vy) + (Iyy - Izz + Ixx) * Om * dphi - ((Ixx - Izz) * mu + Ixz) * Om * Om;
    const T lhs2 = 2.0 * Ixz * dmu * Om;
    const T b0 = Tq[0] - lhs0, b1 = Tq[1] - lhs1, b2 = Tq[2] - lhs2;
    // M = [[Ixx, 0, -(Ixx-Izz) mu - Ixz], [0, Iyy, (Iyy-Izz) phi], [-Ixz, 0, Izz + 2 Ixz mu]]
    const T m02 = -(Ixx - Izz) * mu - Ixz;
    const T m12 = (Iyy - Izz) * phi;
    const T m22 = Izz + 2.0 * Ixz * mu;
    // solve: rows 0 and 2 couple (x0, x2); row 1 gives x1
    const T det = Ixx * m22 + Ixz * m02;
    const T x2 = (Ixx * b2 + Ixz * b0) / det;
    const T x0 = (b0 - m02 * x2) / Ixx;
    const T x1 = (b1 - m12 * x2) / Iyy;

    {
        const T speed = sqrt(vx * vx + vy * vy);        // chassis.hpp:228-244, chassis.h:282-308
        T* ch = out.channels + 36;
        ch[0] = (vx * F[0] + vy * F[1]) / m / speed; ch[1] = (vx * F[1] - vy * F[0]) / m / speed; ch[2] = x2;
        ch[3] = F[0]; ch[4] = F[1]; ch[5] = F[2]; ch[6] = Fa[2]; ch[7] = Fa[0]; ch[8] = Fa[1];
    }
    // state vectors: axle_car_6dof.hpp:245-258, chassis.hpp:379-396, chassis_car_6dof.hpp:218-250
    out.states[0] = omega_axle; out.dstates[0] = domega_axle;
    out.states[1] = vx;         out.dstates[1] = du;
    out.states[2] = vy;         out.dstates[2] = dv;
    out.states[3] = Om;         out.dstates[3] = x2;
    out.states[4] = z;          out.dstates[4] = dz;
    out.states[5] = phi;        out.dstates[5] = dphi;
    out.states[6] = mu;         out.dstates[6] = dmu;
    out.states[7] = dz;         out.dstates[7] = d2z;
    out.states[8] = dphi;       out.dstates[8] = x0;
    out.states[9] = dmu;        out.dstates[9] = x1;
    out.states[10] = q[KART_ITIME]; out.dstates[10] = T(1.0);
    out.states[11] = q[KART_IN];    out.dstates[11] = dr1;
    out.states[12] = q[KART_IALPHA]; out.dstates[12] = dr2;
    if (!curvilinear)
    {
        // road_cartesian: states are x, y, psi with dx = u cos(psi) - v sin(psi), dy = u sin(psi) + v cos(psi)
        const T& psi = q[KART_IALPHA];
        out.dstates[10] = vx * cos(psi) - vy * sin(psi);
        out.dstates[11] = vx * sin(psi) + vy * cos(psi);
        out.dstates[12] = Om;
    }
    for (int i = 0; i < KART_NIN; ++i) out.dstates[i] = out.dstates[i] * dtds;
    out.dtds = dtds;

    out.extra[0] = out.kappa[2]; out.extra[1] = out.kappa[3];
    out.extra[2] = out.lambda[0]; out.extra[3] = out.lambda[1];
    out.extra[4] = out.lambda[2]; out.extra[5] = out.lambda[3];
}

} // namespace oracle
#endif
