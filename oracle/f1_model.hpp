// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
//
// CPU restatement of the reference's limebeer-2014-f1 3DOF car on a curvilinear road, templated on the
// scalar so that the same source gives values (double) and exact derivatives (Jet2<N>).
// The lion::Frame tree of the reference is unrolled by hand; each block cites the reference lines it follows.
// Paths are relative to /root/reference/src/core.
#ifndef ORACLE_F1_MODEL_HPP
#define ORACLE_F1_MODEL_HPP

#include <cmath>
#include "jet.hpp"

namespace oracle {

constexpr double g0 = 9.81;                           // lion constant, pinned by test/chassis/chassis_car_3dof_test.cpp:182-188
constexpr double DEG = 3.14159265358979323846 / 180.0;
constexpr double PI  = 3.14159265358979323846;

inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan(double x) { return std::atan(x); }
inline double tanh(double x) { return std::tanh(x); }
inline double exp(double x) { return std::exp(x); }
// lion smooth_pos(x, eps2), written out in test/chassis/chassis_car_3dof_test.cpp:81-84
template <class T> inline T smooth_pos(const T& x, double eps2) { return 0.5 * (x + sqrt(x * x + eps2)); }
// lion smooth_sign(x, eps): saturating sign; tanh(x/eps) reproduces the golden files, assumed.
template <class T> inline T smooth_sign(const T& x, double eps) { return tanh(x / eps); }

// Per-node track data (plain doubles: vehicles/road_curvilinear.hpp:104-135 evaluates the track in scalar)
struct TrackNode
{
    double theta, mu, phi;        // Frenet frame Euler angles (yaw, pitch, roll)
    double dtheta, dmu, dphi;     // their arclength derivatives
    double wind_north = 0.0, wind_east = 0.0;   // wind velocity [m/s] (chassis.h:276-277)
    // wind (eastward, -northward, 0) along the track's tangent / normal / binormal (road_curvilinear.hpp:121-131): the rows of the
    // transposed road-frame rotation of chassis.hpp:269-274, before the car's own yaw alpha
    void wind_in_track_axes(double& wt, double& wn, double& wb) const
    {
        const double wE = wind_east, wS = -wind_north;
        const double ct = std::cos(theta), st = std::sin(theta), cm = std::cos(mu), sm = std::sin(mu), cp = std::cos(phi), sp = std::sin(phi);
        wt = ct * cm * wE + st * cm * wS;
        wn = (ct * sm * sp - st * cp) * wE + (st * sm * sp + ct * cp) * wS;
        wb = (ct * sm * cp + st * sp) * wE + (st * sm * cp - ct * sp) * wS;
    }
};

// Vehicle parameter vector of the F1 (database/vehicles/f1/limebeer-2014-f1.xml paths in comments)
enum F1Param
{
    F1_MASS = 0, F1_IXX, F1_IYY, F1_IZZ, F1_RHO, F1_AREA, F1_CD, F1_CL,
    F1_COM_X, F1_COM_Y, F1_COM_Z, F1_FRONT_AXLE_X, F1_FRONT_AXLE_Z, F1_REAR_AXLE_X, F1_REAR_AXLE_Z,
    F1_AERO_X, F1_AERO_Y, F1_AERO_Z, F1_BRAKE_BIAS, F1_ROLL_BALANCE, F1_FZ_MAX_REF2,
    F1_F_TRACK, F1_F_INERTIA, F1_F_SMOOTH_THROTTLE, F1_F_BRAKE_TMAX,
    F1_R_TRACK, F1_R_INERTIA, F1_R_SMOOTH_THROTTLE, F1_R_DIFF_STIFFNESS, F1_R_BRAKE_TMAX, F1_R_ENGINE_POWER, F1_R_BOOST_POWER,
    F1_FT_R0, F1_FT_FZ1, F1_FT_FZ2, F1_FT_MUX1, F1_FT_MUX2, F1_FT_KMAX1, F1_FT_KMAX2, F1_FT_MUY1, F1_FT_MUY2,
    F1_FT_LMAX1_DEG, F1_FT_LMAX2_DEG, F1_FT_QX, F1_FT_QY,
    F1_RT_R0, F1_RT_FZ1, F1_RT_FZ2, F1_RT_MUX1, F1_RT_MUX2, F1_RT_KMAX1, F1_RT_KMAX2, F1_RT_MUY1, F1_RT_MUY2,
    F1_RT_LMAX1_DEG, F1_RT_LMAX2_DEG, F1_RT_QX, F1_RT_QY,
    F1_NPARAMS
};

constexpr int F1_NIN = 14;     // inputs: kappa x4, u, v, omega, Fz_g x4, time, n, alpha  (test/vehicles/limebeer2014f1_test.cpp:32-70)
constexpr int F1_NCTRL = 4;    // controls: delta, boost, throttle, brake-bias
constexpr int F1_NEXTRA = 6;   // vehicles/limebeer2014f1.h:181
enum { F1_IKFL = 0, F1_IKFR, F1_IKRL, F1_IKRR, F1_IU, F1_IV, F1_IOMEGA, F1_IFZFL, F1_IFZFR, F1_IFZRL, F1_IFZRR, F1_ITIME, F1_IN, F1_IALPHA };
enum { F1_CDELTA = 0, F1_CBOOST, F1_CTHROTTLE, F1_CBIAS };

template <class T>
struct F1Out
{
    T states[F1_NIN];
    T dstates[F1_NIN];     // already multiplied by dt/ds (vehicles/dynamic_model_car.hpp:104-105)
    T dtds;
    T extra[F1_NEXTRA];    // vehicles/limebeer2014f1.h:266-281
    // diagnostics used by the unit-level tests
    T kappa[4], lambda[4], omega_w[4], Fx_t[4], Fy_t[4], N[4];
    T dL[4];               // wheel angular-momentum balance, unscaled (axle_car_3dof.hpp:263-268)
    // integrands of the integral quantities (vehicles/limebeer2014f1.h:291-306), per unit time: engine power [MW], minus the power
    // dissipated by each tyre [MW] (tire/tire.h:127: F . v_slip)
    T integrands[5];
    // output channels (get_outputs_map: tire/tire.h:170-196, chassis/chassis.h:282-308, chassis.hpp:228-244), in the order of
    // fastest_lap_b200's fl_nlp_output_name: per tyre (fl, fr, rl, rr) lambda, true_kappa, dissipation, force.x, .y, .z, omega, velocity.x,
    // .y; then chassis acceleration.x, .y, .yaw, force.x, .y, .z, aerodynamics.lift, .drag.x, .drag.y
    T channels[45];
};

// Pacejka_simple_model, tire/tire_pacejka.hpp:196-226 and tire/tire_pacejka.h:291-294
struct SimpleTire { double R0, Fz1, Fz2, mux1, mux2, kmax1, kmax2, muy1, muy2, lmax1, lmax2, Qx, Qy, Sx, Sy; };
inline SimpleTire load_tire(const double* P)
{
    SimpleTire t;
    t.R0 = P[0]; t.Fz1 = P[1]; t.Fz2 = P[2]; t.mux1 = P[3]; t.mux2 = P[4]; t.kmax1 = P[5]; t.kmax2 = P[6];
    t.muy1 = P[7]; t.muy2 = P[8];
    t.lmax1 = P[9] * DEG; t.lmax2 = P[10] * DEG;               // tire_pacejka.hpp:186-187
    t.Qx = P[11]; t.Qy = P[12];
    t.Sx = PI / (2.0 * std::atan(t.Qx)); t.Sy = PI / (2.0 * std::atan(t.Qy));   // tire_pacejka.hpp:190-191
    return t;
}
template <class T> inline T tire_kappa_max(const SimpleTire& t, const T& Fz) { return (Fz - t.Fz1) * (t.kmax2 - t.kmax1) / (t.Fz2 - t.Fz1) + t.kmax1; }
template <class T> inline T tire_lambda_max(const SimpleTire& t, const T& Fz) { return (Fz - t.Fz1) * (t.lmax2 - t.lmax1) / (t.Fz2 - t.Fz1) + t.lmax1; }

template <class T>
inline void tire_forces(const SimpleTire& t, const T& kappa, const T& lambda, const T& Fz, T& Fx, T& Fy)
{
    const double mu_min = 1.0;   // tire_pacejka.h:318
    const T mu_x_max = smooth_pos((Fz - t.Fz1) * (t.mux2 - t.mux1) / (t.Fz2 - t.Fz1) + t.mux1 - mu_min, 1.0e-5) + mu_min;
    const T mu_y_max = smooth_pos((Fz - t.Fz1) * (t.muy2 - t.muy1) / (t.Fz2 - t.Fz1) + t.muy1 - mu_min, 1.0e-5) + mu_min;
    const T kappa_max = tire_kappa_max(t, Fz);
    const T lambda_max = tire_lambda_max(t, Fz);
    const T kappa_n = kappa / kappa_max;
    const T lambda_n = lambda / lambda_max;
    const T rho = sqrt(kappa_n * kappa_n + lambda_n * lambda_n + 1.0e-12);
    const T mu_x = mu_x_max * sin(t.Qx * atan(t.Sx * rho));
    const T mu_y = mu_y_max * sin(t.Qy * atan(t.Sy * rho));
    Fx = mu_x * Fz * kappa_n / rho;
    Fy = mu_y * Fz * lambda_n / rho;
}

// (states, dstates_dt) = car(q, u, s): vehicles/dynamic_model_car.hpp:52-108 for limebeer2014f1::curvilinear
template <class T>
inline void f1_car(const T q[F1_NIN], const T u[F1_NCTRL], const double* P, const TrackNode& tk, F1Out<T>& out)
{
    const double m = P[F1_MASS];
    const SimpleTire tf = load_tire(P + F1_FT_R0), tr = load_tire(P + F1_RT_R0);

    const T& vx = q[F1_IU];
    const T& vy = q[F1_IV];
    const T& omega = q[F1_IOMEGA];
    const T& n = q[F1_IN];
    const T& alpha = q[F1_IALPHA];
    const T& delta = u[F1_CDELTA];
    const T& boost = u[F1_CBOOST];
    const T& throttle = u[F1_CTHROTTLE];
    const T& bias = u[F1_CBIAS];

    // --- road: vehicles/road_curvilinear.h:49-60 (curvature), road_curvilinear.hpp:5-21 (update)
    const double kx = tk.dphi - std::sin(tk.mu) * tk.dtheta;
    const double ky = std::cos(tk.phi) * tk.dmu + std::cos(tk.mu) * std::sin(tk.phi) * tk.dtheta;
    const double kz = -std::sin(tk.phi) * tk.dmu + std::cos(tk.mu) * std::cos(tk.phi) * tk.dtheta;

    const T ca = cos(alpha), sa = sin(alpha);
    const T dtds = (1.0 - n * kz) / (vx * ca - vy * sa);
    const T dn = vx * sa + vy * ca;
    const T dalpha = omega - kz / dtds;
    const T w = n * kx / dtds;                       // ground vertical velocity

    // Road frame absolute angular velocity in body axes: rotations Z(theta) Y(mu) X(phi) Z(alpha)
    // (chassis/chassis.hpp:10,119-131; closed form in test/chassis/chassis_car_3dof_3d_test.cpp:90-94)
    const T dtheta_dt = tk.dtheta / dtds, dmu_dt = tk.dmu / dtds, dphi_dt = tk.dphi / dtds;   // road_curvilinear.h:88
    const T p_road = dphi_dt - std::sin(tk.mu) * dtheta_dt;
    const T q_road = std::cos(tk.phi) * dmu_dt + std::cos(tk.mu) * std::sin(tk.phi) * dtheta_dt;
    const T omega_x = ca * p_road + sa * q_road;
    const T omega_y = q_road * ca - sa * p_road;
    const T& omega_z = omega;

    // --- chassis_car_3dof.hpp:291-300, 131-134
    T N[4];
    const int ifz[4] = { F1_IFZFL, F1_IFZFR, F1_IFZRL, F1_IFZRR };
    T FzN[4];
    for (int i = 0; i < 4; ++i)
    {
        FzN[i] = q[ifz[i]] * g0 * m;
        N[i] = smooth_pos(-FzN[i], P[F1_FZ_MAX_REF2]);     // -neg_force_z
    }

    // --- tyres (tire/tire.hpp:79-103): contact point velocities, kappa, lambda, omega
    const double xa[4] = { P[F1_FRONT_AXLE_X], P[F1_FRONT_AXLE_X], P[F1_REAR_AXLE_X], P[F1_REAR_AXLE_X] };
    const double za[4] = { P[F1_FRONT_AXLE_Z], P[F1_FRONT_AXLE_Z], P[F1_REAR_AXLE_Z], P[F1_REAR_AXLE_Z] };
    const double yt[4] = { -0.5 * P[F1_F_TRACK], 0.5 * P[F1_F_TRACK], -0.5 * P[F1_R_TRACK], 0.5 * P[F1_R_TRACK] };   // axle_car_3dof.hpp:31
    const SimpleTire* tires[4] = { &tf, &tf, &tr, &tr };
    const int ikappa[4] = { F1_IKFL, F1_IKFR, F1_IKRL, F1_IKRR };
    const T cd = cos(delta), sd = sin(delta);

    T Fax[4][3], Tax[4][3];     // tyre force / torque (about tyre frame origin) in axle axes
    T rc[4];
    for (int i = 0; i < 4; ++i)
    {
        const SimpleTire& t = *tires[i];
        // tyre frame origin in road frame: (xa, yt, za); point (0,0,R0) -> w = za + R0 (tire.hpp:89-92); contact point (0,0,R0-w)
        const double wdef = za[i] + t.R0;
        rc[i] = T(t.R0 - wdef);
        const double zc = za[i] + (t.R0 - wdef);          // contact point height in road frame (=0)
        // velocity of the contact point in road axes: v0 + omega x r
        const T vcx = vx + omega_y * zc - omega_z * yt[i];
        const T vcy = vy + omega_z * xa[i] - omega_x * zc;
        T vtx, vty;
        if (i < 2) { vtx = cd * vcx + sd * vcy; vty = cd * vcy - sd * vcx; }
        else       { vtx = vcx; vty = vcy; }
        out.lambda[i] = -vty / vtx;                                         // tire.h:157
        const T kappa = q[ikappa[i]] * tire_kappa_max(t, N[i]);              // tire_pacejka.hpp:70
        out.kappa[i] = kappa;
        out.omega_w[i] = (1.0 + kappa) * vtx / t.R0;                         // tire.hpp:102
        out.N[i] = N[i];
        T Fx, Fy;
        tire_forces(t, kappa, out.lambda[i], N[i], Fx, Fy);                  // tire_pacejka.hpp:107-120
        out.Fx_t[i] = Fx; out.Fy_t[i] = Fy;
        out.integrands[1 + i] = -(Fx * (vtx - out.omega_w[i] * t.R0) + Fy * vty) * 1.0e-6;      // tire.h:127, limebeer2014f1.h:295-301
        {
            T* ch = out.channels + 9 * i;
            ch[0] = out.lambda[i]; ch[1] = kappa; ch[2] = Fx * (vtx - out.omega_w[i] * t.R0) + Fy * vty; ch[3] = Fx; ch[4] = Fy; ch[5] = -N[i];
            ch[6] = out.omega_w[i]; ch[7] = vtx; ch[8] = vty;
        }
        // torque at wheel centre: cross((0,0,rc), F)
        const T Ttx = -rc[i] * Fy, Tty = rc[i] * Fx;
        if (i < 2)
        {
            Fax[i][0] = cd * Fx - sd * Fy; Fax[i][1] = sd * Fx + cd * Fy; Fax[i][2] = -N[i];
            Tax[i][0] = cd * Ttx - sd * Tty; Tax[i][1] = sd * Ttx + cd * Tty; Tax[i][2] = T(0.0);
        }
        else
        {
            Fax[i][0] = Fx; Fax[i][1] = Fy; Fax[i][2] = -N[i];
            Tax[i][0] = Ttx; Tax[i][1] = Tty; Tax[i][2] = T(0.0);
        }
    }

    // --- axles: chassis/axle_car_3dof.hpp:219-274
    T dL[4];
    {
        const T brake_f = smooth_pos(-throttle, P[F1_F_SMOOTH_THROTTLE]);
        const T brake_r = smooth_pos(-throttle, P[F1_R_SMOOTH_THROTTLE]);
        T Tq[4];
        Tq[0] = -smooth_sign(out.omega_w[0], 1.0) * (brake_f * P[F1_F_BRAKE_TMAX]) * bias;
        Tq[1] = -smooth_sign(out.omega_w[1], 1.0) * (brake_f * P[F1_F_BRAKE_TMAX]) * bias;
        const T bias_r = 1.0 - bias;                                          // chassis_car_3dof.hpp:138
        Tq[2] = -smooth_sign(out.omega_w[2], 1.0) * (brake_r * P[F1_R_BRAKE_TMAX]) * bias_r;
        Tq[3] = -smooth_sign(out.omega_w[3], 1.0) * (brake_r * P[F1_R_BRAKE_TMAX]) * bias_r;
        const T thr = smooth_pos(throttle, P[F1_R_SMOOTH_THROTTLE]);
        const T omega_mean = 0.5 * (out.omega_w[2] + out.omega_w[3]);
        const T engine_torque = thr * (P[F1_R_ENGINE_POWER] * 1.0e3) / omega_mean;           // actuators/engine.hpp:41-45
        out.integrands[0] = thr * (P[F1_R_ENGINE_POWER] * 1.0e3) * 1.0e-6;                    // engine.hpp:43, limebeer2014f1.h:293
        const T diff_torque = P[F1_R_DIFF_STIFFNESS] * (out.omega_w[2] - out.omega_w[3]);
        const T boost_torque = (thr * boost) * (P[F1_R_BOOST_POWER] * 1.0e3) / omega_mean;
        Tq[2] += 0.5 * (engine_torque + boost_torque) - diff_torque;
        Tq[3] += 0.5 * (engine_torque + boost_torque) + diff_torque;
        for (int i = 0; i < 4; ++i) dL[i] = Tq[i] + (-out.Fx_t[i] * rc[i]);   // tire.h:107
    }

    // axle resultants in axle frame (axle_car_3dof.hpp:270-273), tyre origin at (0, yt, 0)
    T Fa[2][3], Ta[2][3];
    for (int a = 0; a < 2; ++a)
    {
        const int l = 2 * a, r = 2 * a + 1;
        for (int c = 0; c < 3; ++c) Fa[a][c] = Fax[l][c] + Fax[r][c];
        // cross((0,y,0),F) = (y Fz, 0, -y Fx)
        Ta[a][0] = Tax[l][0] + yt[l] * Fax[l][2] + Tax[r][0] + yt[r] * Fax[r][2];
        Ta[a][1] = Tax[l][1] + Tax[r][1];
        Ta[a][2] = Tax[l][2] - yt[l] * Fax[l][0] + Tax[r][2] - yt[r] * Fax[r][0];
    }

    // --- aerodynamics: chassis/chassis.hpp:264-288; wind_body = R^T (E, -N, 0), R = road frame (track frame turned by alpha)
    double wt, wn, wb;
    tk.wind_in_track_axes(wt, wn, wb);
    const T av[3] = { -vx + (ca * wt + sa * wn), -vy + (ca * wn - sa * wt), -w + wb };
    const T avn = sqrt(av[0] * av[0] + av[1] * av[1]);
    const double qd = 0.5 * P[F1_RHO] * P[F1_CD] * P[F1_AREA];
    T Faero[3] = { qd * av[0] * avn, qd * av[1] * avn, qd * av[2] * avn };
    Faero[2] += 0.5 * P[F1_RHO] * P[F1_CL] * P[F1_AREA] * av[0] * av[0];

    // --- chassis resultants: chassis_car_3dof.hpp:145-164
    const double xf[3] = { P[F1_FRONT_AXLE_X], 0.0, P[F1_FRONT_AXLE_Z] }, xr[3] = { P[F1_REAR_AXLE_X], 0.0, P[F1_REAR_AXLE_Z] };
    const double xcom[3] = { P[F1_COM_X], P[F1_COM_Y], P[F1_COM_Z] };
    const double xae[3] = { P[F1_AERO_X] - xcom[0], P[F1_AERO_Y] - xcom[1], P[F1_AERO_Z] - xcom[2] };
    T F[3], Tq3[3];
    for (int c = 0; c < 3; ++c) F[c] = Fa[0][c] + Fa[1][c] + Faero[c];
    auto cross_ds = [](const double a[3], const T b[3], T r[3]) {
        r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
    };
    T c1[3], c2[3], c3[3], c4[3];
    cross_ds(xf, Fa[0], c1); cross_ds(xr, Fa[1], c2); cross_ds(xae, Faero, c3);
    const T mF[3] = { -F[0], -F[1], -F[2] };
    cross_ds(xcom, mF, c4);
    for (int c = 0; c < 3; ++c) Tq3[c] = Ta[0][c] + c1[c] + Ta[1][c] + c2[c] + c3[c] + c4[c];

    // gravity: chassis/chassis.hpp:248-260
    const double smu = std::sin(tk.mu), cmu = std::cos(tk.mu), sph = std::sin(tk.phi), cph = std::cos(tk.phi);
    F[0] += m * g0 * (sa * sph * cmu - ca * smu);
    F[1] += m * g0 * (sa * smu + ca * sph * cmu);
    F[2] += T(m * g0 * cph * cmu);

    // --- equations: chassis_car_3dof.hpp:166-203
    const double Ixx = P[F1_IXX], Iyy = P[F1_IYY], Izz = P[F1_IZZ];
    const double h = -xcom[2];
    const T du = F[0] / m + (vy + h * omega_x) * omega_z - w * omega_y;
    const T dv = F[1] / m - (vx - h * omega_y) * omega_z + w * omega_x;
    const T dw = F[2] / m + (vx - h * omega_y) * omega_y - (vy + h * omega_x) * omega_x;
    const T dLroll = Tq3[0] + (Iyy - Izz) * omega_z * omega_y;
    const T dLpitch = Tq3[1] + (Izz - Ixx) * omega_z * omega_x;
    const T domega = (Tq3[2] + (Ixx - Iyy) * omega_x * omega_y) / Izz;
    const double D = P[F1_ROLL_BALANCE];
    const T roll_balance = (FzN[1] - FzN[0]) * (1.0 - D) + (FzN[2] - FzN[3]) * D;

    // --- state vectors: axle_car_3dof.hpp:287-303, chassis.hpp:379-396, chassis_car_3dof.hpp:222-239, road_curvilinear.hpp:24-41
    const double scf = (P[F1_F_INERTIA] < 1.0e-10 ? 1.0 / m : 1.0), scr = (P[F1_R_INERTIA] < 1.0e-10 ? 1.0 / m : 1.0);
    const double sc[4] = { scf, scf, scr, scr };
    const double Iw[4] = { P[F1_F_INERTIA], P[F1_F_INERTIA], P[F1_R_INERTIA], P[F1_R_INERTIA] };
    for (int i = 0; i < 4; ++i)
    {
        out.states[i] = out.omega_w[i] * Iw[i] * sc[i];
        out.dstates[i] = dL[i] * sc[i];
        out.dL[i] = dL[i];
    }
    out.states[4] = vx - h * omega_y;           out.dstates[4] = du;
    out.states[5] = vy + h * omega_x;           out.dstates[5] = dv;
    out.states[6] = omega;                      out.dstates[6] = domega;
    out.states[7] = w / g0;                     out.dstates[7] = dw / g0;
    out.states[8] = (Ixx * omega_x) / (g0 * m); out.dstates[8] = dLroll / (g0 * m);
    out.states[9] = (Iyy * omega_y) / (g0 * m); out.dstates[9] = dLpitch / (g0 * m);
    out.states[10] = T(0.0);                    out.dstates[10] = roll_balance / (g0 * m);
    out.states[11] = q[F1_ITIME];               out.dstates[11] = T(1.0);
    out.states[12] = n;                         out.dstates[12] = dn;
    out.states[13] = alpha;                     out.dstates[13] = dalpha;
    for (int i = 0; i < F1_NIN; ++i) out.dstates[i] = out.dstates[i] * dtds;     // dynamic_model_car.hpp:104-105
    out.dtds = dtds;

    {
        // chassis channels: accelerations along / across the velocity (chassis.hpp:228-244: dot and cross of (u, v, 0) with F / m)
        const T speed = sqrt(vx * vx + vy * vy);
        T* ch = out.channels + 36;
        ch[0] = (vx * F[0] + vy * F[1]) / m / speed; ch[1] = (vx * F[1] - vy * F[0]) / m / speed; ch[2] = domega;
        ch[3] = F[0]; ch[4] = F[1]; ch[5] = F[2];
        ch[6] = 0.5 * P[F1_RHO] * P[F1_CL] * P[F1_AREA] * av[0] * av[0]; ch[7] = qd * av[0] * avn; ch[8] = qd * av[1] * avn;
    }
    // extra constraints: limebeer2014f1.h:266-281
    for (int i = 0; i < 4; ++i) out.extra[i] = out.lambda[i];
    out.extra[4] = n + 0.5 * P[F1_F_TRACK] * ca;
    out.extra[5] = n - 0.5 * P[F1_F_TRACK] * ca;
}

} // namespace oracle
#endif
