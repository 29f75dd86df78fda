"""The two vehicle models of the hot path, written once against the tracing scalar `E` (expr.py).

What the reference evaluates per collocation node, `car(q, u, s) -> (states, dstates_dt)`
(src/core/vehicles/dynamic_model_car.hpp:52-108), with its lion::Frame tree unrolled into closed form:

  limebeer-2014-f1 (3DOF + 4 algebraic loads)   chassis_car_3dof.hpp:117-239, axle_car_3dof.hpp:219-303,
                                                tire.hpp:79-103, tire_pacejka.hpp:183-226, chassis.hpp:248-288
  roberto-lot-2016-kart (6DOF)                  chassis_car_6dof.hpp:91-250, axle_car_6dof.hpp:109-206,
                                                tire.hpp:56-75, tire_pacejka.hpp:95-179
  curvilinear road                              road_curvilinear.hpp:5-21, road_curvilinear.h:49-60

Parameter vectors follow include/fastestlap_b200.h (FL_F1_* / FL_KART_* enums).
"""
import math
from .expr import E, sqrt, sin, cos, atan, tanh, exp

G0 = 9.81
DEG = math.pi / 180.0

F1_PARAMS = [
    "mass", "Ixx", "Iyy", "Izz", "rho", "area", "cd", "cl", "com_x", "com_y", "com_z",
    "front_axle_x", "front_axle_z", "rear_axle_x", "rear_axle_z", "aero_x", "aero_y", "aero_z",
    "brake_bias", "roll_balance", "Fz_max_ref2",
    "f_track", "f_inertia", "f_smooth_throttle", "f_brake_Tmax",
    "r_track", "r_inertia", "r_smooth_throttle", "r_diff_stiffness", "r_brake_Tmax", "r_engine_power", "r_boost_power",
] + [a + n for a in ("ft_", "rt_") for n in
     ("R0", "Fz1", "Fz2", "mux1", "mux2", "kmax1", "kmax2", "muy1", "muy2", "lmax1_deg", "lmax2_deg", "Qx", "Qy")]
# extra host-computed entries appended after the XML parameters (value-dependent C++ branches of the reference that
# cannot be traced: axle_car_3dof.hpp:294 `I < 1e-10 ? 1/m : 1`)
F1_HOST_PARAMS = ["f_wheel_scale", "r_wheel_scale", "sigma"]

KART_TIRE = ("R0", "kt", "ct", "Fz0", "lambdaFz0", "Fz_max_ref2", "pCx1", "pDx1", "pEx1", "pKx1", "pKx2", "pKx3", "rBx1", "rCx1",
             "pCy1", "pDy1", "pEy1", "pKy1", "pKy2", "pKy4", "rBy1", "rCy1")
KART_PARAMS = [
    "mass", "Ixx", "Iyy", "Izz", "Ixz", "rho", "cd", "cl", "area", "com_x", "com_y", "com_z",
    "front_axle_x", "front_axle_z", "rear_axle_x", "rear_axle_z",
    "f_track", "f_kchassis", "f_kantiroll", "f_beta_l", "f_beta_r",
    "r_track", "r_kchassis", "r_kantiroll", "r_inertia", "r_smooth_throttle", "r_brake_Tmax", "r_engine_power",
] + [a + n for a in ("ft_", "rt_") for n in KART_TIRE]
KART_HOST_PARAMS = ["sigma"]


def smooth_pos(x, eps2):
    return 0.5 * (x + sqrt(x * x + eps2))


def smooth_sign(x, eps):
    # lion smooth_sign: tanh(x/eps) reproduces the reference's golden files
    return tanh(x / eps)


def curvilinear_road(u, v, omega, n, alpha, kz):
    ca, sa = cos(alpha), sin(alpha)
    dtds = (1.0 - n * kz) / (u * ca - v * sa)           # road_curvilinear.hpp:9
    dn = u * sa + v * ca                                # :12
    dalpha = omega - kz / dtds                          # :15
    return ca, sa, dtds, dn, dalpha


def f1_tire(P, pre, kappa_hat, lam, N):
    """Pacejka_simple_model (tire_pacejka.hpp:196-226); kappa = kappa_hat * kappa_max(N) (tire_pacejka.hpp:70)."""
    Fz1, Fz2 = P[pre + "Fz1"], P[pre + "Fz2"]
    dF = 1.0 / (Fz2 - Fz1)
    lin = lambda a1, a2: (N - Fz1) * ((a2 - a1) * dF) + a1
    mu_x_max = smooth_pos(lin(P[pre + "mux1"], P[pre + "mux2"]) - 1.0, 1.0e-5) + 1.0
    mu_y_max = smooth_pos(lin(P[pre + "muy1"], P[pre + "muy2"]) - 1.0, 1.0e-5) + 1.0
    kappa_max = lin(P[pre + "kmax1"], P[pre + "kmax2"])
    lambda_max = lin(P[pre + "lmax1_deg"] * DEG, P[pre + "lmax2_deg"] * DEG)
    kappa = kappa_hat * kappa_max
    kappa_n = kappa_hat                        # (kappa_hat*kappa_max)/kappa_max
    lambda_n = lam / lambda_max
    rho = sqrt(kappa_n * kappa_n + lambda_n * lambda_n + 1.0e-12)
    Qx, Qy = P[pre + "Qx"], P[pre + "Qy"]
    Sx = (0.5 * math.pi) / atan(Qx)
    Sy = (0.5 * math.pi) / atan(Qy)
    mu_x = mu_x_max * sin(Qx * atan(Sx * rho))
    mu_y = mu_y_max * sin(Qy * atan(Sy * rho))
    inv_rho = 1.0 / rho
    Fx = mu_x * N * kappa_n * inv_rho
    Fy = mu_y * N * lambda_n * inv_rho
    return kappa, Fx, Fy


def tire_kappa_max(P, pre, N):
    Fz1, Fz2 = P[pre + "Fz1"], P[pre + "Fz2"]
    return (N - Fz1) * ((P[pre + "kmax2"] - P[pre + "kmax1"]) * (1.0 / (Fz2 - Fz1))) + P[pre + "kmax1"]


def f1_model(q, u, P, trk):
    """q[14], u[4]: E; P: dict name->E; trk: dict with kx, ky, kz, smu, cmu, sph, cph (E or float)."""
    k_hat = q[0:4]
    vx, vy, omega = q[4], q[5], q[6]
    fz_g = q[7:11]
    time, n, alpha = q[11], q[12], q[13]
    delta, boost, throttle, bias = u
    m = P["mass"]
    kx, ky, kz = trk["kx"], trk["ky"], trk["kz"]
    smu, cmu, sph, cph = trk["smu"], trk["cmu"], trk["sph"], trk["cph"]

    ca, sa, dtds, dn, dalpha = curvilinear_road(vx, vy, omega, n, alpha, kz)
    inv_dtds = 1.0 / dtds
    w = n * kx * inv_dtds                                # road_curvilinear.hpp:20
    # road-frame angular velocity in body axes (chassis.hpp:119-131; test/chassis/chassis_car_3dof_3d_test.cpp:90-94):
    # p_road = kx/dtds, q_road = ky/dtds
    p_road, q_road = kx * inv_dtds, ky * inv_dtds
    omega_x = ca * p_road + sa * q_road
    omega_y = ca * q_road - sa * p_road
    omega_z = omega

    mg = m * G0
    FzN = [f * mg for f in fz_g]                                        # chassis_car_3dof.hpp:291-300
    N = [smooth_pos(-F, P["Fz_max_ref2"]) for F in FzN]                 # :131-134

    xa = [P["front_axle_x"]] * 2 + [P["rear_axle_x"]] * 2
    za = [P["front_axle_z"]] * 2 + [P["rear_axle_z"]] * 2
    yt = [-0.5 * P["f_track"], 0.5 * P["f_track"], -0.5 * P["r_track"], 0.5 * P["r_track"]]    # axle_car_3dof.hpp:31
    pre = ["ft_", "ft_", "rt_", "rt_"]
    cd, sd = cos(delta), sin(delta)

    lam, omega_w, Fx_t, Fy_t, rc, vt = [], [], [], [], [], []
    Fax, Tax = [], []
    for i in range(4):
        R0 = P[pre[i] + "R0"]
        # tyre frame origin (xa, yt, za) in the road frame; deformation w_t = za + R0; contact point (0,0,R0 - w_t): on the road plane
        rci = R0 - (za[i] + R0)
        zc = za[i] + rci
        vcx = vx + omega_y * zc - omega_z * yt[i]
        vcy = vy + omega_z * xa[i] - omega_x * zc
        if i < 2:
            vtx, vty = cd * vcx + sd * vcy, cd * vcy - sd * vcx
        else:
            vtx, vty = vcx, vcy
        lam_i = -vty / vtx                                               # tire.h:157
        vt.append((vtx, vty))
        kappa, Fx, Fy = f1_tire(P, pre[i], k_hat[i], lam_i, N[i])
        lam.append(lam_i)
        omega_w.append((1.0 + kappa) * vtx / R0)                         # tire.hpp:102
        Fx_t.append(Fx); Fy_t.append(Fy); rc.append(rci)
        Ttx, Tty = -rci * Fy, rci * Fx                                   # tire_pacejka.hpp:119
        if i < 2:
            Fax.append((cd * Fx - sd * Fy, sd * Fx + cd * Fy, -N[i]))
            Tax.append((cd * Ttx - sd * Tty, sd * Ttx + cd * Tty, 0.0))
        else:
            Fax.append((Fx, Fy, -N[i]))
            Tax.append((Ttx, Tty, 0.0))

    # axle torques: axle_car_3dof.hpp:233-261
    brake_f = smooth_pos(-throttle, P["f_smooth_throttle"])
    brake_r = smooth_pos(-throttle, P["r_smooth_throttle"])
    Tw = [-smooth_sign(omega_w[0], 1.0) * (brake_f * P["f_brake_Tmax"]) * bias,
          -smooth_sign(omega_w[1], 1.0) * (brake_f * P["f_brake_Tmax"]) * bias,
          -smooth_sign(omega_w[2], 1.0) * (brake_r * P["r_brake_Tmax"]) * (1.0 - bias),
          -smooth_sign(omega_w[3], 1.0) * (brake_r * P["r_brake_Tmax"]) * (1.0 - bias)]
    thr = smooth_pos(throttle, P["r_smooth_throttle"])
    inv_omega_mean = 1.0 / (0.5 * (omega_w[2] + omega_w[3]))
    engine_torque = thr * (P["r_engine_power"] * 1.0e3) * inv_omega_mean      # engine.hpp:41-45
    boost_torque = (thr * boost) * (P["r_boost_power"] * 1.0e3) * inv_omega_mean
    diff_torque = P["r_diff_stiffness"] * (omega_w[2] - omega_w[3])
    Tw[2] = Tw[2] + 0.5 * (engine_torque + boost_torque) - diff_torque
    Tw[3] = Tw[3] + 0.5 * (engine_torque + boost_torque) + diff_torque
    dL = [Tw[i] - Fx_t[i] * rc[i] for i in range(4)]                     # tire.h:107

    # axle resultants (axle_car_3dof.hpp:270-273), tyre origins at (0, yt, 0) in the axle frame
    Fa, Ta = [], []
    for a in range(2):
        l, r = 2 * a, 2 * a + 1
        Fa.append(tuple(Fax[l][c] + Fax[r][c] for c in range(3)))
        Ta.append((Tax[l][0] + yt[l] * Fax[l][2] + Tax[r][0] + yt[r] * Fax[r][2],
                   Tax[l][1] + Tax[r][1],
                   -(yt[l] * Fax[l][0]) - yt[r] * Fax[r][0]))

    # aerodynamics (chassis.hpp:264-288).  Wind: (eastward, -northward, 0) in the absolute frame, brought to body axes with the
    # transpose of the road frame's rotation (:269-274).  Its components along the track's tangent / normal / binormal (wt, wn,
    # wb) are per-node constants; the yaw alpha of the car in the road frame is the only rotation left for the node program.
    wt, wn, wb = trk.get("wt", 0.0), trk.get("wn", 0.0), trk.get("wb", 0.0)
    av = (-vx + (ca * wt + sa * wn), -vy + (ca * wn - sa * wt), -w + wb)
    avn = sqrt(av[0] * av[0] + av[1] * av[1])
    qd = 0.5 * P["rho"] * P["cd"] * P["area"]
    Faero = [qd * av[0] * avn, qd * av[1] * avn, qd * av[2] * avn + (0.5 * P["rho"] * P["cl"] * P["area"]) * (av[0] * av[0])]

    def cross(a, b):
        return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

    # chassis resultants about the road-frame origin (chassis_car_3dof.hpp:145-164)
    xf = (P["front_axle_x"], 0.0, P["front_axle_z"])
    xr = (P["rear_axle_x"], 0.0, P["rear_axle_z"])
    xcom = (P["com_x"], P["com_y"], P["com_z"])
    xae = (P["aero_x"] - xcom[0], P["aero_y"] - xcom[1], P["aero_z"] - xcom[2])
    F = [Fa[0][c] + Fa[1][c] + Faero[c] for c in range(3)]
    c1, c2, c3 = cross(xf, Fa[0]), cross(xr, Fa[1]), cross(xae, Faero)
    c4 = cross(xcom, (-F[0], -F[1], -F[2]))
    T = [Ta[0][c] + c1[c] + Ta[1][c] + c2[c] + c3[c] + c4[c] for c in range(3)]
    # gravity (chassis.hpp:248-260)
    F[0] = F[0] + mg * (sa * sph * cmu - ca * smu)
    F[1] = F[1] + mg * (sa * smu + ca * sph * cmu)
    F[2] = F[2] + mg * cph * cmu

    Ixx, Iyy, Izz = P["Ixx"], P["Iyy"], P["Izz"]
    h = -P["com_z"]
    inv_m = 1.0 / m
    du = F[0] * inv_m + (vy + h * omega_x) * omega_z - w * omega_y        # chassis_car_3dof.hpp:186-188
    dv = F[1] * inv_m - (vx - h * omega_y) * omega_z + w * omega_x
    dw = F[2] * inv_m + (vx - h * omega_y) * omega_y - (vy + h * omega_x) * omega_x
    dLroll = T[0] + (Iyy - Izz) * omega_z * omega_y                       # :195-197
    dLpitch = T[1] + (Izz - Ixx) * omega_z * omega_x
    domega = (T[2] + (Ixx - Iyy) * omega_x * omega_y) / Izz
    D = P["roll_balance"]
    roll_balance = (FzN[1] - FzN[0]) * (1.0 - D) + (FzN[2] - FzN[3]) * D  # :203

    inv_mg = 1.0 / mg
    sc = [P["f_wheel_scale"]] * 2 + [P["r_wheel_scale"]] * 2              # axle_car_3dof.hpp:294
    Iw = [P["f_inertia"]] * 2 + [P["r_inertia"]] * 2
    states = [omega_w[i] * (Iw[i] * sc[i]) for i in range(4)]
    dstates = [dL[i] * sc[i] for i in range(4)]
    states += [vx - h * omega_y, vy + h * omega_x, omega, w * (1.0 / G0), (Ixx * omega_x) * inv_mg, (Iyy * omega_y) * inv_mg,
               0.0 * vx, time, n, alpha]
    dstates += [du, dv, domega, dw * (1.0 / G0), dLroll * inv_mg, dLpitch * inv_mg, roll_balance * inv_mg, 1.0 + 0.0 * vx, dn, dalpha]
    dstates = [d * dtds for d in dstates]                                # dynamic_model_car.hpp:104-105
    half_track = 0.5 * P["f_track"]
    extra = [lam[0], lam[1], lam[2], lam[3], n + half_track * ca, n - half_track * ca]     # limebeer2014f1.h:266-281
    # un-scaled time derivatives and wheel balances (the steady-state equations use them, limebeer2014f1.h:148-169)
    rates = dict(du=du, dv=dv, domega=domega, dw_g=dw * (1.0 / G0), dLroll_mg=dLroll * inv_mg, dLpitch_mg=dLpitch * inv_mg,
                 roll_balance_mg=roll_balance * inv_mg, dL=dL, lam=lam)
    # integrands of the integral quantities (limebeer2014f1.h:291-306): engine power [MW] (engine.hpp:43), and minus the power
    # each tyre dissipates, F . v_slip with v_slip = (v_x - omega R0, v_y) at the contact point (tire.h:127)
    integrands = [thr * (P["r_engine_power"] * 1.0e3) * 1.0e-6]
    for i in range(4):
        integrands.append(-(Fx_t[i] * (vt[i][0] - omega_w[i] * P[pre[i] + "R0"]) + Fy_t[i] * vt[i][1]) * 1.0e-6)
    # output channels (get_outputs_map: tire.h:170-196, chassis.h:282-308; accelerations chassis.hpp:228-244)
    outputs = []
    tyre_names = ["front-axle.left-tire", "front-axle.right-tire", "rear-axle.left-tire", "rear-axle.right-tire"]
    kap = [k_hat[i] * tire_kappa_max(P, pre[i], N[i]) for i in range(4)]
    for i, nm in enumerate(tyre_names):
        outputs += [(nm + ".lambda", lam[i]), (nm + ".true_kappa", kap[i]),
                    (nm + ".dissipation", Fx_t[i] * (vt[i][0] - omega_w[i] * P[pre[i] + "R0"]) + Fy_t[i] * vt[i][1]),
                    (nm + ".force.x", Fx_t[i]), (nm + ".force.y", Fy_t[i]), (nm + ".force.z", -N[i]), (nm + ".omega", omega_w[i]),
                    (nm + ".velocity.x", vt[i][0]), (nm + ".velocity.y", vt[i][1])]
    inv_speed = 1.0 / sqrt(vx * vx + vy * vy)
    outputs += [("chassis.acceleration.x", (vx * F[0] + vy * F[1]) * inv_m * inv_speed),
                ("chassis.acceleration.y", (vx * F[1] - vy * F[0]) * inv_m * inv_speed),
                ("chassis.acceleration.yaw", domega),
                ("chassis.force.x", F[0]), ("chassis.force.y", F[1]), ("chassis.force.z", F[2]),
                ("chassis.aerodynamics.lift", (0.5 * P["rho"] * P["cl"] * P["area"]) * (av[0] * av[0])),
                ("chassis.aerodynamics.drag.x", qd * av[0] * avn), ("chassis.aerodynamics.drag.y", qd * av[1] * avn)]
    return dict(states=states, dstates=dstates, dtds=dtds, extra=extra, rates=rates, integrands=integrands, outputs=outputs)


def kart_tire(P, pre, kappa, lam, N, only_lateral):
    """Pacejka_standard_model (tire_pacejka.hpp:123-179)."""
    p = lambda n: P[pre + n]
    Fz0p = p("lambdaFz0") * p("Fz0")
    if only_lateral:
        # front tyres: pDx1 == 0 -> force_pure_longitudinal_magic returns 0 (tire_pacejka.hpp:132-133)
        Fx = None
    else:
        dfz = (N - Fz0p) / Fz0p
        Bx = (p("pKx1") + p("pKx2") * dfz) * exp(p("pKx3") * dfz) / (p("pCx1") * p("pDx1"))
        Bk = Bx * kappa
        Fx0 = p("pDx1") * N * sin(p("pCx1") * atan(Bk - p("pEx1") * (Bk - atan(Bk))))
        Fx = cos(p("rCx1") * atan(p("rBx1") * lam)) * Fx0
    By = (p("pKy1") * Fz0p * sin(p("pKy4") * atan(N / (Fz0p * p("pKy2"))))) / (p("pCy1") * p("pDy1") * N)
    Bl = By * lam
    Fy0 = p("pDy1") * N * sin(p("pCy1") * atan(Bl - p("pEy1") * (Bl - atan(Bl))))
    if only_lateral:
        Fy = Fy0          # kappa == 0: cos(rCy1 atan(0)) = 1
    else:
        Fy = cos(p("rCy1") * atan(p("rBy1") * kappa)) * Fy0
    return Fx, Fy


def kart_model(q, u, P, trk, direct_torque=False, curvilinear=True):
    omega_axle, vx, vy, Om, z, phi, mu, dz, dphi, dmu = q[0:10]
    r0, r1, r2 = q[10], q[11], q[12]          # time, n, alpha  |  x, y, psi
    delta, throttle = u
    m = P["mass"]

    if curvilinear:
        ca, sa, dtds, dr1, dr2 = curvilinear_road(vx, vy, Om, r1, r2, trk["kz"])
        dr0 = None
    else:
        dtds = None
        dr0 = vx * cos(r2) - vy * sin(r2)      # road_cartesian.hpp:7-14
        dr1 = vx * sin(r2) + vy * cos(r2)
        dr2 = Om

    if direct_torque:                          # axle_car_6dof.hpp:186-206
        T_ax = throttle
    else:
        thr = smooth_pos(throttle, P["r_smooth_throttle"])
        brk = smooth_pos(-throttle, P["r_smooth_throttle"])
        T_ax = thr * (P["r_engine_power"] * 1.0e3) / omega_axle - smooth_sign(omega_axle, 1.0) * (brk * P["r_brake_Tmax"])

    cd, sd = cos(delta), sin(delta)
    F = [0.0, 0.0, 0.0]
    T = [0.0, 0.0, 0.0]
    kappa_l, lambda_l = [], []
    tyre_out = []
    sum_long_torque = 0.0
    for a, ax in enumerate(("f_", "r_")):
        pre = "ft_" if a == 0 else "rt_"
        R0, kt = P[pre + "R0"], P[pre + "kt"]
        xa_ = P["front_axle_x"] if a == 0 else P["rear_axle_x"]
        za_ = P["front_axle_z"] if a == 0 else P["rear_axle_z"]
        track, kch, kar = P[ax + "track"], P[ax + "kchassis"], P[ax + "kantiroll"]
        axle_z = P["com_z"] + z + za_ - mu * xa_        # chassis_car_6dof.h:152-156 on top of the chassis frame origin
        daxle_z = dz - dmu * xa_
        displ_sym, ddispl_sym = axle_z + R0, daxle_z     # axle_car_6dof.hpp:127-157
        displ_asym, ddispl_asym = 0.5 * track * phi, 0.5 * track * dphi
        if a == 0:
            displ_asym = displ_asym + P["f_beta_r"] * delta
        sym = 1.0 / (kch + kt)
        asym = 2.0 * kar * kt * sym / (2.0 * kar + kch + kt)
        ksym = kch * sym
        wl = ksym * (displ_sym - displ_asym) - displ_asym * asym
        wr = ksym * (displ_sym + displ_asym) + displ_asym * asym
        dwl = ksym * (ddispl_sym - ddispl_asym) - asym * ddispl_asym
        dwr = ksym * (ddispl_sym + ddispl_asym) + asym * ddispl_asym
        s_t = (wl - displ_sym + displ_asym, wr - displ_sym - displ_asym)
        ds_t = (dwl - ddispl_sym + ddispl_asym, dwr - ddispl_sym - ddispl_asym)
        y_t = (-0.5 * track, 0.5 * track)
        beta = (P["f_beta_l"], P["f_beta_r"]) if a == 0 else (None, None)
        for side in range(2):
            tz = s_t[side] + phi * y_t[side]             # axle_car_6dof.h:253-263
            if a == 0:
                tz = tz + beta[side] * delta
            dtz = ds_t[side] + dphi * y_t[side]
            w = axle_z + tz + R0                         # tire.hpp:62-64
            dw = daxle_z + dtz
            vcx = vx - Om * y_t[side]
            vcy = vy + Om * xa_
            if a == 0:
                vtx, vty = cd * vcx + sd * vcy, cd * vcy - sd * vcx
            else:
                vtx, vty = vcx, vcy
            inv_vtx = 1.0 / vtx
            lam = -vty * inv_vtx                         # tire.h:157
            if a == 0:
                kappa = None                             # "only lateral": omega = vx/R0 -> kappa = 0 (tire.hpp:69-70)
            else:
                kappa = (omega_axle * R0 - vtx) * inv_vtx    # tire.h:154
            N = smooth_pos(kt * w + P[pre + "ct"] * dw, P[pre + "Fz_max_ref2"])    # tire_pacejka.hpp:100
            Fx, Fy = kart_tire(P, pre, kappa, lam, N, a == 0)
            kappa_l.append(kappa); lambda_l.append(lam)
            tyre_out.append(dict(lam=lam, kappa=kappa, Fx=Fx, Fy=Fy, N=N, vtx=vtx, vty=vty, omega=(omega_axle if a == 1 else vtx * (1.0 / R0)), R0=R0))
            if a == 1:
                sum_long_torque = sum_long_torque - Fx * (R0 - w)    # tire.h:107
            if a == 0:
                Fr = (-(sd * Fy), cd * Fy, -N)           # Fx == 0
            else:
                Fr = (Fx, Fy, -N)
            # total lever arm of the contact point about the road-frame origin is (xa, y_t, 0)
            F = [F[c] + Fr[c] for c in range(3)]
            T = [T[0] + y_t[side] * Fr[2], T[1] - xa_ * Fr[2], T[2] + (xa_ * Fr[1] - y_t[side] * Fr[0])]

    domega_axle = (T_ax + sum_long_torque) / P["r_inertia"]              # axle_car_6dof.hpp:160-164
    F[2] = F[2] + m * G0                                                  # chassis_car_6dof.hpp:127
    if curvilinear and trk is not None and "wt" in trk:
        wt, wn = trk["wt"], trk["wn"]                                     # wind along the track's tangent / normal (see f1_model)
        av = (-vx + (ca * wt + sa * wn), -vy + (ca * wn - sa * wt))
    else:
        av = (-vx, -vy)
    avn = sqrt(av[0] * av[0] + av[1] * av[1])
    qd = 0.5 * P["rho"] * P["cd"] * P["area"]
    Fa = (qd * av[0] * avn, qd * av[1] * avn, (0.5 * P["rho"] * P["cl"] * P["area"]) * (av[0] * av[0]))
    F = [F[c] + Fa[c] for c in range(3)]
    cx, cy, cz = P["com_x"], P["com_y"], P["com_z"] + z                   # chassis frame origin (:135-136)
    T = [T[0] + (cy * Fa[2] - cz * Fa[1]), T[1] + (cz * Fa[0] - cx * Fa[2]), T[2] + (cx * Fa[1] - cy * Fa[0])]

    inv_m = 1.0 / m
    du = Om * vy + F[0] * inv_m                                           # :139-146, 157-164
    dv = F[1] * inv_m - Om * vx
    d2z = F[2] * inv_m
    Ixx, Iyy, Izz, Ixz = P["Ixx"], P["Iyy"], P["Izz"], P["Ixz"]
    h = -P["com_z"]
    lhs0 = m * (dv * h + (h - z) * Om * vx) - (Iyy - Izz + Ixx) * Om * dmu - (Iyy - Izz) * Om * Om * phi      # :167-188
    lhs1 = m * ((z - h) * du + h * Om * vy) + (Iyy - Izz + Ixx) * Om * dphi - ((Ixx - Izz) * mu + Ixz) * Om * Om
    lhs2 = 2.0 * Ixz * dmu * Om
    b0, b1, b2 = T[0] - lhs0, T[1] - lhs1, T[2] - lhs2
    m02 = -(Ixx - Izz) * mu - Ixz                                         # :191-200
    m12 = (Iyy - Izz) * phi
    m22 = Izz + 2.0 * Ixz * mu
    det = Ixx * m22 + Ixz * m02
    x2 = (Ixx * b2 + Ixz * b0) / det
    x0 = (b0 - m02 * x2) / Ixx
    x1 = (b1 - m12 * x2) / Iyy

    states = [omega_axle, vx, vy, Om, z, phi, mu, dz, dphi, dmu, r0, r1, r2]
    dstates = [domega_axle, du, dv, x2, dz, dphi, dmu, d2z, x0, x1, (1.0 + 0.0 * vx) if curvilinear else dr0, dr1, dr2]
    if curvilinear:
        dstates = [d * dtds for d in dstates]
    extra = [kappa_l[2], kappa_l[3], lambda_l[0], lambda_l[1], lambda_l[2], lambda_l[3]]      # lot2016kart.h:200-211
    # output channels (tire.h:170-196, chassis.h:282-308); the front tyres are "only lateral": kappa = 0, no longitudinal force
    outputs = []
    for t, nm in zip(tyre_out, ["front-axle.left-tire", "front-axle.right-tire", "rear-axle.left-tire", "rear-axle.right-tire"]):
        zero = 0.0 * vx
        fx = zero if t["Fx"] is None else t["Fx"]
        outputs += [(nm + ".lambda", t["lam"]), (nm + ".true_kappa", zero if t["kappa"] is None else t["kappa"]),
                    (nm + ".dissipation", fx * (t["vtx"] - t["omega"] * t["R0"]) + t["Fy"] * t["vty"]),
                    (nm + ".force.x", fx), (nm + ".force.y", t["Fy"]), (nm + ".force.z", -t["N"]), (nm + ".omega", t["omega"]),
                    (nm + ".velocity.x", t["vtx"]), (nm + ".velocity.y", t["vty"])]
    inv_speed = 1.0 / sqrt(vx * vx + vy * vy)
    outputs += [("chassis.acceleration.x", (vx * F[0] + vy * F[1]) * inv_m * inv_speed),
                ("chassis.acceleration.y", (vx * F[1] - vy * F[0]) * inv_m * inv_speed),
                ("chassis.acceleration.yaw", x2),
                ("chassis.force.x", F[0]), ("chassis.force.y", F[1]), ("chassis.force.z", F[2]),
                ("chassis.aerodynamics.lift", Fa[2]), ("chassis.aerodynamics.drag.x", Fa[0]), ("chassis.aerodynamics.drag.y", Fa[1])]
    return dict(states=states, dstates=dstates, dtds=dtds, extra=extra, force=F, torque=T,
                kappa=kappa_l, lam=lambda_l, outputs=outputs)
