"""Pins the CPU oracle against the reference's golden solution files.

Every golden optimal-laptime XML of the reference is a known-answer test of the NLP callbacks:
  * eval_g: all equality rows vanish at the stored point (the reference solves to constr_viol_tol = 1e-10),
  * eval_f: the time integral equals the stored lap time (written with 6 decimals),
  * eval_grad_f + eval_jac_g: grad f + J^T lambda - zl + zu = 0 with the stored multipliers (tol = 1e-10 scaled).
"""
import numpy as np
import pytest

from tests.helpers import load_golden, oracle_nlp, equality_rows

CASES = ["f1_catalunya_discrete", "f1_catalunya_3d", "f1_catalunya_chicane", "f1_catalunya_chicane_derivative", "kart_vendrell",
         "f1_ovaltrack_closed", "f1_ovaltrack_open", "kart_ovaltrack_derivative", "kart_catalunya_direct", "kart_catalunya_derivative",
         "kart_catalunya_derivative_throttle", "f1_catalunya_adapted", "f1_melbourne_direct", "f1_melbourne_derivative", "f1_catalunya_2022_3d", "f1_laguna_seca_3d",
         "f1_catalunya_engine_energy"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    p, ex = load_golden(request.param)
    r = oracle_nlp(p).eval(ex["x"], ex["lam"], 1.0)
    return p, ex, r


def test_equality_rows_vanish(case):
    p, ex, r = case
    g = r["g"].reshape(p.n_elements, p.ncpe)
    assert np.abs(g[:, equality_rows(p)]).max() < 1.0e-10          # the constr_viol_tol the reference solved them to


def test_laptime(case):
    p, ex, r = case
    # objective = lap time + control-rate penalty; lap time = sum h (0.5 dtds_i + 0.5 dtds_j) (optimal_laptime.hpp:616-634)
    o = oracle_nlp(p)
    p0 = p._copy(dissipation=[0.0, 0.0])
    t = oracle_nlp(p0).eval(ex["x"], None, 1.0, derivs=False)["f"]
    t0 = 0.0 if p.closed else float(ex["time"][0])      # open sections keep integrating the time input of their first node
    assert abs(t0 + t - float(ex["laptime"])) < 1.0e-6
    assert r["f"] >= t


def test_kkt_stationarity(case):
    p, ex, r = case
    jr, jc, jv = r["jac"]
    JTl = np.zeros(p.n)
    np.add.at(JTl, jc, jv * ex["lam"][jr])
    kkt = r["grad"] + JTl - ex["zl"] + ex["zu"]
    assert np.abs(kkt).max() < 5.0e-10


def test_inequality_rows_within_bounds(case):
    p, ex, r = case
    g = r["g"].reshape(p.n_elements, p.ncpe)
    n_eq = p.ncpe - 6 - (0 if p.form == 0 else 2)
    ext = g[:, n_eq:n_eq + 6]
    if p.model == 0:
        # track limits: n +- t/2 cos(alpha) in [-wl, wr] at the element's right node (limebeer2014f1.h:266-281)
        j = (np.arange(p.n_elements) + 1) % p.n_points
        assert np.all(ext[:, 4] <= p.wr[j] + 1e-6) and np.all(ext[:, 5] >= -p.wl[j] - 1e-6)
    else:
        assert np.abs(ext).max() <= 0.11 + 1e-8          # lot2016kart.h:194-211


def test_hessian_matches_finite_differences_of_gradient():
    """Second derivatives are not pinned by any reference fixture: cross-check the oracle's jets against
    central differences of its own first derivatives on a short open section."""
    p, ex = load_golden("f1_catalunya_chicane")
    o = oracle_nlp(p)
    rng = np.random.default_rng(0)
    lam = rng.standard_normal(p.m)
    x = ex["x"]
    r = o.eval(x, lam, 0.7)
    hr, hc, hv = r["hess"]
    H = np.zeros((p.n, p.n))
    np.add.at(H, (hr, hc), hv)
    H = H + np.tril(H, -1).T

    def grad_lagr(xx):
        rr = o.eval(xx, lam, 0.7)
        jr, jc, jv = rr["jac"]
        gl = 0.7 * rr["grad"]
        np.add.at(gl, jc, jv * lam[jr])
        return gl
    for k in rng.choice(p.n, 12, replace=False):
        h = 1.0e-6 * max(1.0, abs(x[k]))
        e = np.zeros(p.n); e[k] = h
        col = (grad_lagr(x + e) - grad_lagr(x - e)) / (2 * h)
        assert np.abs(col - H[:, k]).max() <= 2.0e-6 * max(1.0, np.abs(H[:, k]).max())


def test_wind_in_the_aerodynamic_force():
    """chassis_car_3dof_test.cpp:296-336 (aerodynamic_force_wind): with a northward wind of 50 km/h and an eastward one of 30 km/h
    the wind in body axes is (E cos(psi) - N sin(psi), -N cos(psi) - E sin(psi)), the aerodynamic velocity -(u, v) + wind, drag
    0.5 rho cd A |va| va and lift 0.5 rho cl A va_x^2.  The oracle's F1 node is checked through its accelerations: the wind
    changes nothing but the aerodynamic force, so m (du, dv)(wind) - m (du, dv)(calm) is the change of the drag, and the
    change of the vertical force the change of the lift.  psi = track heading + alpha."""
    from oracle import oracle as orc
    from fastest_lap_b200 import workloads
    P = workloads._vehicles()["limebeer_2014_f1"].copy()
    KMH, DEG = 1.0 / 3.6, np.pi / 180.0
    N, E = 50.0 * KMH, 30.0 * KMH
    theta, alpha = 35.0 * DEG, -20.0 * DEG
    psi = theta + alpha
    u, v = 60.0, 1.5
    q = np.array([0.01, 0.01, 0.02, 0.02, u, v, 0.1, -0.25, -0.25, -0.3, -0.3, 0.0, 0.5, alpha])
    ctl = np.array([0.02, 0.0, 0.3, P[18]])
    track6 = np.array([theta, 0.0, 0.0, 0.01, 0.0, 0.0])
    orc.set_wind(0.0, 0.0)
    calm = orc.f1_car(q, ctl, P, track6)
    orc.set_wind(N, E)
    wind = orc.f1_car(q, ctl, P, track6)
    orc.set_wind(0.0, 0.0)
    m, rho, A, cd, cl = P[0], P[4], P[5], P[6], P[7]
    wx, wy = E * np.cos(psi) - N * np.sin(psi), -N * np.cos(psi) - E * np.sin(psi)

    def aero(vax, vay):
        nrm = np.hypot(vax, vay)
        return 0.5 * rho * cd * A * nrm * vax, 0.5 * rho * cd * A * nrm * vay, 0.5 * rho * cl * A * vax * vax
    d0, d1 = aero(-u, -v), aero(-u + wx, -v + wy)
    # dstates are s-derivatives: time derivatives times dt/ds (the same in both evaluations)
    du = (wind["dstates"][4] - calm["dstates"][4]) / calm["dtds"]
    dv = (wind["dstates"][5] - calm["dstates"][5]) / calm["dtds"]
    dw = (wind["dstates"][7] - calm["dstates"][7]) / calm["dtds"] * 9.81          # the vertical equation is scaled by 1 / g0
    assert abs(m * du - (d1[0] - d0[0])) <= 2e-11 * max(1.0, abs(d1[0]))
    assert abs(m * dv - (d1[1] - d0[1])) <= 2e-11 * max(1.0, abs(d1[1]))
    assert abs(m * dw - (d1[2] - d0[2])) <= 2e-9 * max(1.0, abs(d1[2]))
    assert abs(d1[0] - d0[0]) > 100.0          # the wind is felt


def test_engine_energy_limited_run_is_a_kkt_point_of_the_node_wise_formulation():
    """f1_optimal_laptime_test.cpp:1057-1117 (engine energy <= 24 MJ): the stored run, mapped to the node-wise statement of the
    integral constraint (one accumulated-energy state per node), satisfies every row, holds 24 MJ at node 0 (EXPECT_NEAR 24, 1e-4,
    :1115), reproduces the stored lap time and is stationary with the stored multipliers -- which pins the integrand
    (limebeer2014f1.h:293, engine.hpp:43), its weights in the closing element (optimal_laptime.hpp:1409) and their derivatives."""
    from tests.helpers import augmented_point
    p, ex = load_golden("f1_catalunya_engine_energy")
    x, lam, zl, zu = augmented_point(p, ex)
    r = oracle_nlp(p).eval(x, lam, 1.0)
    g = r["g"].reshape(p.n_elements, p.ncpe)
    assert np.abs(g[:, equality_rows(p)]).max() < 1.0e-10
    assert abs(x.reshape(p.n_points, p.nv)[0, p.aug_pos] - 24.0) < 1.0e-4
    jr, jc, jv = r["jac"]
    JTl = np.zeros(p.n)
    np.add.at(JTl, jc, jv * lam[jr])
    kkt = r["grad"] + JTl - zl + zu
    assert np.abs(kkt).max() < 5.0e-10, np.abs(kkt).max()
    p0 = p._copy(dissipation=[0.0, 0.0])
    assert abs(oracle_nlp(p0).eval(x, None, 1.0, derivs=False)["f"] - float(ex["laptime"])) < 1.0e-6
