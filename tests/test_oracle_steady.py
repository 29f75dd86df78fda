"""Steady-state (G-G) NLP of the kart: the oracle's restatement of lot2016kart::steady_state_equations against the
reference's converged steady states (src/test/applications/data/steady_state.xml -> tests/golden/kart_steady_state.npz)."""
import os
import numpy as np

from oracle import oracle as orc

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kart_steady_state.npz"))


def test_golden_steady_states_satisfy_the_equations():
    worst = 0.0
    for name, v, x in zip(GOLD["names"], GOLD["v"], GOLD["x"]):
        r = orc.kart_steady(x, GOLD["params"], v, ax=x[6], ay=x[7])
        worst = max(worst, np.max(np.abs(r["c"][:6])))
        # tyre-slip inequality rows within their bounds (lot2016kart.h:117-121), with the solver's tolerance
        assert np.all(np.abs(r["c"][6:8]) <= 0.11 + 1e-6) and np.all(np.abs(r["c"][8:]) <= 0.09 + 1e-6), name
    assert worst <= 1e-9            # steady_state_ad_test's tolerance class (Ipopt tol 1e-8)


def test_free_and_fixed_accelerations_agree():
    x, v = GOLD["x"][3], GOLD["v"][3]
    a = orc.kart_steady(x, GOLD["params"], v, ax=x[6], ay=x[7])
    b = orc.kart_steady(x, GOLD["params"], v, free_ax=True, free_ay=True, cax=-1.0)
    assert np.allclose(a["c"], b["c"], rtol=0, atol=1e-14)
    assert b["f"] == -x[6] and np.all(a["jac"][:, 6:] == 0.0) and np.any(b["jac"][:, 6] != 0.0)


def test_derivatives_match_finite_differences():
    rng = np.random.default_rng(0)
    x, v = GOLD["x"][5] * (1 + 1e-3 * rng.standard_normal(8)), GOLD["v"][5]
    lam = rng.standard_normal(12)
    r = orc.kart_steady(x, GOLD["params"], v, free_ax=True, free_ay=True, cax=-1.0, cay=0.5, lam=lam)
    for k in range(8):
        h = 1e-6 * max(1.0, abs(x[k]))
        xp, xm = x.copy(), x.copy(); xp[k] += h; xm[k] -= h
        rp = orc.kart_steady(xp, GOLD["params"], v, free_ax=True, free_ay=True, cax=-1.0, cay=0.5, lam=lam)
        rm = orc.kart_steady(xm, GOLD["params"], v, free_ax=True, free_ay=True, cax=-1.0, cay=0.5, lam=lam)
        dj = (rp["c"] - rm["c"]) / (2 * h)
        assert np.max(np.abs(dj - r["jac"][:, k]) / np.maximum(1.0, np.abs(dj))) <= 1e-6
        gl = lambda q: q["grad"] + q["jac"].T @ lam
        dh = (gl(rp) - gl(rm)) / (2 * h)
        assert np.max(np.abs(dh - r["hess"][:, k]) / np.maximum(1.0, np.abs(dh))) <= 1e-5


def test_f1_steady_state_oracle_reproduces_the_reference_maximum_lateral_acceleration():
    """limebeer2014f1::steady_state_equations restated (oracle/steady.hpp) + scipy's trust-constr: the maximum lateral
    acceleration at two speeds of f1_steady_state_test.xml (f1_steady_state_test.cpp:47-80, tolerance 2e-4)."""
    import scipy.optimize as so
    import warnings
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f1_steady_state.npz"))
    P = g["params"]
    lb = np.array([-1.15, -1.15, -1.15, -1.15, 0, -2, -2, -2, -2, 0, -1, -8.0, -8.0])
    ub = np.array([0.25, 0.25, 1, 1, 1.5, 0.1, 0.1, 0.1, 0.1, 1.5, 1, 8.0, 8.0])
    cl, cu = np.array([0.0] * 11 + [-1.0] * 4), np.array([0.0] * 11 + [1.2] * 4)

    def solve(v, x0, free, cay):
        F = lambda x: orc.f1_steady(x, P, v, 0.0, 0.0, free, free, 0.0, cay)
        cons = [so.NonlinearConstraint(lambda x: F(x)["c"], cl, cu, jac=lambda x: F(x)["jac"])]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return so.minimize(lambda x: F(x)["f"], x0, jac=lambda x: F(x)["grad"], method="trust-constr", constraints=cons,
                               bounds=so.Bounds(lb, ub), options=dict(xtol=1e-12, gtol=1e-10, maxiter=2000)).x
    x0 = np.array([0, 0, 0, 0, 0, -0.25, -0.25, -0.25, -0.25, 0, 0, 0, 0.0])
    for i in (0, 50):
        x = solve(g["v"][i], solve(g["v"][i], x0, False, 0.0), True, -1.0)
        assert abs(x[12] * 9.81 - g["ay_max"][i]) <= 2e-4 and abs(x[11] * 9.81 - g["ax_at_ay_max"][i]) <= 2e-4
