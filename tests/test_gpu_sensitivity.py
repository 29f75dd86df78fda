"""Parameter sensitivities of a converged lap (the reference's Optimal_laptime with check_optimality: dx_dp, dlaptime_dp,
src/core/applications/optimal_laptime.hpp:564-588, 651-729), checked the way the reference checks its own
(src/test/applications/f1_sensitivity_analysis.cpp:20-26): against the difference of laps re-solved with the parameter moved."""
import os
import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _moved(p, col, delta):
    q = p.params.copy()
    q[col] += delta
    return p._copy(params=q)


def test_mass_sensitivity_matches_resolved_laps():
    """F1, Catalunya, 500 nodes (the reference's Catalunya_discrete case): d/d(mass) of every NLP variable and of the lap time
    from ONE linear solve with the KKT matrix of the solution equals the central difference of two laps re-solved with the mass
    moved by +-0.05 kg.  Tolerances: the reference's -- 1e-6 on the lap time, 1e-5 .. 7e-5 (absolute) on states and controls."""
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.ipm import LapSolver, lap_time, lap_sensitivities, dlaptime_dp
    from fastest_lap_b200.vehicle import F1_PARAM_PATHS
    p, ex = load_golden("f1_catalunya_discrete")
    col = F1_PARAM_PATHS.index("chassis/mass")
    nlp = CollocationNLP(p)
    solver = LapSolver(nlp)
    base = solver.solve(ex["x"], warm=(ex["lam"], ex["zl"], ex["zu"]), warm_mu=1e-6)
    assert base["status"][0] == 1
    sigma = solver.sigma()
    solver.close(); nlp.close()
    x, lam = base["x"][0], base["y"][0]
    h = 1.0e-6 * p.params[col]
    dx, dy = lap_sensitivities(p, x, lam, sigma, [(_moved(p, col, h), _moved(p, col, -h), 2.0 * h)])
    dT, _ = dlaptime_dp(p, x, dx[0])

    dm = 0.05
    laps = {}
    for sign in (1.0, -1.0):
        q = _moved(p, col, sign * dm)
        n2 = CollocationNLP(q)
        s2 = LapSolver(n2)
        o = s2.solve(x, warm=(lam, base["zl"][0], base["zu"][0]), warm_mu=1e-6)
        assert o["status"][0] == 1
        laps[sign] = (o["x"][0].copy(), lap_time(q, o["x"][0], o["obj"][0]))
        s2.close(); n2.close()
    fd_x = (laps[1.0][0] - laps[-1.0][0]) / (2.0 * dm)
    fd_T = (laps[1.0][1] - laps[-1.0][1]) / (2.0 * dm)
    print("d laptime / d mass: %.9f s/kg (linear solve)  %.9f (re-solved laps)" % (dT, fd_T))
    print("max |dx/dp - difference|: %.3e, max |dx/dp| %.3e" % (np.max(np.abs(dx[0] - fd_x)), np.max(np.abs(dx[0]))))
    assert abs(dT - fd_T) <= 1.0e-6
    assert dT > 0.0                                           # a heavier car is slower
    assert np.max(np.abs(dx[0] - fd_x)) <= 1.0e-5
    t0 = lap_time(p, x, base["obj"][0])
    assert abs(t0 - float(ex["laptime"])) <= 1e-6 * t0          # the analysis does not move the solution


def test_sensitivities_through_the_api(tmp_path):
    """`<compute_sensitivity>` of optimal_laptime (fastestlapc.cpp:1254, 1525, 1585-1615, 1687-1706): parameters declared with
    vehicle_declare_new_constant_parameter get `<prefix>derivatives/laptime/<alias>` and `<prefix>derivatives/<variable>/<alias>`."""
    from fastest_lap_b200 import api as fl, workloads
    from fastest_lap_b200.track import track_from_npz
    from tests.test_gpu_api import _write_xml_fixtures
    paths = _write_xml_fixtures(tmp_path)
    fl.delete_variable("*")
    fl.create_vehicle_from_xml("car", paths["f1"])
    fl._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    fl.vehicle_declare_new_constant_parameter("car", "vehicle/chassis/mass", "mass", 660.0)
    fl.vehicle_declare_new_constant_parameter("car", "vehicle/rear-axle/engine/maximum-power", "power", 735.499)
    s = np.linspace(0.0, fl.track_download_length("catalunya"), 501)[:-1]
    opts = "<options><compute_sensitivity>true</compute_sensitivity><output_variables><prefix>run/</prefix></output_variables></options>"
    fl.optimal_laptime("car", "catalunya", s, opts)
    dm, dp = fl.download_scalar("run/derivatives/laptime/mass"), fl.download_scalar("run/derivatives/laptime/power")
    print("d laptime / d mass %.6f s/kg, / d power %.6f s/kW" % (dm, dp))
    assert dm > 0.0 and dp < 0.0
    du = fl.download_vector("run/derivatives/chassis.velocity.x/mass")
    dt = fl.download_vector("run/derivatives/time/mass")
    assert du.shape == s.shape and dt.shape == s.shape
    assert dt[0] == 0.0 and np.all(np.diff(dt) > -1e-9)      # a heavier car loses time all around the lap
    fl.delete_variable("*")


def test_mass_sensitivity_of_an_open_section():
    """The reference's Catalunya_chicane case (f1_sensitivity_analysis.cpp:214-291): the open section of the adapted Catalunya mesh
    (wheel inertias 1.00 / 1.55), mass as the parameter; dx/dp and d(time of the last node)/dp against laps re-solved with the mass
    moved by +-0.01 kg (the reference moves it by 0.01 kg and asks 1e-6)."""
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.ipm import LapSolver, lap_sensitivities, dlaptime_dp
    from fastest_lap_b200.vehicle import F1_PARAM_PATHS
    from fastest_lap_b200.api import trajectory
    p, ex = load_golden("f1_catalunya_chicane")
    assert not p.closed
    col = F1_PARAM_PATHS.index("chassis/mass")

    def last_node_time(q, x):
        # time of the last node: trapezoid rule of dt/ds from the given first node (optimal_laptime.hpp:598-631)
        first = np.concatenate([np.delete(q.q0, 11), [q.u0[0], q.u0[2]]])
        X = np.vstack([first, x.reshape(-1, q.nv)])
        t, lap, _, _ = trajectory(q.track_nodes, np.zeros((q.n_points, 3)), q.s, q.track_length, False, X[:, 4], X[:, 5], X[:, 11], X[:, 12])
        return lap

    def solve(q, x0):
        nlp = CollocationNLP(q)
        solver = LapSolver(nlp)
        out = solver.solve(x0)
        assert out["status"][0] == 1
        sig = solver.sigma()
        solver.close(); nlp.close()
        return out, sig

    base, sigma = solve(p, ex["x"])
    x, lam = base["x"][0], base["y"][0]
    h = 1.0e-6 * p.params[col]
    dx, _ = lap_sensitivities(p, x, lam, sigma, [(_moved(p, col, h), _moved(p, col, -h), 2.0 * h)])
    dT, d = dlaptime_dp(p, x, dx[0])
    assert d[0] == 0.0 and d.shape == (p.n_points,)
    dm = 0.01
    xp, _ = solve(_moved(p, col, dm), x)
    xm, _ = solve(_moved(p, col, -dm), x)
    fd_x = (xp["x"][0] - xm["x"][0]) / (2.0 * dm)
    fd_T = (last_node_time(_moved(p, col, dm), xp["x"][0]) - last_node_time(_moved(p, col, -dm), xm["x"][0])) / (2.0 * dm)
    print("open section: d time / d mass %.9f s/kg (linear solve) %.9f (re-solved); max |dx/dp - difference| %.2e of %.2e" % (
        dT, fd_T, np.max(np.abs(dx[0] - fd_x)), np.max(np.abs(dx[0]))))
    assert abs(dT - fd_T) <= 1.0e-6
    assert np.max(np.abs(dx[0] - fd_x)) <= 1.0e-5


def test_sensitivity_to_the_values_of_a_parameter_that_varies_along_the_lap(tmp_path):
    """A mass declared as a piecewise-linear function of the arclength (vehicle_declare_new_variable_parameter, aliases "m0;m1"): the
    lap time's derivative with respect to each of its two values, against laps re-solved with that value moved."""
    from fastest_lap_b200 import api as fl, workloads
    from fastest_lap_b200.track import track_from_npz
    from tests.test_gpu_api import _write_xml_fixtures
    paths = _write_xml_fixtures(tmp_path)
    fl.delete_variable("*")
    fl.create_vehicle_from_xml("car", paths["f1"])
    fl._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    L = fl.track_download_length("catalunya")
    s = np.linspace(0.0, L, 501)[:-1]

    def declare(m0, m1):
        fl.vehicle_declare_new_variable_parameter("car", "vehicle/chassis/mass", "m0; m1", [m0, m1], [0, 1], [0.0, 0.9 * L])
    declare(660.0, 670.0)
    fl.optimal_laptime("car", "catalunya", s, "<options><compute_sensitivity>true</compute_sensitivity><output_variables><prefix>run/</prefix></output_variables></options>")
    d0, d1 = fl.download_scalar("run/derivatives/laptime/m0"), fl.download_scalar("run/derivatives/laptime/m1")
    fd = []
    for k in range(2):
        t = []
        for sign in (1.0, -1.0):
            m = [660.0, 670.0]; m[k] += sign * 0.05
            declare(*m)
            fl.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>moved/</prefix></output_variables></options>")
            t.append(fl.download_scalar("moved/laptime"))
        fd.append((t[0] - t[1]) / 0.1)
    print("d laptime / d m0 %.8f (re-solved %.8f), / d m1 %.8f (re-solved %.8f) s/kg" % (d0, fd[0], d1, fd[1]))
    assert d0 > 0.0 and d1 > 0.0
    assert abs(d0 - fd[0]) <= 1e-6 and abs(d1 - fd[1]) <= 1e-6
    # the two values share the lap: together they weigh what a constant mass does (0.0395 s/kg on this lap)
    assert abs(d0 + d1 - 0.0395) <= 2e-3
    fl.delete_variable("*")
