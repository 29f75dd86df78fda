"""Host-only part of the fastest_lap.py-style surface (no GPU): the table of named objects, vehicle files, track data."""
import os
import numpy as np
import pytest

from fastest_lap_b200 import api, workloads
from fastest_lap_b200.track import track_from_npz
from fastest_lap_b200.vehicle import vehicle_from_xml


def test_vehicle_table_and_files(tmp_path, capsys):
    api.delete_variable("*")
    api.create_vehicle_empty("car", "f1-3dof")
    assert api.variable_type("car") == "f1-3dof"
    with pytest.raises(api.FastestLapError):
        api.create_vehicle_empty("car", "f1-3dof")              # the name is taken
    with pytest.raises(api.FastestLapError):
        api.create_vehicle_empty("other", "truck")
    ref = workloads._vehicles()["limebeer_2014_f1"]
    for path, value in zip(api._get("car").paths, ref):
        api.vehicle_set_parameter("car", "vehicle/" + path, value)
    api.vehicle_declare_new_constant_parameter("car", "vehicle/chassis/mass", "mass", 700.0)
    xml = str(tmp_path / "car.xml")
    api.vehicle_save_as_xml("car", xml)
    back = vehicle_from_xml(xml)
    assert back.model == api._get("car").model and back.get_parameter("vehicle/chassis/mass") == 700.0
    assert np.array_equal(np.delete(back.params, 0), np.delete(ref, 0))
    assert api.vehicle_type_get_sizes("f1-3dof") == (14, 4, 4) and api.vehicle_type_get_sizes("kart-6dof")[:2] == (13, 2)
    key, q, u, out = api.vehicle_type_get_names("kart-6dof")
    assert key == "road.arclength" and len(q) == 13 and len(u) == 2 and "laptime" in out
    api.print_variable("car")
    assert "<vehicle type=\"f1-3dof\">" in capsys.readouterr().out
    api.copy_variable("car", "car2"); api.move_variable("car2", "car3")
    assert api.variable_type("car3") == "f1-3dof"
    with pytest.raises(api.FastestLapError):
        api.download_scalar("car2")


def test_track_data():
    api.delete_variable("*")
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    api.create_vehicle_empty("car", "f1-3dof")
    api.vehicle_change_track("car", "catalunya")
    with pytest.raises(api.FastestLapError):
        api.vehicle_change_track("car", "monza")
    L = api.track_download_length("catalunya")
    assert 4600.0 < L < 4700.0
    x, y, xl, yl, xr, yr, th = api.track_coordinates("catalunya")
    nl, nr = api.track_download_data("catalunya", "distance-left-boundary"), api.track_download_data("catalunya", "distance-right-boundary")
    np.testing.assert_allclose(np.hypot(xl - x, yl - y), nl, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.hypot(xr - x, yr - y), nr, rtol=0, atol=1e-9)
    # the same through the C API's names (fastestlapc.cpp track_download_data: left.x, left.y, right.x, right.y)
    for key, ref in (("left.x", xl), ("left.y", yl), ("right.x", xr), ("right.y", yr)):
        np.testing.assert_array_equal(api.track_download_data("catalunya", key), ref)
    # the left boundary is to the left of the direction of travel
    assert np.all((xl - x) * -np.sin(th) + (yl - y) * np.cos(th) <= 0.0) or np.all((xl - x) * -np.sin(th) + (yl - y) * np.cos(th) >= 0.0)
    with pytest.raises(api.FastestLapError):
        api.track_download_data("catalunya", "banking")


def test_trajectory_on_a_track_with_elevation():
    """Time, x, y and psi exported with a solution (optimal_laptime.hpp:598-631) on Catalunya 3-D: the reference's stored
    trajectory (f1_optimal_laptime_catalunya_3d.xml, fixture by tests/golden/make_golden.py::trajectory_case) is reproduced
    from its stored states -- the flat-track formulas are 8 cm off in x, y on this track."""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f1_catalunya_3d_trajectory.npz"))
    q = d["inputs"]
    time, laptime, xyz, theta = api.trajectory(d["track_nodes"], d["centreline"], d["s"], float(d["track_length"]), True, q[:, 4], q[:, 5], q[:, 12], q[:, 13])
    assert abs(laptime - float(d["laptime"])) <= 1e-6
    assert np.max(np.abs(time - q[:, 11])) <= 1e-9
    assert np.max(np.abs(xyz[:, 0] - d["x_coord"])) <= 1e-9 and np.max(np.abs(xyz[:, 1] - d["y_coord"])) <= 1e-9
    assert np.max(np.abs(theta + q[:, 13] - d["psi"])) <= 1e-12
    yaw = d["track_nodes"][:, 0]
    assert np.max(np.abs(d["centreline"][:, 0] - q[:, 12] * np.sin(yaw) - d["x_coord"])) > 1e-2


def test_variable_parameter_follows_model_parameter():
    """Model_parameter::operator()(s) (src/core/foundation/model_parameter.h:51-77): upper_bound on the mesh points, the FIRST
    VALUE before the first point and the LAST VALUE after the last one (whatever the mesh indexes say), linear blend of the
    two indexed values in between; a mesh may reuse a value."""
    from fastest_lap_b200.vehicle import Vehicle, MODEL_F1
    veh = Vehicle(MODEL_F1, workloads._vehicles()["limebeer_2014_f1"].copy())
    values, idx, pts = [700.0, 650.0, 720.0], [1, 0, 2, 0], [100.0, 200.0, 400.0, 500.0]
    veh.declare_variable_parameter("vehicle/chassis/mass", values, idx, pts)
    s = np.array([0.0, 99.9, 100.0, 150.0, 200.0, 300.0, 399.9, 450.0, 500.0, 600.0])
    got = veh.params_at(s)[:, 0]
    # before the first point: values.front(); [100, 200): 650 -> 700; [200, 400): 700 -> 720; [400, 500): 720 -> 700; from 500: values.back()
    expect = np.array([700.0, 700.0, 650.0, 675.0, 700.0, 710.0, 700.0 + 20.0 * 199.9 / 200.0, 710.0, 720.0, 720.0])
    np.testing.assert_allclose(got, expect, rtol=0, atol=1e-12)
    other = veh.params_at(s)
    assert np.array_equal(other[:, 1:], np.tile(veh.params[1:], (len(s), 1)))          # every other parameter is constant
    with pytest.raises(KeyError):
        veh.declare_variable_parameter("vehicle/chassis/wings", values, idx, pts)
    api.delete_variable("*")
    api._table["car"] = veh.copy()
    api.vehicle_declare_new_variable_parameter("car", "vehicle/chassis/aerodynamics/cd", "cd_a;cd_b", [0.9, 1.1], [0, 1], [0.0, 1000.0])
    assert api._get("car").has_variable_parameters() and abs(api._get("car").params_at(np.array([500.0]))[0, 6] - 1.0) <= 1e-15
    with pytest.raises(api.FastestLapError):
        api.vehicle_declare_new_variable_parameter("car", "vehicle/chassis/cd", "a", [1.0], [0], [0.0])


def test_wind_and_engine_curve_in_the_vehicle_file(tmp_path):
    """Optional parameters outside the kernels' parameter vector: the wind is read (default 0) and settable by its reference
    path; an engine given as a power curve is refused by name (lion's sPolynomial is not in the reference tree), never dropped."""
    from fastest_lap_b200.vehicle import Vehicle, vehicle_from_xml, MODEL_F1
    veh = Vehicle(MODEL_F1, workloads._vehicles()["limebeer_2014_f1"].copy())
    assert veh.wind == [0.0, 0.0]
    veh.set_parameter("vehicle/chassis/aerodynamics/wind_velocity/northward", 5.0)
    veh.set_parameter("vehicle/chassis/aerodynamics/wind_velocity/eastward", -3.0)
    back = vehicle_from_xml(veh.to_xml())
    assert back.wind == [5.0, -3.0] and back.get_parameter("vehicle/chassis/aerodynamics/wind_velocity/eastward") == -3.0
    assert np.array_equal(back.params, veh.params)
    xml = veh.to_xml().replace("<maximum-power>", "<rpm-data>1000 2000</rpm-data><maximum-power>", 1)
    with pytest.raises(ValueError, match="power curves"):
        vehicle_from_xml(xml)
    with pytest.raises(KeyError):
        veh.set_parameter("vehicle/chassis/aerodynamics/wind_velocity/upward", 1.0)


def test_sweep_bounds_follow_each_laps_vehicle():
    """The tyre-slip rows are bounded by +-max(lambda_max(0), lambda_max(m g0)) of each lap's own vehicle (limebeer2014f1.h:232-262):
    a sweep over mass / aerodynamics / power shares one set of bounds (None), a sweep that changes a tyre's lambda-max gets its own
    set per lap, equal to what a single-lap problem of that vehicle has."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.ipm import sweep_bounds
    from fastest_lap_b200.nlp import problem_bounds
    from fastest_lap_b200.vehicle import MODEL_F1, F1_PARAM_PATHS
    p = workloads.f1_catalunya(60)
    P = workloads.sweep_parameters(MODEL_F1, 5)
    assert sweep_bounds(p, P) is None
    k = F1_PARAM_PATHS.index("rear-tire/lambda-max-1")
    P[3, k] *= 1.2
    b = sweep_bounds(p, P)
    assert b is not None and b[2].shape == (5, p.m)
    single = problem_bounds(p._copy(params=P[3]))
    np.testing.assert_array_equal(b[2][3], single[2]); np.testing.assert_array_equal(b[3][3], single[3])
    np.testing.assert_array_equal(b[2][0], problem_bounds(p._copy(params=P[0]))[2])
    assert np.any(b[3][3] != b[3][0])
