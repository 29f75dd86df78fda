"""The reference's Python surface (src/main/python/fastest_lap.py) on the GPU path: the README example of the reference --
limebeer-2014-f1 on Catalunya, 500 points -- and the kart G-G diagram."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _write_xml_fixtures(tmp_path):
    """Vehicle / track XML files rebuilt from the packaged data (the reference's database is not on the GPU box)."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.vehicle import F1_PARAM_PATHS, KART_PARAM_PATHS
    from fastest_lap_b200.track import track_from_npz
    import xml.etree.ElementTree as ET
    veh = workloads._vehicles()
    paths = {}
    for name, typ, plist, key in (("f1", "f1-3dof", F1_PARAM_PATHS, "limebeer_2014_f1"), ("kart", "kart-6dof", KART_PARAM_PATHS, "roberto_lot_kart_2016")):
        root = ET.Element("vehicle", type=typ)
        for pth, val in zip(plist, veh[key]):
            node = root
            for part in pth.split("/"):
                nxt = node.find(part)
                node = ET.SubElement(node, part) if nxt is None else nxt
            node.text = repr(float(val))
        paths[name] = str(tmp_path / (name + ".xml"))
        ET.ElementTree(root).write(paths[name])
    return paths


def test_optimal_laptime_readme_example(tmp_path):
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    L = api.track_download_length("catalunya")
    s = np.linspace(0.0, L, 501)[:-1]
    prefix, names = api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>run/</prefix></output_variables></options>")
    assert prefix == "run/" and "chassis.velocity.x" in names and "laptime" in names
    gold = np.load(os.path.join(HERE, "golden", "f1_catalunya_discrete.npz"))
    assert abs(api.download_scalar("run/laptime") - float(gold["laptime"])) <= 1e-6 * float(gold["laptime"])
    u = api.download_vector("run/chassis.velocity.x")
    assert u.shape == (500,) and np.max(np.abs(u - gold["x"].reshape(500, 15)[:, 4])) <= 1e-4
    t = api.download_vector("run/time")
    assert np.max(np.abs(t - gold["time"])) <= 1e-5
    with pytest.raises(api.FastestLapError):
        api.create_vehicle_from_xml("car", paths["f1"])          # name already taken
    api.delete_variable("run/*")
    with pytest.raises(api.FastestLapError):
        api.download_scalar("run/laptime")


def test_gg_diagram_of_the_kart(tmp_path):
    from fastest_lap_b200 import api
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("kart", paths["kart"])
    ay, ay_minus, ax_max, ax_min = api.gg_diagram("kart", 50.0 / 3.6, 12)
    g = np.load(os.path.join(HERE, "golden", "kart_steady_state.npz"))
    i = list(g["names"]).index("velocity_50/max_lat_acc")
    assert abs(ay[-1] * 9.81 - g["x"][i, 7]) <= 1e-6 * g["x"][i, 7]
    assert len(ay) == 12 and ay[0] == 0.0 and all(a >= b - 1e-9 for a, b in zip(ax_max, ax_min))
    assert ay_minus[3] == -ay[3]


def test_optimal_laptime_of_the_kart(tmp_path):
    """kart-6dof on Vendrell, 250 points, the C API's defaults for karts (derivative form; start from the GPU steady state)."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("kart", paths["kart"])
    api._table["vendrell"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "vendrell.npz"), "vendrell")
    L = api.track_download_length("vendrell")
    s = np.linspace(0.0, L, 251)[:-1]
    prefix, names = api.optimal_laptime("kart", "vendrell", s, "<options><output_variables><prefix>k/</prefix></output_variables></options>")
    t = api.download_scalar("k/laptime")
    assert 55.0 < t < 70.0
    u = api.download_vector("k/chassis.velocity.x")
    assert u.shape == (250,) and u.min() > 5.0 and u.max() < 40.0
    time = api.download_vector("k/time")
    assert np.all(np.diff(time) > 0.0) and abs(time[-1] + (t - time[-1]) - t) < 1e-9


def test_warm_start_parameter_sweep(tmp_path):
    """The reference's use of <save_warm_start>/<warm_start> (fastestlapc.cpp:1561-1568, 1712-1713): solve once, change a
    vehicle parameter, restart from the saved primal-dual solution.  The warm run agrees with a cold run of the modified
    vehicle and takes fewer iterations."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api._warm_start.clear()
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    s = np.linspace(0.0, api.track_download_length("catalunya"), 501)[:-1]
    with pytest.raises(api.FastestLapError):
        api.optimal_laptime("car", "catalunya", s, "<options><warm_start>true</warm_start></options>")      # nothing saved yet
    api.optimal_laptime("car", "catalunya", s, "<options><save_warm_start>true</save_warm_start><output_variables><prefix>base/</prefix></output_variables></options>")
    api.vehicle_set_parameter("car", "vehicle/chassis/mass", 680.0)
    api.optimal_laptime("car", "catalunya", s, "<options><warm_start>true</warm_start><output_variables><prefix>warm/</prefix></output_variables></options>")
    api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>cold/</prefix></output_variables></options>")
    t0, tw, tc = (api.download_scalar(k + "/laptime") for k in ("base", "warm", "cold"))
    assert tw > t0 + 1e-3                         # 20 kg heavier is slower
    assert abs(tw - tc) <= 2e-5 * tc              # two converged runs from different starts: neighbouring local minima, 4e-4 s apart
    iw, ic = (api.download_scalar(k + "/optimization_data.number_of_iterations") for k in ("warm", "cold"))
    assert iw < ic


def test_open_section_with_initial_condition_and_solution_file(tmp_path):
    """<closed_simulation> false with <initial_condition> from the tables (fastestlapc.cpp:1237-1249): nodes 0..99 of the
    500-node lap, started from the closed lap's state at node 0 replicated.  The first node stays fixed; with its end
    free the section cannot be slower than the same stretch of the closed lap; <write_xml> writes the reference's solution
    format, which the loader reads back to the same vectors."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    from fastest_lap_b200.solution import solution_from_xml, names_of
    from fastest_lap_b200.vehicle import MODEL_F1
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    s = np.linspace(0.0, api.track_download_length("catalunya"), 501)[:-1]
    api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>lap/</prefix></output_variables></options>")
    inputs, controls = names_of(MODEL_F1)
    i0, i1 = 0, 100          # the main straight and the first corners
    api.create_vector("q_start", [api.download_vector("lap/" + nm)[i0] for nm in inputs])
    api.create_vector("u_start", [api.download_vector("lap/" + nm)[i0] for nm in controls])
    xml = str(tmp_path / "section.xml")
    opts = ("<options><closed_simulation>false</closed_simulation><initial_condition><q from_table=\"q_start\"/><u from_table=\"u_start\"/></initial_condition>"
            "<write_xml>true</write_xml><xml_file_name>%s</xml_file_name><output_variables><prefix>open/</prefix></output_variables></options>" % xml)
    with pytest.raises(api.FastestLapError):
        api.optimal_laptime("car", "catalunya", s[i0:i1], "<options><closed_simulation>false</closed_simulation></options>")
    api.optimal_laptime("car", "catalunya", s[i0:i1], opts)
    t_lap = api.download_vector("lap/time")
    u = api.download_vector("open/chassis.velocity.x")
    t = api.download_vector("open/time")
    assert u.shape == (i1 - i0,) and u[0] == api.download_vector("lap/chassis.velocity.x")[i0] and t[0] == t_lap[i0]
    assert api.download_scalar("open/laptime") == t[-1]
    assert t[-1] - t[0] <= t_lap[i1 - 1] - t_lap[i0] + 1e-6
    assert t[-1] - t[0] >= 0.9 * (t_lap[i1 - 1] - t_lap[i0])
    # the first stretch is the closed lap's (the free end only changes the approach to the last corner)
    assert np.max(np.abs(u[:20] - api.download_vector("lap/chassis.velocity.x")[i0:i0 + 20])) <= 0.5
    sol = solution_from_xml(xml, MODEL_F1)
    assert not sol.is_closed and sol.is_direct and sol.n_points == i1 - i0
    assert np.array_equal(sol.inputs[:, inputs.index("chassis.velocity.x")], u) and np.array_equal(sol.inputs[:, inputs.index("time")], t)
    assert np.array_equal(sol.controls[2].values, api.download_vector("open/chassis.throttle")) and sol.controls[3].kind == "dont optimize"
    assert abs(sol.laptime - t[-1]) <= 1e-6
    assert sol.lam.size == (i1 - i0 - 1) * 19 and sol.zl.size == (i1 - i0 - 1) * 15 and sol.pack_x().size == sol.zl.size


def test_variable_bounds_option(tmp_path):
    """<variable_bounds> (fastestlapc.cpp:1308-1376): a speed limit of 70 m/s is respected and costs lap time; malformed
    entries and unknown names raise the reference's errors."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    s = np.linspace(0.0, api.track_download_length("catalunya"), 501)[:-1]
    api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>free/</prefix></output_variables></options>")
    api.optimal_laptime("car", "catalunya", s, "<options><variable_bounds><chassis.velocity.x><upper>70.0</upper></chassis.velocity.x></variable_bounds>"
                                               "<output_variables><prefix>limited/</prefix></output_variables></options>")
    assert api.download_vector("free/chassis.velocity.x").max() > 80.0
    assert api.download_vector("limited/chassis.velocity.x").max() <= 70.0 + 1e-6
    assert api.download_scalar("limited/laptime") > api.download_scalar("free/laptime") + 0.5
    with pytest.raises(api.FastestLapError, match="shall have only children"):
        api.optimal_laptime("car", "catalunya", s, "<options><variable_bounds><chassis.velocity.x><max>70.0</max></chassis.velocity.x></variable_bounds></options>")
    with pytest.raises(api.FastestLapError, match="not found"):
        api.optimal_laptime("car", "catalunya", s, "<options><variable_bounds><chassis.speed><upper>70.0</upper></chassis.speed></variable_bounds></options>")


def test_hypermesh_control_and_integral_constraint_options(tmp_path):
    """<control_variables><chassis.brake-bias optimal_control_type="hypermesh"> and <integral_constraints> (fastestlapc.cpp:1255-1306):
    the brake bias comes back constant inside each stretch and the lap is not slower than with the fixed bias; with the engine
    energy restricted to 26 MJ the integral ends at its bound and the lap is slower; what is not offered raises by name."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    s = np.linspace(0.0, api.track_download_length("catalunya"), 501)[:-1]
    api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>free/</prefix></output_variables></options>")
    api.optimal_laptime("car", "catalunya", s, "<options><control_variables><chassis.brake-bias optimal_control_type=\"hypermesh\"><hypermesh>0.0, 2000.0</hypermesh>"
                                               "</chassis.brake-bias></control_variables><output_variables><prefix>hm/</prefix></output_variables></options>")
    bb = api.download_vector("hm/chassis.brake-bias")
    first = s < 2000.0
    assert np.ptp(bb[first]) <= 1e-9 and np.ptp(bb[~first]) <= 1e-9 and 0.0 < bb.min() and bb.max() < 1.0
    assert api.download_scalar("hm/laptime") <= api.download_scalar("free/laptime") + 1e-6
    # "Engine energy limits require a higher value of the smooth throttle coefficient" (f1_optimal_laptime_test.cpp:1068-1069)
    api.vehicle_set_parameter("car", "vehicle/rear-axle/smooth_throttle_coeff", 1.0e-2)
    _, names = api.optimal_laptime("car", "catalunya", s, "<options><integral_constraints><engine-energy><lower_bound>0.0</lower_bound><upper_bound>26.0</upper_bound>"
                                                          "</engine-energy></integral_constraints><output_variables><prefix>en/</prefix></output_variables></options>")
    assert "integral_quantities.engine-energy" in names
    assert abs(api.download_scalar("en/integral_quantities.engine-energy") - 26.0) <= 1e-4
    assert api.download_scalar("en/laptime") > api.download_scalar("free/laptime") + 0.1
    assert api.download_vector("en/chassis.velocity.x").shape == (500,)
    with pytest.raises(api.FastestLapError, match="one hypermesh / constant control or one integral constraint"):
        api.optimal_laptime("car", "catalunya", s, "<options><integral_constraints><tire-fl-energy><lower_bound>0</lower_bound><upper_bound>0.8</upper_bound></tire-fl-energy>"
                                                   "<tire-fr-energy><lower_bound>0</lower_bound><upper_bound>0.8</upper_bound></tire-fr-energy></integral_constraints></options>")
    with pytest.raises(api.FastestLapError, match="boost-time"):
        api.optimal_laptime("car", "catalunya", s, "<options><integral_constraints><boost-time><lower_bound>0</lower_bound><upper_bound>10</upper_bound></boost-time></integral_constraints></options>")
    with pytest.raises(api.FastestLapError, match="only the F1's chassis.brake-bias"):
        api.optimal_laptime("car", "catalunya", s, "<options><control_variables><chassis.throttle optimal_control_type=\"constant\"/></control_variables></options>")


def test_output_channels_through_the_api(tmp_path):
    """<output_variables><variables> with channels of the model (fastestlapc.cpp:1215-1233, 1643-1690) and vehicle_get_output
    (fastestlapc.h:72): tyre dissipation integrated over the lap equals what the integral quantity of the same name accumulates;
    the channel of one node equals vehicle_get_output at that node's inputs and controls; unknown names raise."""
    from fastest_lap_b200 import api, workloads
    from fastest_lap_b200.track import track_from_npz
    paths = _write_xml_fixtures(tmp_path)
    api.delete_variable("*")
    api.create_vehicle_from_xml("car", paths["f1"])
    api._table["catalunya"] = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"), "catalunya")
    api.vehicle_change_track("car", "catalunya")
    s = np.linspace(0.0, api.track_download_length("catalunya"), 501)[:-1]
    want = ["rear-axle.left-tire.dissipation", "rear-axle.left-tire.true_kappa", "chassis.acceleration.x", "chassis.acceleration.y", "chassis.aerodynamics.drag.x"]
    api.optimal_laptime("car", "catalunya", s, "<options><output_variables><prefix>run/</prefix><variables>%s</variables></output_variables></options>" %
                        "".join("<%s/>" % w for w in want))
    ax, ay = api.download_vector("run/chassis.acceleration.x"), api.download_vector("run/chassis.acceleration.y")
    assert ax.shape == (500,) and -60.0 < ax.min() < -20.0 and 5.0 < ax.max() < 25.0 and np.abs(ay).max() > 20.0      # an F1 car: ~4 g braking, >2 g lateral
    assert api.download_vector("run/chassis.aerodynamics.drag.x").max() < 0.0                                       # drag opposes the motion
    names_in, names_u = api.vehicle_type_get_names("f1-3dof")[1:3]
    k = 137
    q = np.array([api.download_vector("run/" + nm)[k] for nm in names_in])
    u = np.array([api.download_vector("run/" + nm)[k] for nm in names_u])
    for w in want:
        assert abs(api.vehicle_get_output("car", q, u, s[k], w) - api.download_vector("run/" + w)[k]) <= 1e-9 * max(1.0, abs(api.download_vector("run/" + w)[k])), w
    with pytest.raises(api.FastestLapError, match="is not offered"):
        api.optimal_laptime("car", "catalunya", s, "<options><output_variables><variables><chassis.banana/></variables></output_variables></options>")
    with pytest.raises(api.FastestLapError, match="is not offered"):
        api.vehicle_get_output("car", q, u, s[k], "chassis.banana")
