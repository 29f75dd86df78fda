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
