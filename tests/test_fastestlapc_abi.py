"""The reference's C API (include/fastestlapc.h <- src/main/c/fastestlapc.h:28-107) exported by libfastestlap_b200.so: bound
with ctypes exactly as the reference's Python module binds libfastestlapc (src/main/python/fastest_lap.py:10-14 and the
wrappers below it: c_char_p names, c_double arrays), host-side entries on the CPU, the README example on the GPU."""
import ctypes as c
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fastest_lap_b200 import build
    L = c.CDLL(build.build())
    # fastest_lap.py:12-14
    L.download_scalar.restype = c.c_double
    L.track_download_length.restype = c.c_double
    L.vehicle_type_get_sizes.argtypes = [c.POINTER(c.c_int), c.POINTER(c.c_int), c.POINTER(c.c_int), c.c_char_p]
    L.fastestlapc_last_error.restype = c.c_char_p
    L.delete_variable(b"*")
    return L


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    """Vehicle / track database files rebuilt from the packaged data (the reference's database is not on the GPU box)."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.vehicle import Vehicle, MODEL_F1, MODEL_KART
    from fastest_lap_b200.track import track_from_npz, track_to_xml
    d = tmp_path_factory.mktemp("db")
    veh = workloads._vehicles()
    out = {}
    for name, model, key in (("f1", MODEL_F1, "limebeer_2014_f1"), ("kart", MODEL_KART, "roberto_lot_kart_2016")):
        out[name] = str(d / (name + ".xml"))
        open(out[name], "w").write(Vehicle(model, veh[key]).to_xml())
    data = os.path.join(os.path.dirname(workloads.__file__), "data")
    for name in ("catalunya", "vendrell"):
        out[name] = str(d / (name + ".xml"))
        open(out[name], "w").write(track_to_xml(track_from_npz(os.path.join(data, name + ".npz"))))
    return out


def b(s):
    return c.c_char_p(s.encode("utf-8"))


def err(lib):
    return lib.fastestlapc_last_error().decode()


def test_library_exports_the_reference_c_api(lib):
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fastestlapc.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b([a-z_]+)\s*\(", text)) - {"defined"})
    expected = {"set_print_level", "print_variables", "print_variable", "print_variable_to_string", "create_vehicle_from_xml", "create_vehicle_empty",
                "create_track_from_xml", "create_vector", "create_scalar", "copy_variable", "move_variable", "delete_variable", "variable_type",
                "download_scalar", "download_vector_size", "download_vector", "vehicle_type_get_sizes", "vehicle_type_get_names", "vehicle_get_output",
                "vehicle_save_as_xml", "track_download_number_of_points", "track_download_data", "track_download_length", "vehicle_set_parameter",
                "vehicle_declare_new_constant_parameter", "vehicle_declare_new_variable_parameter", "vehicle_change_track",
                "track_set_track_limit_correction", "steady_state", "propagate_vehicle", "gg_diagram", "optimal_laptime", "circuit_preprocessor"}
    assert expected <= set(names), expected - set(names)          # the 33 functions of fastestlapc.h:28-107
    for n in names:
        assert hasattr(lib, n), n


def test_tables_and_loaders_on_the_host(lib, files, tmp_path):
    lib.delete_variable(b"*")
    # scalars and vectors
    lib.create_scalar(b("a/k"), c.c_double(2.5))
    v = (c.c_double * 3)(1.0, 2.0, 3.0)
    lib.create_vector(b("a/v"), c.c_int(3), v)
    assert lib.download_scalar(b("a/k")) == 2.5 and lib.download_vector_size(b("a/v")) == 3
    back = (c.c_double * 3)()
    lib.download_vector(back, c.c_int(3), b("a/v"))
    assert list(back) == [1.0, 2.0, 3.0]
    lib.create_scalar(b("a/k"), c.c_double(1.0))
    assert "already exists" in err(lib)
    lib.copy_variable(b("a/v"), b("b/v")); lib.move_variable(b("a/k"), b("b/k"))
    assert lib.download_scalar(b("b/k")) == 2.5 and lib.download_vector_size(b("b/v")) == 3
    lib.download_scalar(b("a/k"))
    assert "does not exist" in err(lib)
    lib.delete_variable(b("b/*"))
    lib.download_vector_size(b("b/v"))
    assert "does not exist" in err(lib)
    # vehicles
    lib.create_vehicle_from_xml(b("car"), b(files["f1"]))
    assert err(lib) == ""
    buf = c.create_string_buffer(64)
    lib.variable_type(buf, c.c_int(64), b("car"))
    assert buf.value == b"f1-3dof"
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/mass"), c.c_double(700.0))
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/wings"), c.c_double(1.0))
    assert "was not found" in err(lib)
    out = str(tmp_path / "car.xml")
    lib.vehicle_save_as_xml(b("car"), b(out))
    from fastest_lap_b200.vehicle import vehicle_from_xml
    from fastest_lap_b200 import workloads
    again = vehicle_from_xml(out)
    ref = workloads._vehicles()["limebeer_2014_f1"].copy(); ref[0] = 700.0
    assert np.array_equal(again.params, ref)
    ni, nc, no = c.c_int(), c.c_int(), c.c_int()
    lib.vehicle_type_get_sizes(c.byref(ni), c.byref(nc), c.byref(no), b("kart-6dof"))
    assert (ni.value, nc.value, no.value) == (13, 2, 4)
    vals, idx, pts = (c.c_double * 2)(700.0, 640.0), (c.c_int * 2)(0, 1), (c.c_double * 2)(0.0, 4000.0)
    lib.vehicle_declare_new_variable_parameter(b("car"), b("vehicle/chassis/mass"), b("m0;m1"), c.c_int(2), vals, c.c_int(2), idx, pts)
    assert err(lib) == ""
    # tracks
    lib.create_track_from_xml(b("catalunya"), b(files["catalunya"]))
    assert err(lib) == ""
    n = lib.track_download_number_of_points(b("catalunya"))
    from fastest_lap_b200.track import track_from_npz
    trk = track_from_npz(os.path.join(os.path.dirname(workloads.__file__), "data", "catalunya.npz"))
    assert n == trk.n_points and abs(lib.track_download_length(b("catalunya")) - trk.track_length) == 0.0
    data = (c.c_double * n)()
    for key, mine in (("arclength", trk.s_data), ("centerline.x", trk("x", trk.s_data)), ("heading-angle", trk("yaw", trk.s_data)),
                      ("distance-left-boundary", trk("nl", trk.s_data))):
        lib.track_download_data(data, b("catalunya"), c.c_int(n), b(key))
        assert np.array_equal(np.array(data), mine), key
    lib.track_download_data(data, b("catalunya"), c.c_int(n), b("left.x"))
    th, x, nl = trk("yaw", trk.s_data), trk("x", trk.s_data), trk("nl", trk.s_data)
    np.testing.assert_allclose(np.array(data), x + nl * np.sin(th), rtol=0, atol=1e-12)
    lib.track_download_data(data, b("catalunya"), c.c_int(n), b("banking"))
    assert "is not valid" in err(lib)
    lib.vehicle_change_track(b("car"), b("monza"))
    assert "does not exist" in err(lib)
    lib.circuit_preprocessor(b("<options/>"))
    assert "not part of the B200 path" in err(lib)
    lib.delete_variable(b"*")


@pytest.mark.gpu
def test_readme_example_through_the_c_api(lib, files):
    """The reference's README example (limebeer-2014-f1, Catalunya, 500 points) exactly as fastest_lap.py drives libfastestlapc:
    create_vehicle_from_xml, create_track_from_xml, optimal_laptime(vehicle, track, s, options), download_*: lap time
    77.791438 s (f1_optimal_laptime_catalunya_discrete.xml of the reference, Ixx = Iyy = 0 as in its test)."""
    lib.delete_variable(b"*")
    lib.create_vehicle_from_xml(b("car"), b(files["f1"]))
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/inertia/Ixx"), c.c_double(0.0))
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/inertia/Iyy"), c.c_double(0.0))
    lib.create_track_from_xml(b("catalunya"), b(files["catalunya"]))
    L = lib.track_download_length(b("catalunya"))
    s = np.linspace(0.0, L, 501)[:-1]
    c_s = (c.c_double * len(s))(*s)
    options = "<options><output_variables><prefix>run/</prefix></output_variables><print_level>0</print_level></options>"
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b(options))
    assert err(lib) == ""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "f1_catalunya_discrete.npz"))
    assert abs(lib.download_scalar(b("run/laptime")) - float(gold["laptime"])) <= 1e-6 * float(gold["laptime"])
    assert lib.download_vector_size(b("run/chassis.velocity.x")) == 500
    u = (c.c_double * 500)()
    lib.download_vector(u, c.c_int(500), b("run/chassis.velocity.x"))
    assert np.max(np.abs(np.array(u) - gold["x"].reshape(500, 15)[:, 4])) <= 1e-4
    t = (c.c_double * 500)()
    lib.download_vector(t, c.c_int(500), b("run/time"))
    assert np.max(np.abs(np.array(t) - gold["time"])) <= 1e-5
    # output channels of the model along the lap (<output_variables><variables>, fastestlapc.cpp:1215-1233, 1643-1690) and the same
    # channel at one node through vehicle_get_output (fastestlapc.h:72)
    lib.vehicle_change_track(b("car"), b("catalunya"))
    lib.vehicle_get_output.restype = c.c_double
    lib.vehicle_get_output.argtypes = [c.c_char_p, c.POINTER(c.c_double), c.POINTER(c.c_double), c.c_double, c.c_char_p]
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><output_variables><prefix>ch/</prefix><variables><chassis.acceleration.y/>"
                                                                          "<rear-axle.left-tire.dissipation/></variables></output_variables></options>"))
    assert err(lib) == ""
    ay = (c.c_double * 500)()
    lib.download_vector(ay, c.c_int(500), b("ch/chassis.acceleration.y"))
    assert np.abs(np.array(ay)).max() > 20.0
    names_in = ["front-axle.left-tire.kappa", "front-axle.right-tire.kappa", "rear-axle.left-tire.kappa", "rear-axle.right-tire.kappa", "chassis.velocity.x",
                "chassis.velocity.y", "chassis.omega.z", "chassis.force_z_fl_g", "chassis.force_z_fr_g", "chassis.force_z_rl_g", "chassis.force_z_rr_g", "time",
                "road.lateral-displacement", "road.track-heading-angle"]
    names_u = ["front-axle.steering-angle", "rear-axle.boost", "chassis.throttle", "chassis.brake-bias"]
    k = 211

    def column(name):
        v = (c.c_double * 500)()
        lib.download_vector(v, c.c_int(500), b("ch/" + name))
        return v[k]
    q = (c.c_double * 14)(*[column(nm) for nm in names_in])
    uu = (c.c_double * 4)(*[column(nm) for nm in names_u])
    val = lib.vehicle_get_output(b("car"), q, uu, c.c_double(s[k]), b("chassis.acceleration.y"))
    assert err(lib) == "" and abs(val - ay[k]) <= 1e-9 * max(1.0, abs(ay[k]))
    lib.vehicle_get_output(b("car"), q, uu, c.c_double(s[k]), b("chassis.banana"))
    assert "is not offered" in err(lib)
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><integral_constraints><x><upper_bound>1</upper_bound></x></integral_constraints></options>"))
    assert "needs <lower_bound> and <upper_bound>" in err(lib)
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><integral_constraints><boost-time><lower_bound>0</lower_bound><upper_bound>10</upper_bound></boost-time></integral_constraints></options>"))
    assert "is not offered" in err(lib)
    lib.delete_variable(b"*")


@pytest.mark.gpu
def test_engine_energy_limited_lap_through_the_c_api(lib, files):
    """<integral_constraints> of the C entry (fastestlapc.cpp:1298-1306) on the vehicle of Catalunya_engine_energy_limited
    (f1_optimal_laptime_test.cpp:1057-1117: smooth throttle coefficient 1e-2, no boost power, throttle dissipation 8e-4): the lap
    ends with 24 MJ of engine energy and within 2e-4 of the stored lap time (several local optima, tests/test_gpu_ipm.py); the
    brake bias as a constant control comes back constant."""
    lib.delete_variable(b"*")
    lib.create_vehicle_from_xml(b("car"), b(files["f1"]))
    lib.vehicle_set_parameter(b("car"), b("vehicle/rear-axle/smooth_throttle_coeff"), c.c_double(1.0e-2))
    lib.vehicle_set_parameter(b("car"), b("vehicle/rear-axle/boost/maximum-power"), c.c_double(0.0))
    lib.create_track_from_xml(b("catalunya"), b(files["catalunya"]))
    s = np.linspace(0.0, lib.track_download_length(b("catalunya")), 501)[:-1]
    c_s = (c.c_double * len(s))(*s)
    options = ("<options><integral_constraints><engine-energy><lower_bound>0.0</lower_bound><upper_bound>24.0</upper_bound></engine-energy></integral_constraints>"
               "<control_variables><front-axle.steering-angle optimal_control_type=\"full-mesh\"><dissipation>50.0</dissipation></front-axle.steering-angle>"
               "<chassis.throttle optimal_control_type=\"full-mesh\"><dissipation>8.0e-4</dissipation></chassis.throttle></control_variables>"
               "<output_variables><prefix>en/</prefix></output_variables></options>")
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b(options))
    assert err(lib) == ""
    assert abs(lib.download_scalar(b("en/integral_quantities.engine-energy")) - 24.0) <= 1e-4
    assert abs(lib.download_scalar(b("en/laptime")) - 79.733713) <= 2e-4 * 79.733713
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s,
                        b("<options><control_variables><chassis.brake-bias optimal_control_type=\"constant\"/></control_variables><output_variables><prefix>bb/</prefix></output_variables></options>"))
    assert err(lib) == ""
    bb = (c.c_double * 500)()
    lib.download_vector(bb, c.c_int(500), b("bb/chassis.brake-bias"))
    assert np.ptp(np.array(bb)) <= 1e-9 and 0.3 < bb[0] < 0.9
    lib.delete_variable(b"*")


@pytest.mark.gpu
def test_gg_diagram_through_the_c_api(lib, files):
    """gg_diagram(ay, ax_max, ax_min, vehicle, v, n) (fastestlapc.h:103) for the kart at 70 km/h: the envelope equals the Python
    surface's (same sequence of GPU solves), in m/s2 as the reference's C entry returns it."""
    from fastest_lap_b200 import gg, workloads
    lib.delete_variable(b"*")
    lib.create_vehicle_from_xml(b("kart"), b(files["kart"]))
    n = 16
    ay, amax, amin = (c.c_double * n)(), (c.c_double * n)(), (c.c_double * n)()
    lib.gg_diagram(ay, amax, amin, b("kart"), c.c_double(70.0 / 3.6), c.c_int(n))
    assert err(lib) == ""
    r = gg.gg_diagram(workloads._vehicles()["roberto_lot_kart_2016"], np.array([70.0 / 3.6]), n_lat=n)
    assert r["ax_max_solved"].all() and r["ax_min_solved"].all()
    np.testing.assert_allclose(np.array(ay), r["ay"][0], rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.array(amax), r["ax_max"][0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array(amin), r["ax_min"][0], rtol=0, atol=1e-6)
    assert np.all(np.array(amax) >= np.array(amin) - 1e-9)
    lib.delete_variable(b"*")


@pytest.mark.gpu
def test_warm_start_and_open_section_through_the_c_api(lib, files):
    """<save_warm_start> / <warm_start> (fastestlapc.cpp:1561-1568, 1712-1713) and <closed_simulation> false with <initial_condition>
    from the tables (:1237-1249) through the C entry: the warm-started lap of a heavier car agrees with its cold lap in fewer
    iterations; the open section keeps its first node and is not slower than the same stretch of the closed lap."""
    lib.delete_variable(b"*")
    lib.create_vehicle_from_xml(b("car"), b(files["f1"]))
    lib.create_track_from_xml(b("catalunya"), b(files["catalunya"]))
    s = np.linspace(0.0, lib.track_download_length(b("catalunya")), 501)[:-1]
    c_s = (c.c_double * len(s))(*s)

    def run(options, points=None):
        m = c_s if points is None else (c.c_double * len(points))(*points)
        lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s) if points is None else len(points)), m, b("<options>%s</options>" % options))
        return err(lib)

    def vec(name, n=500):
        v = (c.c_double * n)()
        lib.download_vector(v, c.c_int(n), b(name))
        return np.array(v)
    assert "without a simulation saved" in run("<warm_start>true</warm_start>")
    assert run("<save_warm_start>true</save_warm_start><output_variables><prefix>base/</prefix></output_variables>") == ""
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/mass"), c.c_double(680.0))
    assert run("<warm_start>true</warm_start><output_variables><prefix>warm/</prefix></output_variables>") == ""
    assert run("<output_variables><prefix>cold/</prefix></output_variables>") == ""
    t0, tw, tc = (lib.download_scalar(b(k + "/laptime")) for k in ("base", "warm", "cold"))
    assert tw > t0 + 1e-3 and abs(tw - tc) <= 2e-5 * tc
    assert lib.download_scalar(b("warm/optimization_data.number_of_iterations")) < lib.download_scalar(b("cold/optimization_data.number_of_iterations"))
    # open section: nodes 0..99 from the closed lap's first node
    names_in = ["front-axle.left-tire.kappa", "front-axle.right-tire.kappa", "rear-axle.left-tire.kappa", "rear-axle.right-tire.kappa", "chassis.velocity.x",
                "chassis.velocity.y", "chassis.omega.z", "chassis.force_z_fl_g", "chassis.force_z_fr_g", "chassis.force_z_rl_g", "chassis.force_z_rr_g", "time",
                "road.lateral-displacement", "road.track-heading-angle"]
    names_u = ["front-axle.steering-angle", "rear-axle.boost", "chassis.throttle", "chassis.brake-bias"]
    q = (c.c_double * 14)(*[vec("cold/" + nm)[0] for nm in names_in])
    u = (c.c_double * 4)(*[vec("cold/" + nm)[0] for nm in names_u])
    lib.create_vector(b("q_start"), c.c_int(14), q)
    lib.create_vector(b("u_start"), c.c_int(4), u)
    assert "initial condition must be provided" in run("<closed_simulation>false</closed_simulation>", s[:100])
    assert run("<closed_simulation>false</closed_simulation><initial_condition><q from_table=\"q_start\"/><u from_table=\"u_start\"/></initial_condition>"
               "<output_variables><prefix>open/</prefix></output_variables>", s[:100]) == ""
    uo, to, tl = vec("open/chassis.velocity.x", 100), vec("open/time", 100), vec("cold/time")
    assert uo[0] == q[4] and to[0] == tl[0] and lib.download_scalar(b("open/laptime")) == to[-1]
    assert 0.9 * (tl[99] - tl[0]) <= to[-1] - to[0] <= tl[99] - tl[0] + 1e-6
    assert np.max(np.abs(uo[:20] - vec("cold/chassis.velocity.x")[:20])) <= 0.5
    # sensitivities of the open section (the reference's Catalunya_chicane case is an open section, f1_sensitivity_analysis.cpp:214-291):
    # the given first node does not move, the derivative of the time of the last node against the section re-solved with the mass moved
    lib.vehicle_declare_new_constant_parameter(b("car"), b("vehicle/chassis/mass"), b("mass"), c.c_double(680.0))
    open_opts = "<closed_simulation>false</closed_simulation><initial_condition><q from_table=\"q_start\"/><u from_table=\"u_start\"/></initial_condition>"
    assert run(open_opts + "<compute_sensitivity>true</compute_sensitivity><output_variables><prefix>os/</prefix></output_variables>", s[:100]) == ""
    dT = lib.download_scalar(b("os/derivatives/laptime/mass"))
    du, dt = vec("os/derivatives/chassis.velocity.x/mass", 100), vec("os/derivatives/time/mass", 100)
    assert du[0] == 0.0 and dt[0] == 0.0 and abs(dt[-1] - dT) <= 1e-15 and dT > 0.0
    times = {}
    for sign in (1.0, -1.0):
        lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/mass"), c.c_double(680.0 + sign * 0.05))
        assert run(open_opts + "<output_variables><prefix>om/</prefix></output_variables>", s[:100]) == ""
        times[sign] = lib.download_scalar(b("om/laptime"))
    print("open section: d time / d mass %.8f s/kg, re-solved %.8f" % (dT, (times[1.0] - times[-1.0]) / 0.1))
    assert abs(dT - (times[1.0] - times[-1.0]) / 0.1) <= 1.0e-6
    lib.delete_variable(b"*")


@pytest.mark.gpu
def test_sensitivities_through_the_c_api(lib, files):
    """<compute_sensitivity> (fastestlapc.cpp:1254, 1525): parameters declared with vehicle_declare_new_constant_parameter get
    "<prefix>derivatives/laptime/<alias>" and "<prefix>derivatives/<variable>/<alias>" (:1585-1615, 1687-1706).  Checked as the
    reference checks its own analysis (f1_sensitivity_analysis.cpp:20-26): against the difference of two laps re-solved with the
    parameter moved, lap time to 1e-6, states to 1e-5."""
    lib.delete_variable(b"*")
    lib.create_vehicle_from_xml(b("car"), b(files["f1"]))
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/inertia/Ixx"), c.c_double(0.0))
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/inertia/Iyy"), c.c_double(0.0))
    lib.vehicle_declare_new_constant_parameter(b("car"), b("vehicle/chassis/mass"), b("mass"), c.c_double(660.0))
    lib.vehicle_declare_new_constant_parameter(b("car"), b("vehicle/chassis/aerodynamics/cd"), b("cd"), c.c_double(0.9))
    assert err(lib) == ""
    lib.vehicle_declare_new_constant_parameter(b("car"), b("vehicle/chassis/mass"), b("mass"), c.c_double(660.0))
    assert "already declared" in err(lib)
    lib.create_track_from_xml(b("catalunya"), b(files["catalunya"]))
    s = np.linspace(0.0, lib.track_download_length(b("catalunya")), 501)[:-1]
    c_s = (c.c_double * len(s))(*s)
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><compute_sensitivity>true</compute_sensitivity><output_variables><prefix>run/</prefix></output_variables></options>"))
    assert err(lib) == ""
    dT = {a: lib.download_scalar(b("run/derivatives/laptime/" + a)) for a in ("mass", "cd")}
    print("d laptime / d mass %.8f s/kg, / d cd %.8f s" % (dT["mass"], dT["cd"]))
    assert dT["mass"] > 0.0 and dT["cd"] > 0.0

    def vector(name):
        v = (c.c_double * 500)()
        lib.download_vector(v, c.c_int(500), b(name))
        return np.array(v)
    du = vector("run/derivatives/chassis.velocity.x/mass")
    t0, u0 = lib.download_scalar(b("run/laptime")), vector("run/chassis.velocity.x")
    # the same lap with the mass moved by +-0.05 kg
    laps = {}
    for sign in (1.0, -1.0):
        lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/mass"), c.c_double(660.0 + sign * 0.05))
        lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><output_variables><prefix>moved/</prefix></output_variables></options>"))
        assert err(lib) == ""
        laps[sign] = (lib.download_scalar(b("moved/laptime")), vector("moved/chassis.velocity.x"))
    fd_T = (laps[1.0][0] - laps[-1.0][0]) / 0.1
    fd_u = (laps[1.0][1] - laps[-1.0][1]) / 0.1
    print("re-solved laps: %.8f s/kg; max |du/dm - difference| %.2e of %.2e" % (fd_T, np.max(np.abs(du - fd_u)), np.max(np.abs(du))))
    assert abs(dT["mass"] - fd_T) <= 1.0e-6
    assert np.max(np.abs(du - fd_u)) <= 1.0e-5
    assert abs(t0 - 0.5 * (laps[1.0][0] + laps[-1.0][0])) <= 1.0e-6
    lib.vehicle_declare_new_variable_parameter(b("car"), b("vehicle/chassis/aerodynamics/cl"), b("cl0;cl1"), c.c_int(2), (c.c_double * 2)(3.0, 3.1), c.c_int(2),
                                               (c.c_int * 2)(0, 1), (c.c_double * 2)(0.0, 3000.0))
    # the values of a parameter that varies along the lap are parameters of the analysis too, one alias each (fastestlapc.cpp:962-981)
    lib.vehicle_set_parameter(b("car"), b("vehicle/chassis/mass"), c.c_double(660.0))
    lib.optimal_laptime(b("car"), b("catalunya"), c.c_int(len(s)), c_s, b("<options><compute_sensitivity>true</compute_sensitivity><output_variables><prefix>var/</prefix></output_variables></options>"))
    assert err(lib) == ""
    dcl = [lib.download_scalar(b("var/derivatives/laptime/cl%d" % k)) for k in range(2)]
    print("d laptime / d cl0 %.6f, / d cl1 %.6f s" % tuple(dcl))
    assert err(lib) == "" and dcl[0] < 0.0 and dcl[1] < 0.0          # more downforce, shorter lap
    lib.vehicle_declare_new_variable_parameter(b("car"), b("vehicle/chassis/aerodynamics/cl"), b("only_one"), c.c_int(2), (c.c_double * 2)(3.0, 3.1), c.c_int(2),
                                               (c.c_int * 2)(0, 1), (c.c_double * 2)(0.0, 3000.0))
    assert "one alias per parameter value" in err(lib)
    lib.delete_variable(b"*")
