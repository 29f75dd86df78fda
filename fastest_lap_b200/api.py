"""The user-facing surface of the reference's Python module, on the GPU path.

Mirror of src/main/python/fastest_lap.py over src/main/c/fastestlapc.cpp: a table of named objects
(vehicles, tracks, scalars, vectors), `create_vehicle_from_xml`, `create_track_from_xml`, `vehicle_set_parameter`,
`optimal_laptime(vehicle, track, s, options)`, `gg_diagram(vehicle, speed, n_points)`, `download_*`, `delete_variable`.
Same function names, argument meaning and error behaviour (exceptions with the reference's messages where it has them);
the computation runs in libfastestlap_b200.so (callback kernels + device interior-point solver): there is no CPU path.

Supported options of `optimal_laptime` (fastestlapc.cpp:1379-1492): <sigma>, <steady_state_speed> (km/h; the reference
reads <initial_speed> for it, quirk ix of SURVEY 8a-1: both spellings are accepted), <output_variables><prefix>,
<control_variables> dissipations of the two full-mesh controls, <print_level>.  Closed and open meshes; F1 flat and 3-D
tracks (direct form), kart (derivative form, the C API's default); <save_warm_start> keeps the solution (per vehicle
type, fastestlapc.cpp:60-71, 1712-1713) and <warm_start> restarts from it with its multipliers (:1561-1568) -- the use
the reference makes of it is a parameter sweep with vehicle_set_parameter.  Hypermesh controls, integral constraints and
sensitivities are not supported yet (a clear exception is raised).
"""
import xml.etree.ElementTree as ET
import numpy as np

from .vehicle import vehicle_from_xml, Vehicle, MODEL_F1, MODEL_KART
from .track import track_from_xml, Track
from .problem import LapProblem, FORM_DIRECT, FORM_DERIVATIVE, DEFAULT_DISSIPATION
from .solution import names_of, KEY_NAME
from . import workloads

KMH = 1.0 / 3.6
_table = {}
_warm_start = {}          # per vehicle type: the last simulation saved with <save_warm_start> (fastestlapc.cpp:60-71)
_print_level = 0


class FastestLapError(RuntimeError):
    """fastest_lap_exception of the reference (lion)."""


def set_print_level(print_level):
    global _print_level
    _print_level = int(print_level)


def _check_free(name):
    if name in _table:
        raise FastestLapError("variable \"%s\" already exists" % name)          # fastestlapc.cpp:61-84 check_variable_exists_in_any_table


def _get(name, kind=None):
    if name not in _table:
        raise FastestLapError("variable \"%s\" does not exist" % name)
    v = _table[name]
    if kind is not None and not isinstance(v, kind):
        raise FastestLapError("variable \"%s\" has not the expected type" % name)
    return v


def print_variables():
    for k, v in _table.items():
        print("%-40s %s" % (k, type(v).__name__))


def create_vehicle_from_xml(name, database_file):
    _check_free(name)
    _table[name] = vehicle_from_xml(database_file, name)


def create_track_from_xml(name, track_file):
    _check_free(name)
    _table[name] = track_from_xml(track_file, name)


def create_vector(name, data):
    _check_free(name)
    _table[name] = np.array(data, dtype=np.float64)


def create_scalar(name, data):
    _check_free(name)
    _table[name] = float(data)


def copy_variable(old_name, new_name):
    _check_free(new_name)
    v = _get(old_name)
    _table[new_name] = v.copy() if hasattr(v, "copy") else v


def move_variable(old_name, new_name):
    _check_free(new_name)
    _table[new_name] = _table.pop(old_name) if old_name in _table else _get(old_name)


def delete_variable(name):
    if name.endswith("*"):                      # the reference deletes by prefix with a trailing '*' (fastestlapc.cpp:246-330)
        for k in [k for k in _table if k.startswith(name[:-1])]:
            del _table[k]
        return
    _get(name)
    del _table[name]


def variable_type(name):
    v = _get(name)
    if isinstance(v, Vehicle):
        return "f1-3dof" if v.model == MODEL_F1 else "kart-6dof"
    if isinstance(v, Track):
        return "track"
    return "scalar" if isinstance(v, float) else "vector"


def download_scalar(name):
    return _get(name, float)


def download_vector_size(name):
    return len(_get(name, np.ndarray))


def download_vector(name):
    return _get(name, np.ndarray).copy()


def download_variables(prefix, variable_list):
    out = {}
    for v in variable_list:
        x = _get(prefix + v)
        out[v] = x if isinstance(x, float) else x.copy()
    return out


def vehicle_set_parameter(vehicle, parameter_name, parameter_value):
    _get(vehicle, Vehicle).set_parameter(parameter_name, float(parameter_value))


def track_download_length(track_name):
    return _get(track_name, Track).track_length


def track_download_data(track_name, variable_name):
    t = _get(track_name, Track)
    keys = {"arclength": None, "centerline.x": "x", "centerline.y": "y", "heading-angle": "yaw", "curvature": "yaw_dot",
            "distance-left-boundary": "nl", "distance-right-boundary": "nr"}
    if variable_name not in keys:
        raise FastestLapError("[ERROR] track_download_data -> variable_name \"%s\" is not valid" % variable_name)
    return t.s_data.copy() if keys[variable_name] is None else t(keys[variable_name], t.s_data)


def vehicle_type_get_names(vehicle_type_name):
    model = {"f1-3dof": MODEL_F1, "kart-6dof": MODEL_KART}.get(vehicle_type_name)
    if model is None:
        raise FastestLapError("[ERROR] vehicle type \"%s\" not recognized" % vehicle_type_name)
    inputs, controls = names_of(model)
    return KEY_NAME, list(inputs), list(controls), ["laptime", "x", "y", "psi"]


# ---------------------------------------------------------------------------------------------------------------------
def _options(options):
    root = ET.fromstring(options) if options and options.strip() else ET.Element("options")
    for tag in ("integral_constraints", "compute_sensitivity"):
        node = root.find(tag)
        if node is not None and (node.text or "").strip().lower() in ("true", "1") or (tag == "integral_constraints" and node is not None and len(node)):
            raise FastestLapError("[ERROR] optimal_laptime -> option <%s> is not supported by the B200 path yet" % tag)
    get = lambda tag, default: float(root.find(tag).text) if root.find(tag) is not None else default
    speed = get("steady_state_speed", get("initial_speed", 50.0))
    prefix = ""
    ov = root.find("output_variables")
    if ov is not None and ov.find("prefix") is not None:
        prefix = (ov.find("prefix").text or "").strip()
    diss = None
    cv = root.find("control_variables")
    if cv is not None:
        vals = [float(n.find("dissipation").text) for n in cv if n.attrib.get("optimal_control_type", "") == "full-mesh" and n.find("dissipation") is not None]
        if len(vals) == 2:
            diss = vals
    closed = True
    if root.find("closed_simulation") is not None:
        closed = (root.find("closed_simulation").text or "").strip().lower() in ("true", "1")
    flag = lambda tag: root.find(tag) is not None and (root.find(tag).text or "").strip().lower() in ("true", "1")
    return dict(warm_start=flag("warm_start"), save_warm_start=flag("save_warm_start"), sigma=get("sigma", 0.5), speed=speed, prefix=prefix, dissipation=diss, print_level=int(get("print_level", _print_level)), closed=closed)


def _kart_start_node(vehicle, speed, form):
    """Steady state of the kart at `speed` (0 g), solved on the GPU, as one node of the lap NLP (fastestlapc.cpp:1508-1545;
    the throttle of the kart is then set to 0, :1512-1513)."""
    from . import gg
    out = gg.solve_batch(vehicle.params, np.array([[speed, 0, 0, 0, 0, 0, 0.0]]), gg.X0[None, :])
    if out["status"][0] < 1:
        raise FastestLapError("[ERROR] steady-state computation did not converge")
    w, z, phi, mu, psi, delta = out["x"][0, :6]
    q = [(w + 1.0) * speed / 0.139, speed * np.cos(psi), -speed * np.sin(psi), 0.0, z, phi, mu, 0.0, 0.0, 0.0, 0.0, psi]   # time dropped; n = 0, alpha = psi
    node = q + [delta, 0.0]
    if form == FORM_DERIVATIVE:
        node += [0.0, 0.0]
    return np.array(node)


def optimal_laptime(vehicle, track, s, options=""):
    """Computes the optimal lap of `vehicle` on `track` over the mesh `s` and stores the outputs in the table under
    `<prefix>`.  Returns (prefix, variable names), as the reference's Python wrapper does."""
    from .nlp import CollocationNLP
    from .ipm import LapSolver
    veh, trk = _get(vehicle, Vehicle), _get(track, Track)
    o = _options(options)
    s = np.asarray(s, dtype=np.float64)
    if len(s) < 2:
        raise FastestLapError("[ERROR] optimal_laptime -> provide at least two values of arclength")
    closed = o["closed"]          # <closed_simulation>, true by default (fastestlapc.cpp:1237, 1386)
    form = FORM_DIRECT if veh.model == MODEL_F1 else FORM_DERIVATIVE
    if not closed:
        raise FastestLapError("[ERROR] optimal_laptime -> open meshes need <initial_condition> (not supported by the B200 path yet)")
    warm = None
    if o["warm_start"]:
        # the saved simulation's mesh, start point and multipliers with the vehicle as it is now (fastestlapc.cpp:1561-1568)
        if veh.model not in _warm_start:
            raise FastestLapError("[ERROR] optimal_laptime -> <warm_start> without a simulation saved by <save_warm_start>")
        ws = _warm_start[veh.model]
        s = ws["s"].copy()
        if ws["track"] is not trk:
            raise FastestLapError("[ERROR] optimal_laptime -> the warm start was saved for another track")
    p = LapProblem.from_vehicle_track(veh, trk, s, form=form, closed=True, dissipation=o["dissipation"], sigma=o["sigma"])
    if o["warm_start"]:
        x0, warm = ws["x"], (ws["lam"], ws["zl"], ws["zu"])
    elif veh.model == MODEL_F1:
        x0 = workloads.f1_start_point(p, o["speed"])
    else:
        x0 = np.tile(_kart_start_node(veh, o["speed"] * KMH, form), p.n_points)
    nlp = CollocationNLP(p)
    solver = LapSolver(nlp)
    out = solver.solve(x0, verbose=o["print_level"] >= 5, warm=warm)
    solver.close(); nlp.close()
    if o["save_warm_start"] and out["status"][0] >= 1:
        _warm_start[veh.model] = dict(s=s.copy(), track=trk, x=out["x"][0].copy(), lam=out["y"][0].copy(), zl=out["zl"][0].copy(), zu=out["zu"][0].copy())
    if out["status"][0] < 1:
        raise FastestLapError("[ERROR] optimal_laptime -> the optimization did not converge (status %d)" % out["status"][0])
    X = out["x"][0].reshape(p.n_points, p.nv)
    inputs, controls = names_of(veh.model)
    n_in = len(inputs)
    h = np.diff(np.append(s, trk.track_length))
    # time: trapezoid integral of dt/ds, recovered from the objective's lap-time part through the callbacks' own values is
    # not exposed; integrate 1 / (ds/dt) with the curvilinear kinematics (road_curvilinear.hpp:9)
    iu, iv = (4, 5) if veh.model == MODEL_F1 else (1, 2)
    n_col, a_col = n_in - 3, n_in - 2          # positions of n and alpha in x (time removed)
    kz = trk("yaw_dot", s)
    dtds = (1.0 - X[:, n_col] * kz) / (X[:, iu] * np.cos(X[:, a_col]) - X[:, iv] * np.sin(X[:, a_col]))
    time = np.concatenate([[0.0], np.cumsum(h[:-1] * 0.5 * (dtds[:-1] + dtds[1:]))])
    laptime = float(np.sum(h * 0.5 * (dtds + np.roll(dtds, -1))))
    pre = o["prefix"]
    names = []

    def put(name, value):
        _table[pre + name] = value
        names.append(name)
    put(KEY_NAME, s.copy())
    k = 0
    for i, nm in enumerate(inputs):
        if nm == "time":
            put("time", time)
        else:
            put(nm, X[:, k].copy()); k += 1
    full_mesh = (0, 2) if veh.model == MODEL_F1 else (0, 1)
    for c, nm in enumerate(controls):
        put(nm, X[:, k + full_mesh.index(c)].copy() if c in full_mesh else np.full(p.n_points, p.ctrl_default[c]))
    pos = trk.position(s)
    yaw = trk("yaw", s)
    put("x", pos[:, 0] - X[:, n_col] * np.sin(yaw)); put("y", pos[:, 1] + X[:, n_col] * np.cos(yaw)); put("psi", yaw + X[:, a_col])
    put("laptime", laptime)
    put("optimization_data.number_of_iterations", float(out["iters"][0]))
    return pre, names


def gg_diagram(vehicle, speed, n_points):
    """(ay, -ay, ax_max, ax_min) in g at `speed` [m/s] for `n_points` lateral accelerations (fastest_lap.py:269-287)."""
    from . import gg
    veh = _get(vehicle, Vehicle)
    r = gg.gg_diagram(veh.params, np.array([float(speed)]), n_lat=int(n_points))
    if not (r["zero_g_solved"].all() and r["ay_max_solved"].all()):
        raise FastestLapError("[ERROR] gg_diagram -> steady-state computation did not converge")
    to_g = 1.0 if veh.model == MODEL_F1 else 1.0 / 9.81          # the F1's steady-state NLP works in g already (limebeer2014f1.h:51)
    ay = r["ay"][0] * to_g
    return list(ay), list(-ay), list(r["ax_max"][0] * to_g), list(r["ax_min"][0] * to_g)
