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
the reference makes of it is a parameter sweep with vehicle_set_parameter.  Open meshes take their first node from
<initial_condition><q from_table=.../><u from_table=.../> (fastestlapc.cpp:1237-1249) and start from it replicated
(:1546-1557); <write_xml> / <xml_file_name> write the reference's solution file (optimal_laptime.hpp xml()).
<control_variables> with chassis.brake-bias optimal_control_type="constant" | "hypermesh" (+ <hypermesh>), or ONE entry of
<integral_constraints> (engine-energy, tire-*-energy) on flat closed F1 laps: carried node-wise, see problem.py.  More than one
of them, other controls on a hypermesh and boost-time raise an exception that names what was asked.
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


def print_variable(variable_name):
    v = _get(variable_name)
    if isinstance(v, Vehicle):
        print(v.to_xml())
    elif isinstance(v, Track):
        print("track \"%s\": length %.6f m, %s" % (variable_name, v.track_length, "closed" if v.closed else "open"))
    else:
        print(v)


def create_vehicle_empty(name, vehicle_type):
    """A vehicle of the given type with every parameter 0, to be filled with vehicle_set_parameter (fastestlapc.cpp
    create_vehicle_empty)."""
    _check_free(name)
    model = {"f1-3dof": MODEL_F1, "kart-6dof": MODEL_KART}.get(vehicle_type)
    if model is None:
        raise FastestLapError("[ERROR] vehicle type \"%s\" not recognized" % vehicle_type)
    from .vehicle import F1_PARAM_PATHS, KART_PARAM_PATHS
    _table[name] = Vehicle(model, np.zeros(len(F1_PARAM_PATHS if model == MODEL_F1 else KART_PARAM_PATHS)), name)


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


def vehicle_save_as_xml(vehicle_name, xml_file_name):
    with open(xml_file_name, "w") as fh:
        fh.write(_get(vehicle_name, Vehicle).to_xml())


def vehicle_declare_new_constant_parameter(vehicle_name, parameter_path, parameter_alias, parameter_value):
    """The reference registers `parameter_path` under an alias and sets its value (fastestlapc.cpp:899-939); the aliases name the
    derivatives `<compute_sensitivity>` of optimal_laptime stores."""
    try:
        _get(vehicle_name, Vehicle).declare_parameter(parameter_path, parameter_alias.strip(), float(parameter_value))
    except KeyError as e:
        raise FastestLapError("[ERROR] vehicle_declare_new_constant_parameter -> %s" % e.args[0])


def vehicle_declare_new_variable_parameter(vehicle_name, parameter_path, parameter_alias, parameter_values, mesh_parameter_indexes, mesh_points):
    """A vehicle parameter as a piecewise-linear function of the arclength (fastest_lap.py / fastestlapc.cpp:941-981; evaluated
    by Model_parameter::operator()(s), model_parameter.h:51-77): optimal_laptime evaluates it at every mesh node and the kernels
    read one parameter vector per node.  (The aliases name the parameter in the reference's sensitivity analysis, not computed here.)"""
    try:
        _get(vehicle_name, Vehicle).declare_variable_parameter(parameter_path, parameter_values, mesh_parameter_indexes, mesh_points, aliases=parameter_alias)
    except (KeyError, ValueError) as e:
        raise FastestLapError("[ERROR] vehicle_declare_new_variable_parameter -> %s" % e)


_vehicle_track = {}          # vehicle name -> Track (vehicle_change_track): the road vehicle_get_output evaluates on


def vehicle_change_track(vehicle_name, track_name):
    """Vehicles of the reference carry their road; here the track is an argument of optimal_laptime, and remembered for
    vehicle_get_output."""
    _get(vehicle_name, Vehicle)
    _vehicle_track[vehicle_name] = _get(track_name, Track)


def vehicle_type_get_sizes(vehicle_type_name):
    model = {"f1-3dof": MODEL_F1, "kart-6dof": MODEL_KART}.get(vehicle_type_name)
    if model is None:
        raise FastestLapError("[ERROR] vehicle type \"%s\" not recognized" % vehicle_type_name)
    inputs, controls = names_of(model)
    return len(inputs), len(controls), 4          # outputs stored by this path: laptime, x, y, psi


def track_coordinates(track):
    """Centreline, left and right boundary and heading (fastest_lap.py:329-338)."""
    t = _get(track, Track)
    s = t.s_data
    x, y, th = t("x", s), t("y", s), t("yaw", s)
    nl, nr = t("nl", s), t("nr", s)
    return x, y, x + nl * np.sin(th), y - nl * np.cos(th), x - nr * np.sin(th), y + nr * np.cos(th), th


def track_download_length(track_name):
    return _get(track_name, Track).track_length


def track_download_data(track_name, variable_name):
    t = _get(track_name, Track)
    keys = {"arclength": None, "centerline.x": "x", "centerline.y": "y", "heading-angle": "yaw", "curvature": "yaw_dot",
            "distance-left-boundary": "nl", "distance-right-boundary": "nr"}
    if variable_name in ("left.x", "left.y", "right.x", "right.y"):
        # track boundaries: centreline -+ distance * (sin(theta), -cos(theta)) (fastestlapc.cpp track_download_data)
        x, y, xl, yl, xr, yr, _ = track_coordinates(track_name)
        return {"left.x": xl, "left.y": yl, "right.x": xr, "right.y": yr}[variable_name]
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
    node = root.find("compute_sensitivity")
    sensitivity = node is not None and (node.text or "").strip().lower() in ("true", "1")
    # <integral_constraints><name><lower_bound/><upper_bound/></name></integral_constraints>: fastestlapc.cpp:1298-1306
    integral = []
    node = root.find("integral_constraints")
    if node is not None:
        for var in node:
            if var.find("lower_bound") is None or var.find("upper_bound") is None:
                raise FastestLapError("[ERROR] optimal_laptime -> integral constraint \"%s\" needs <lower_bound> and <upper_bound>" % var.tag)
            integral.append((var.tag, float(var.find("lower_bound").text), float(var.find("upper_bound").text)))
    get = lambda tag, default: float(root.find(tag).text) if root.find(tag) is not None else default
    speed = get("steady_state_speed", get("initial_speed", 50.0))
    prefix = ""
    variables = None          # <output_variables><variables><name/>...: what to store (fastestlapc.cpp:1215-1233); default: everything standard
    ov = root.find("output_variables")
    if ov is not None and ov.find("prefix") is not None:
        prefix = (ov.find("prefix").text or "").strip()
    if ov is not None and ov.find("variables") is not None:
        variables = [v.tag for v in ov.find("variables")]
    diss = None
    hypermesh = {}           # control name -> hypermesh (a constant control = the hypermesh [0]): fastestlapc.cpp:1276-1283
    cv = root.find("control_variables")
    if cv is not None:
        vals = [float(n.find("dissipation").text) for n in cv if n.attrib.get("optimal_control_type", "") == "full-mesh" and n.find("dissipation") is not None]
        if len(vals) == 2:
            diss = vals
        for n in cv:
            kind = n.attrib.get("optimal_control_type", "")
            if kind == "constant":
                hypermesh[n.tag] = [0.0]
            elif kind == "hypermesh":
                if n.find("hypermesh") is None:
                    raise FastestLapError("[ERROR] optimal_laptime -> control variable \"%s\" needs its <hypermesh>" % n.tag)
                hypermesh[n.tag] = [float(v) for v in (n.find("hypermesh").text or "").replace(",", " ").split()]
            elif kind not in ("dont optimize", "full-mesh"):
                raise FastestLapError("[ERROR] Optimal control type \"%s\" not recognized" % kind)          # the reference's message
    closed = True
    if root.find("closed_simulation") is not None:
        closed = (root.find("closed_simulation").text or "").strip().lower() in ("true", "1")
    flag = lambda tag: root.find(tag) is not None and (root.find(tag).text or "").strip().lower() in ("true", "1")
    q_start = u_start = None
    if not closed:
        ic = root.find("initial_condition")
        if ic is None or ic.find("q") is None or ic.find("u") is None:
            raise FastestLapError("For open simulations, the initial condition must be providedin 'options/initial_condition'")     # the reference's message
        q_start = _get(ic.find("q").attrib.get("from_table", ""), np.ndarray)
        u_start = _get(ic.find("u").attrib.get("from_table", ""), np.ndarray)
    # <variable_bounds><name><lower/><upper/></name>...: fastestlapc.cpp:1308-1376, same checks and messages
    var_bounds = {}
    vb = root.find("variable_bounds")
    if vb is not None:
        for var in vb:
            kids = [k.tag for k in var]
            if not (1 <= len(kids) <= 2) or any(k not in ("lower", "upper") for k in kids) or len(set(kids)) != len(kids):
                raise FastestLapError("[ERROR] Optimal_laptime_configuration -> variable_bounds \"%s\" shall have only children named \"lower\" and/or \"upper\"" % var.tag)
            var_bounds[var.tag] = (float(var.find("lower").text) if var.find("lower") is not None else None,
                                   float(var.find("upper").text) if var.find("upper") is not None else None)
    xml_file = None
    if flag("write_xml"):
        if root.find("xml_file_name") is None:
            raise FastestLapError("[ERROR] optimal_laptime -> <write_xml> needs <xml_file_name>")
        xml_file = (root.find("xml_file_name").text or "").strip()
    return dict(sensitivity=sensitivity, variables=variables, integral=integral, hypermesh=hypermesh, q_start=q_start, u_start=u_start, xml_file=xml_file, var_bounds=var_bounds, warm_start=flag("warm_start"), save_warm_start=flag("save_warm_start"), sigma=get("sigma", 0.5), speed=speed, prefix=prefix, dissipation=diss, print_level=int(get("print_level", _print_level)), closed=closed)


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


def trajectory(track_nodes, centreline, s, track_length, closed, u, v, n, alpha, t_start=0.0):
    """Time, lap time, position and heading of a solution, the quantities Optimal_laptime exports besides the NLP variables
    (optimal_laptime.hpp:598-631): time = trapezoid integral of dt/ds = (1 - n kz) / (u cos(alpha) - v sin(alpha))
    (road_curvilinear.hpp:9) with kz the z component of the 3-D road-frame curvature, -sin(phi) dmu/ds + cos(mu) cos(phi)
    dtheta/ds (road_curvilinear.h:49-60; the formula the kernels' per-node constants are built with, csrc/fl_capi.cu);
    position = centreline + n * normal vector of the pitched and rolled road frame, psi = alpha + theta
    (road_curvilinear.hpp:17-18, 125-127).  track_nodes: [n][6] theta, mu, phi and their s-derivatives."""
    nd = np.asarray(track_nodes)
    theta, mu_t, phi_t = nd[:, 0], nd[:, 1], nd[:, 2]
    kz = -np.sin(phi_t) * nd[:, 4] + np.cos(mu_t) * np.cos(phi_t) * nd[:, 3]
    dtds = (1.0 - n * kz) / (u * np.cos(alpha) - v * np.sin(alpha))
    h = np.diff(np.append(s, track_length)) if closed else np.diff(s)
    if closed:
        time = np.concatenate([[0.0], np.cumsum(h[:-1] * 0.5 * (dtds[:-1] + dtds[1:]))])
        laptime = float(np.sum(h * 0.5 * (dtds + np.roll(dtds, -1))))
    else:
        time = t_start + np.concatenate([[0.0], np.cumsum(h * 0.5 * (dtds[:-1] + dtds[1:]))])
        laptime = float(time[-1])          # the time of the last node (optimal_laptime.hpp:631)
    normal = np.stack([np.cos(theta) * np.sin(mu_t) * np.sin(phi_t) - np.sin(theta) * np.cos(phi_t),
                       np.sin(theta) * np.sin(mu_t) * np.sin(phi_t) + np.cos(theta) * np.cos(phi_t),
                       np.cos(mu_t) * np.sin(phi_t)], axis=1)
    xyz = np.asarray(centreline) + np.asarray(n)[:, None] * normal
    return time, laptime, xyz, theta


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
    inputs, controls = names_of(veh.model)
    n_in = len(inputs)
    full_mesh = (0, 2) if veh.model == MODEL_F1 else (0, 1)
    if not closed:
        if len(o["q_start"]) != n_in or len(o["u_start"]) != len(controls):
            raise FastestLapError("[ERROR] optimal_laptime -> the initial condition must have %d states and %d controls" % (n_in, len(controls)))
        if o["warm_start"]:
            raise FastestLapError("[ERROR] optimal_laptime -> <warm_start> of open simulations is not supported by the B200 path yet")
    warm = None
    if o["warm_start"]:
        # the saved simulation's mesh, start point and multipliers with the vehicle as it is now (fastestlapc.cpp:1561-1568)
        if veh.model not in _warm_start:
            raise FastestLapError("[ERROR] optimal_laptime -> <warm_start> without a simulation saved by <save_warm_start>")
        ws = _warm_start[veh.model]
        s = ws["s"].copy()
        if ws["track"] is not trk:
            raise FastestLapError("[ERROR] optimal_laptime -> the warm start was saved for another track")
    p = LapProblem.from_vehicle_track(veh, trk, s, form=form, closed=closed, dissipation=o["dissipation"], sigma=o["sigma"],
                                      q0=o["q_start"], u0=o["u_start"])
    # One quantity that couples the whole lap -- the brake bias on a hypermesh / as a constant, or one restricted integral quantity --
    # is carried node-wise (problem.py, codegen/nlpnode.py); more than one, or another control, is refused by name
    if o["hypermesh"] or o["integral"]:
        try:
            if len(o["hypermesh"]) + len(o["integral"]) > 1:
                raise ValueError("one hypermesh / constant control or one integral constraint per optimal_laptime call (asked: %s)" %
                                 ", ".join(list(o["hypermesh"]) + [q[0] for q in o["integral"]]))
            if o["hypermesh"]:
                (name, sh), = o["hypermesh"].items()
                if veh.model != MODEL_F1 or name != "chassis.brake-bias":
                    raise ValueError("control \"%s\": only the F1's chassis.brake-bias can be optimised as a constant / on a hypermesh" % name)
                p = p.with_hypermesh_brake_bias(sh)
            else:
                p = p.with_integral_constraint(*o["integral"][0])
        except ValueError as e:
            raise FastestLapError("[ERROR] optimal_laptime -> %s" % e)
        if o["warm_start"] or o["save_warm_start"] or o["var_bounds"] or o["xml_file"]:
            raise FastestLapError("[ERROR] optimal_laptime -> <warm_start>, <save_warm_start>, <variable_bounds> and <write_xml> together with "
                                  "hypermesh controls or integral constraints are not supported")
    if o["warm_start"]:
        x0, warm = ws["x"], (ws["lam"], ws["zl"], ws["zu"])
    elif not closed:
        # the given first node replicated over the free nodes (fastestlapc.cpp:1546-1557); control rates 0
        q = np.delete(np.asarray(o["q_start"], dtype=np.float64), n_in - 3)          # time is not an NLP variable
        node = np.concatenate([q, [o["u_start"][c] for c in full_mesh], [0.0, 0.0] if form == FORM_DERIVATIVE else []])
        x0 = np.tile(node, p.n_points - 1)
    elif veh.model == MODEL_F1:
        x0 = workloads.f1_start_point(p, o["speed"])
    else:
        x0 = np.tile(_kart_start_node(veh, o["speed"] * KMH, form), p.n_points)
    bounds = None
    if o["var_bounds"]:
        from .nlp import problem_bounds
        bounds = problem_bounds(p)
        XL, XU = bounds[0].reshape(-1, p.nv), bounds[1].reshape(-1, p.nv)
        state_cols = [nm for nm in inputs if nm != "time"]
        for nm, (lo, hi) in o["var_bounds"].items():
            if nm in state_cols:
                col = state_cols.index(nm)
            elif nm in controls:
                if controls.index(nm) not in full_mesh:
                    continue                                  # bounds of a control that is not optimised have no effect
                col = len(state_cols) + full_mesh.index(controls.index(nm))
            else:
                raise FastestLapError("[ERROR] Optimal_laptime_configuration -> variable \"%s\" not found" % nm)
            if lo is not None:
                XL[:, col] = lo
            if hi is not None:
                XU[:, col] = hi
    declared = list(getattr(veh, "declared", []))
    if o["sensitivity"]:
        # (closed laps and open sections; not together with lap-coupling quantities)
        if p.aug:
            raise FastestLapError("[ERROR] optimal_laptime -> <compute_sensitivity> is not offered together with hypermesh controls or integral constraints")
    # the values of the parameters that vary along the lap are parameters of the analysis too, one alias each (fastestlapc.cpp:962-981)
    declared = [(a, key, None) for a, key in declared]
    for key, names in getattr(veh, "variable_aliases", {}).items():
        declared += [(a, key, k) for k, a in enumerate(names)]
    nlp = CollocationNLP(p)
    solver = LapSolver(nlp, bounds=bounds)
    out = solver.solve(x0, verbose=o["print_level"] >= 5, warm=warm)
    sigma = solver.sigma() if o["sensitivity"] and out["status"][0] >= 1 else None
    solver.close()
    if out["status"][0] < 1 and closed and warm is None:
        # The solver has no restoration phase: a closed lap that does not converge from the requested starting speed is tried
        # again with a larger initial barrier parameter, from the same and from faster steady states (measured over the
        # reference's 35 track files with the F1: 27 converge at once, zolder and three more with these retries; the rest are
        # karting tracks too tight for the car).  ipm.solve_sweep does the same for batches;
        # the reference wraps such retry loops around Ipopt itself, steady_state.hpp:200-226).
        from .ipm import default_options
        opts = default_options()
        opts.mu_init = 1.0
        for factor in (1.0, 2.0, 3.0):
            if veh.model == MODEL_F1:
                x_try = workloads.f1_start_point(p, factor * o["speed"])
            else:
                x_try = np.tile(_kart_start_node(veh, factor * o["speed"] * KMH, form), p.n_points)
            solver = LapSolver(nlp, opts, bounds=bounds)
            out = solver.solve(x_try, verbose=o["print_level"] >= 5)
            sigma = solver.sigma() if o["sensitivity"] and out["status"][0] >= 1 else None
            solver.close()
            if out["status"][0] >= 1:
                break
    channels = {}
    if out["status"][0] >= 1 and o["variables"]:
        # output channels of the model along the lap, evaluated on the GPU at the solution (fastestlapc.cpp:1643-1690)
        known = set(nlp.output_names)
        if any(v in known for v in o["variables"]):
            if not closed:
                raise FastestLapError("[ERROR] optimal_laptime -> output channels of open simulations are not supported")
            ch = nlp.eval_outputs(out["x"][0])
            channels = {v: ch[v][0].copy() for v in o["variables"] if v in known}
    nlp.close()
    if o["save_warm_start"] and out["status"][0] >= 1:
        _warm_start[veh.model] = dict(s=s.copy(), track=trk, x=out["x"][0].copy(), lam=out["y"][0].copy(), zl=out["zl"][0].copy(), zu=out["zu"][0].copy())
    if out["status"][0] < 1:
        raise FastestLapError("[ERROR] optimal_laptime -> the optimization did not converge (status %d)" % out["status"][0])
    X = out["x"][0].reshape(-1, p.nv)
    aug_values = None
    if p.aug:
        aug_values = X[:, p.aug_pos].copy()                   # the control at every node / the integral accumulated up to the node
        X = np.delete(X, [p.aug_pos, p.aug_pos + 1], axis=1)  # back to the reference's node layout
    if not closed:
        X = np.vstack([node, X])          # the fixed first node
    iu, iv = (4, 5) if veh.model == MODEL_F1 else (1, 2)
    n_col, a_col = n_in - 3, n_in - 2          # positions of n and alpha in x (time removed)
    t_start = 0.0 if closed else float(o["q_start"][n_in - 3])
    time, laptime, xyz, yaw = trajectory(trk.node_data(s), trk.position(s), s, trk.track_length, closed, X[:, iu], X[:, iv], X[:, n_col], X[:, a_col], t_start)
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
    for c, nm in enumerate(controls):
        if p.aug and p.aug["kind"] == 1 and nm == "chassis.brake-bias":
            put(nm, aug_values)                                # constant inside every stretch of the hypermesh
        else:
            put(nm, X[:, k + full_mesh.index(c)].copy() if c in full_mesh else np.full(p.n_points, p.ctrl_default[c]))
    if p.aug and p.aug["kind"] == 2:
        put("integral_quantities." + p.aug["name"], float(aug_values[0]))      # fastestlapc.cpp:1619-1641
    put("x", xyz[:, 0].copy()); put("y", xyz[:, 1].copy()); put("psi", yaw + X[:, a_col])
    put("laptime", laptime)
    put("optimization_data.number_of_iterations", float(out["iters"][0]))
    for nm, val in channels.items():
        put(nm, val)
    if o["variables"]:
        unknown = [v for v in o["variables"] if v not in names and v not in channels]
        if unknown:
            raise FastestLapError("[ERROR] optimal_laptime -> requested output variable \"%s\" is not offered by the B200 path" % unknown[0])
    if o["sensitivity"] and declared:
        # derivatives with respect to the declared parameters (fastestlapc.cpp:1585-1615, 1687-1706: "derivatives/<variable>/<alias>";
        # the reference fills chassis.velocity.x and time and leaves the other requested vectors at zero -- here every state and
        # optimised control gets its derivative)
        from .ipm import lap_sensitivities, dlaptime_dp
        perturbed = []
        for alias, key, k_value in declared:
            col = veh.paths.index(key)
            base_value = veh.params[col] if k_value is None else veh.variable[key][0][k_value]
            step = 1.0e-6 * max(abs(base_value), 1.0e-3)
            pm = []
            for sign in (1.0, -1.0):
                v2 = veh.copy()
                if k_value is None:
                    v2.params[col] += sign * step
                else:
                    vals, idx_m, pts_m = v2.variable[key]
                    vals = vals.copy(); vals[k_value] += sign * step
                    v2.variable[key] = (vals, idx_m, pts_m)
                    v2.params[col] = vals[0]
                pm.append(LapProblem.from_vehicle_track(v2, trk, s, form=form, closed=closed, dissipation=o["dissipation"], sigma=o["sigma"],
                                                        q0=o["q_start"], u0=o["u_start"]))
            perturbed.append((pm[0], pm[1], 2.0 * step))
        dxs, _ = lap_sensitivities(p, out["x"][0], out["y"][0], sigma, perturbed)
        h_el = np.diff(np.append(s, trk.track_length)) if closed else np.append(np.diff(s), 0.0)
        for (alias, _, _), dx in zip(declared, dxs):
            dT, d = dlaptime_dp(p, out["x"][0], dx)
            DX = dx.reshape(-1, p.nv)
            if not closed:
                DX = np.vstack([np.zeros(p.nv), DX])          # the given first node does not move (optimal_laptime.hpp:651-659)
            put("derivatives/laptime/" + alias, dT)
            k = 0
            for nm in inputs:
                if nm == "time":
                    # the trapezoid rule over d/dp of dt/ds (optimal_laptime.hpp:700-716)
                    put("derivatives/time/" + alias, np.concatenate([[0.0], np.cumsum(h_el[:-1] * ((1.0 - o["sigma"]) * d[:-1] + o["sigma"] * d[1:]))]))
                else:
                    put("derivatives/%s/%s" % (nm, alias), DX[:, k].copy()); k += 1
            for c, nm in enumerate(controls):
                put("derivatives/%s/%s" % (nm, alias), DX[:, k + full_mesh.index(c)].copy() if c in full_mesh else np.zeros(p.n_points))
    if o["xml_file"]:
        _write_solution_xml(o["xml_file"], veh.model, p, form, closed, s, time, X, out, laptime, controls, full_mesh, xyz, yaw, n_col, a_col)
    return pre, names


def _write_solution_xml(path, model, p, form, closed, s, time, X, out, laptime, controls, full_mesh, xyz, yaw, n_col, a_col):
    """The reference's solution file (Optimal_laptime::xml(), optimal_laptime.hpp: mesh, inputs, controls, x, y, psi,
    multipliers), the format solution.solution_from_xml reads back and <warm_start> files use."""
    from .solution import Solution, Control
    sol = Solution(model)
    sol.is_closed, sol.is_direct, sol.n_points, sol.laptime, sol.s = closed, form == FORM_DIRECT, len(s), laptime, s.copy()
    n_in = len(names_of(model)[0])
    sol.inputs = np.insert(X[:, :n_in - 1], n_in - 3, time, axis=1)
    k = n_in - 1
    for c in range(len(controls)):
        if c in full_mesh:
            j = full_mesh.index(c)
            ctl = Control("full-mesh", values=X[:, k + j].copy(), derivatives=X[:, k + 2 + j].copy() if form == FORM_DERIVATIVE else None)
        else:
            ctl = Control("dont optimize")
        sol.controls.append(ctl)
    sol.x_coord, sol.y_coord, sol.psi = xyz[:, 0].copy(), xyz[:, 1].copy(), yaw + X[:, a_col]
    sol.iter_count, sol.zl, sol.zu, sol.lam = int(out["iters"][0]), out["zl"][0], out["zu"][0], out["y"][0]
    sol.n_variables_per_point, sol.n_constraints_per_element = p.nv, p.ncpe
    with open(path, "w") as fh:
        fh.write(sol.to_xml())


def vehicle_get_output(vehicle, inputs, controls, s, property_name):
    """One output channel of the vehicle at (inputs, controls), at arclength s of the vehicle's track (fastestlapc.h:72,
    fastestlapc.cpp vehicle_get_output): evaluated on the GPU as node 0 of a two-node open section."""
    from .nlp import CollocationNLP
    veh = _get(vehicle, Vehicle)
    trk = _vehicle_track.get(vehicle)
    if trk is None:
        raise FastestLapError("[ERROR] vehicle_get_output -> the vehicle has no track: call vehicle_change_track first")
    n_in, n_u = (14, 4) if veh.model == MODEL_F1 else (13, 2)
    q, u = np.asarray(inputs, dtype=np.float64), np.asarray(controls, dtype=np.float64)
    if q.size != n_in or u.size != n_u:
        raise FastestLapError("[ERROR] vehicle_get_output -> %d inputs and %d controls are expected" % (n_in, n_u))
    form = FORM_DIRECT if veh.model == MODEL_F1 else FORM_DERIVATIVE
    if veh.model == MODEL_F1:
        veh = veh.copy()
        veh.params[18] = u[3]          # the brake bias is a parameter of the node program when it is not optimised (boost: not evaluated)
    ds = min(1.0, 0.5 * (trk.track_length - float(s)))
    p = LapProblem.from_vehicle_track(veh, trk, np.array([float(s), float(s) + ds]), form=form, closed=False, q0=q, u0=u, du0=np.zeros(2))
    nlp = CollocationNLP(p)
    try:
        if property_name not in nlp.output_names:
            raise FastestLapError("[ERROR] vehicle_get_output -> output \"%s\" is not offered (available: %s)" % (property_name, ", ".join(nlp.output_names)))
        node = np.concatenate([np.delete(q, n_in - 3), [u[c] for c in ((0, 2) if veh.model == MODEL_F1 else (0, 1))], [0.0, 0.0] if form == FORM_DERIVATIVE else []])
        return float(nlp.eval_outputs(node)[property_name][0, 0])
    finally:
        nlp.close()


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
