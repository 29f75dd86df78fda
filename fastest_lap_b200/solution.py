"""Optimal-laptime solution / warm-start files.

Host-side mirror of the reference's `Optimal_laptime_data::construct_from_xml` / `xml()`
(src/core/applications/detail/optimal_laptime_data.h:66-155, 158-339): same element names, same attribute names,
arrays written with 17 significant digits separated by ", ", `laptime` written with six decimals (std::to_string,
optimal_laptime_data.h:190).  The state names are the reference's `get_state_and_control_names()` of each model.
"""
import xml.etree.ElementTree as ET
import numpy as np

from .vehicle import MODEL_F1, MODEL_KART

# input (state) names, in the order of the model's input vector (src/test/vehicles/limebeer2014f1_test.cpp:115-140,
# src/test/vehicles/car_road_curvilinear.cpp:46-76)
F1_INPUT_NAMES = ["front-axle.left-tire.kappa", "front-axle.right-tire.kappa", "rear-axle.left-tire.kappa", "rear-axle.right-tire.kappa",
                  "chassis.velocity.x", "chassis.velocity.y", "chassis.omega.z",
                  "chassis.force_z_fl_g", "chassis.force_z_fr_g", "chassis.force_z_rl_g", "chassis.force_z_rr_g",
                  "time", "road.lateral-displacement", "road.track-heading-angle"]
F1_CONTROL_NAMES = ["front-axle.steering-angle", "rear-axle.boost", "chassis.throttle", "chassis.brake-bias"]
KART_INPUT_NAMES = ["rear-axle.omega", "chassis.velocity.x", "chassis.velocity.y", "chassis.omega.z",
                    "chassis.position.z", "chassis.attitude.phi", "chassis.attitude.mu",
                    "chassis.velocity.z", "chassis.omega.x", "chassis.omega.y",
                    "time", "road.lateral-displacement", "road.track-heading-angle"]
KART_CONTROL_NAMES = ["front-axle.steering-angle", "rear-axle.throttle"]
KEY_NAME = "road.arclength"

CONTROL_TYPES = ("dont optimize", "constant", "hypermesh", "full-mesh")


def names_of(model):
    if model == MODEL_F1:
        return F1_INPUT_NAMES, F1_CONTROL_NAMES
    if model == MODEL_KART:
        return KART_INPUT_NAMES, KART_CONTROL_NAMES
    raise ValueError("unknown model")


def _vec(text):
    if text is None:
        return np.zeros(0)
    return np.array([float(t) for t in text.replace("\n", " ").split(",") if t.strip()], dtype=np.float64)


def _fmt(a):
    return ", ".join("%.17g" % float(v) for v in np.asarray(a).ravel())


class Control:
    def __init__(self, kind="dont optimize", values=None, derivatives=None, hypermesh=None):
        if kind not in CONTROL_TYPES:
            raise ValueError("optimal_control_type attribute not recognize. Options are: 'dont optimize', 'constant', 'hypermesh', 'full-mesh'")
        self.kind = kind
        self.values = np.zeros(0) if values is None else np.asarray(values, dtype=np.float64)
        self.derivatives = np.zeros(0) if derivatives is None else np.asarray(derivatives, dtype=np.float64)
        self.hypermesh = np.zeros(0) if hypermesh is None else np.asarray(hypermesh, dtype=np.float64)


class Solution:
    """One optimal-laptime result: mesh, inputs at every node, controls, multipliers."""

    def __init__(self, model):
        self.model = model
        self.is_closed = True
        self.is_direct = True
        self.n_points = 0
        self.laptime = 0.0
        self.s = np.zeros(0)
        self.inputs = np.zeros((0, 0))      # [n_points][n_inputs]
        self.controls = []                  # list of Control, in control-index order
        self.x_coord = np.zeros(0)
        self.y_coord = np.zeros(0)
        self.psi = np.zeros(0)
        self.iter_count = 0
        self.zl = np.zeros(0)
        self.zu = np.zeros(0)
        self.lam = np.zeros(0)
        self.n_variables_per_point = 0
        self.n_constraints_per_element = 0

    @property
    def n_elements(self):
        return self.n_points if self.is_closed else self.n_points - 1

    def full_mesh_controls(self):
        return [i for i, c in enumerate(self.controls) if c.kind == "full-mesh"]

    def pack_x(self):
        """NLP variable vector in the reference's node-major order (optimal_laptime.hpp:333-375, 1269-1303):
        per node [inputs without time, full-mesh controls, (derivative form) their rates]; open problems skip node 0."""
        n_in = self.inputs.shape[1]
        time_idx = n_in - 3
        cols = [self.inputs[:, j] for j in range(n_in) if j != time_idx]
        fm = self.full_mesh_controls()
        cols += [self.controls[i].values for i in fm]
        if not self.is_direct:
            cols += [self.controls[i].derivatives for i in fm]
        X = np.stack(cols, axis=1)
        if not self.is_closed:
            X = X[1:]
        return np.ascontiguousarray(X).ravel()

    def to_xml(self):
        in_names, u_names = names_of(self.model)
        root = ET.Element("optimal_laptime", dict(n_points=str(self.n_points), type="closed" if self.is_closed else "open",
                                                  is_direct="true" if self.is_direct else "false"))
        ET.SubElement(root, "laptime").text = "%.6f" % self.laptime
        ET.SubElement(root, KEY_NAME).text = _fmt(self.s)
        for j, name in enumerate(in_names):
            ET.SubElement(root, name).text = _fmt(self.inputs[:, j])
        cv = ET.SubElement(root, "control_variables")
        for i, name in enumerate(u_names):
            c = self.controls[i]
            node = ET.SubElement(cv, name, dict(optimal_control_type=c.kind))
            val = ET.SubElement(node, "values")
            if c.kind == "constant":
                val.text = _fmt(c.values[:1])
            elif c.kind in ("hypermesh", "full-mesh"):
                val.text = _fmt(c.values)
            if c.kind == "hypermesh":
                ET.SubElement(node, "hypermesh").text = _fmt(c.hypermesh)
            if c.kind == "full-mesh" and not self.is_direct:
                ET.SubElement(node, "derivatives").text = _fmt(c.derivatives)
        ET.SubElement(root, "x").text = _fmt(self.x_coord)
        ET.SubElement(root, "y").text = _fmt(self.y_coord)
        ET.SubElement(root, "psi").text = _fmt(self.psi)
        od = ET.SubElement(root, "optimization_data", dict(n_variables_per_point=str(self.n_variables_per_point),
                                                           n_constraints_per_element=str(self.n_constraints_per_element)))
        ET.SubElement(od, "number_of_iterations").text = str(int(self.iter_count))
        ET.SubElement(od, "zl").text = _fmt(self.zl)
        ET.SubElement(od, "zu").text = _fmt(self.zu)
        ET.SubElement(od, "lambda").text = _fmt(self.lam)
        ET.indent(root, space="    ")
        return ET.tostring(root, encoding="unicode")


def solution_from_xml(source, model):
    text = source if source.lstrip().startswith("<") else open(source).read()
    root = ET.fromstring(text)
    in_names, u_names = names_of(model)
    sol = Solution(model)
    t = root.attrib.get("type")
    if t not in ("open", "closed"):
        raise ValueError('Incorrect track type, should be "open" or "closed"')
    sol.is_closed = t == "closed"
    d = root.attrib.get("is_direct")
    if d not in ("true", "false"):
        raise ValueError('Incorrect "is_direct" attribute, should be "true" or "false"')
    sol.is_direct = d == "true"
    sol.n_points = int(root.attrib["n_points"])
    sol.laptime = float(root.findtext("laptime"))
    sol.s = _vec(root.findtext(KEY_NAME))
    sol.inputs = np.stack([_vec(root.findtext(n)) for n in in_names], axis=1)
    for name in u_names:
        node = root.find("control_variables/" + name)
        kind = node.attrib.get("optimal_control_type")
        c = Control(kind, _vec(node.findtext("values")))
        if kind == "hypermesh":
            c.hypermesh = _vec(node.findtext("hypermesh"))
        if not sol.is_direct and node.find("derivatives") is not None:
            c.derivatives = _vec(node.findtext("derivatives"))
        sol.controls.append(c)
    sol.x_coord, sol.y_coord, sol.psi = _vec(root.findtext("x")), _vec(root.findtext("y")), _vec(root.findtext("psi"))
    od = root.find("optimization_data")
    sol.n_variables_per_point = int(od.attrib.get("n_variables_per_point", 0))
    sol.n_constraints_per_element = int(od.attrib.get("n_constraints_per_element", 0))
    sol.iter_count = int(od.findtext("number_of_iterations"))
    sol.zl, sol.zu, sol.lam = _vec(od.findtext("zl")), _vec(od.findtext("zu")), _vec(od.findtext("lambda"))
    return sol
