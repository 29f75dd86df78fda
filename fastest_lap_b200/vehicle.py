"""Vehicle XML -> flat parameter vector.

Host-side mirror of the reference's `create_vehicle_from_xml` (src/main/c/fastestlapc.cpp:103-142): the root
attribute `type` selects the model ("f1-3dof" | "kart-6dof"), every parameter is addressed by its XML path exactly as
the reference's DECLARE_PARAMS tables do (src/core/chassis/chassis.h:255-270, chassis_car_3dof.h, axle_car_3dof.h,
tire_pacejka.h:329-342, chassis_car_6dof.h, axle_car_6dof.h).  The order of the vector is the ABI shared with the CUDA
kernels (include/fastestlap_b200.h: FL_F1_* / FL_KART_* enums).
"""
import xml.etree.ElementTree as ET
import numpy as np

MODEL_F1 = 0
MODEL_KART = 1

_F1_TIRE = ["radius", "reference-load-1", "reference-load-2", "mu-x-max-1", "mu-x-max-2", "kappa-max-1", "kappa-max-2",
            "mu-y-max-1", "mu-y-max-2", "lambda-max-1", "lambda-max-2", "Qx", "Qy"]

F1_PARAM_PATHS = (
    ["chassis/mass", "chassis/inertia/Ixx", "chassis/inertia/Iyy", "chassis/inertia/Izz",
     "chassis/aerodynamics/rho", "chassis/aerodynamics/area", "chassis/aerodynamics/cd", "chassis/aerodynamics/cl",
     "chassis/com/x", "chassis/com/y", "chassis/com/z",
     "chassis/front_axle/x", "chassis/front_axle/z", "chassis/rear_axle/x", "chassis/rear_axle/z",
     "chassis/pressure_center/x", "chassis/pressure_center/y", "chassis/pressure_center/z",
     "chassis/brake_bias", "chassis/roll_balance_coefficient", "chassis/Fz_max_ref2",
     "front-axle/track", "front-axle/inertia", "front-axle/smooth_throttle_coeff", "front-axle/brakes/max_torque",
     "rear-axle/track", "rear-axle/inertia", "rear-axle/smooth_throttle_coeff", "rear-axle/differential_stiffness",
     "rear-axle/brakes/max_torque", "rear-axle/engine/maximum-power", "rear-axle/boost/maximum-power"]
    + ["front-tire/" + p for p in _F1_TIRE] + ["rear-tire/" + p for p in _F1_TIRE])

_KART_TIRE = ["radius", "radial-stiffness", "radial-damping", "nominal-vertical-load", "lambdaFz0", "Fz-max-ref2",
              "longitudinal/pure/pCx1", "longitudinal/pure/pDx1", "longitudinal/pure/pEx1", "longitudinal/pure/pKx1",
              "longitudinal/pure/pKx2", "longitudinal/pure/pKx3", "longitudinal/combined/rBx1", "longitudinal/combined/rCx1",
              "lateral/pure/pCy1", "lateral/pure/pDy1", "lateral/pure/pEy1", "lateral/pure/pKy1", "lateral/pure/pKy2",
              "lateral/pure/pKy4", "lateral/combined/rBy1", "lateral/combined/rCy1"]

KART_PARAM_PATHS = (
    ["chassis/mass", "chassis/inertia/Ixx", "chassis/inertia/Iyy", "chassis/inertia/Izz", "chassis/inertia/Ixz",
     "chassis/aerodynamics/rho", "chassis/aerodynamics/cd", "chassis/aerodynamics/cl", "chassis/aerodynamics/area",
     "chassis/com/x", "chassis/com/y", "chassis/com/z",
     "chassis/front_axle/x", "chassis/front_axle/z", "chassis/rear_axle/x", "chassis/rear_axle/z",
     "front-axle/track", "front-axle/stiffness/chassis", "front-axle/stiffness/antiroll",
     "front-axle/beta-steering/left", "front-axle/beta-steering/right",
     "rear-axle/track", "rear-axle/stiffness/chassis", "rear-axle/stiffness/antiroll", "rear-axle/inertia",
     "rear-axle/smooth_throttle_coeff", "rear-axle/brakes/max_torque", "rear-axle/engine/maximum-power"]
    + ["front-tire/" + p for p in _KART_TIRE] + ["rear-tire/" + p for p in _KART_TIRE])

# optional parameters that are not part of the kernels' parameter vector: (northward, eastward) wind (chassis.h:276-277, default 0)
WIND_PATHS = ("chassis/aerodynamics/wind_velocity/northward", "chassis/aerodynamics/wind_velocity/eastward")

_VEC3 = {"x": 0, "y": 1, "z": 2}


def _lookup(root, path):
    """Value of `path` below the <vehicle> root.  Three-vectors may be written either as <x>,<y>,<z> children
    (F1 database files) or as one whitespace separated text node (kart database files)."""
    node = root.find(path)
    if node is not None and node.text is not None and node.text.strip():
        return float(node.text.split()[0]) if len(node.text.split()) == 1 else None
    parent_path, _, leaf = path.rpartition("/")
    parent = root.find(parent_path)
    if parent is not None and leaf in _VEC3 and parent.text and len(parent.text.split()) == 3:
        return float(parent.text.split()[_VEC3[leaf]])
    return None


class Vehicle:
    def __init__(self, model, params, name=""):
        self.model = model
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self.name = name
        self.paths = F1_PARAM_PATHS if model == MODEL_F1 else KART_PARAM_PATHS
        self.wind = [0.0, 0.0]

    def copy(self):
        v = Vehicle(self.model, self.params.copy(), self.name)
        v.wind = list(self.wind)
        v.variable = {k: tuple(np.array(a) for a in val) for k, val in getattr(self, "variable", {}).items()}
        v.declared = list(getattr(self, "declared", []))
        v.variable_aliases = {k: list(a) for k, a in getattr(self, "variable_aliases", {}).items()}
        return v

    def declare_parameter(self, path, alias, value):
        """A constant parameter registered under an alias (vehicle_declare_new_constant_parameter, fastestlapc.cpp:899-939; the
        reference's Model_parameters list): its value is set, and optimal_laptime's sensitivity analysis differentiates with
        respect to it."""
        key = path[len("vehicle/"):] if path.startswith("vehicle/") else path
        self.set_parameter(path, value)
        if key not in self.paths:
            raise KeyError('Parameter "%s" cannot be differentiated' % path)
        if not hasattr(self, "declared"):
            self.declared = []
        if any(a == alias for a, _ in self.declared):
            raise KeyError('Parameter alias "%s" already declared' % alias)
        self.declared.append((alias, key))

    def declare_variable_parameter(self, path, values, mesh_indexes, mesh_points, aliases=None):
        """A parameter that varies with the arclength (vehicle_declare_new_variable_parameter, fastestlapc.cpp:941-981;
        Model_parameter, src/core/foundation/model_parameter.h:18-77): `values` and a mesh of (arclength, index into values)
        pairs, blended piecewise linearly between consecutive mesh points."""
        key = path[len("vehicle/"):] if path.startswith("vehicle/") else path
        if key not in self.paths:
            raise KeyError('Parameter "%s" was not found' % path)
        values, idx, pts = np.asarray(values, dtype=np.float64), np.asarray(mesh_indexes, dtype=int), np.asarray(mesh_points, dtype=np.float64)
        if len(idx) != len(pts) or len(values) == 0 or (len(pts) and (idx.min() < 0 or idx.max() >= len(values))):
            raise ValueError("mesh indexes must address the parameter values, one per mesh point")
        if np.any(np.diff(pts) < 0.0):
            raise ValueError("mesh points must be sorted")
        if not hasattr(self, "variable"):
            self.variable = {}
        self.variable[key] = (values, idx, pts)
        self.params[self.paths.index(key)] = values[0]
        # one alias per value ("a; b; c" in the reference, fastestlapc.cpp:962-969): the names of the derivatives of the sensitivity analysis
        if not hasattr(self, "variable_aliases"):
            self.variable_aliases = {}
        if aliases is not None:
            names = [a.strip() for a in aliases.split(";")] if isinstance(aliases, str) else [str(a).strip() for a in aliases]
            if len(names) != len(values):
                raise ValueError("one alias per parameter value (got %d for %d values)" % (len(names), len(values)))
            self.variable_aliases[key] = names

    def has_variable_parameters(self):
        return bool(getattr(self, "variable", {}))

    def params_at(self, s):
        """[len(s)][n_params]: the parameter vector at every arclength of `s` -- Model_parameter::operator()(s)
        (model_parameter.h:51-77): the first value before the first mesh point, the last value after the last one (the
        reference returns _values.front() / _values.back() there, whatever the mesh indexes say), linear in between."""
        s = np.asarray(s, dtype=np.float64)
        out = np.tile(self.params, (len(s), 1))
        for key, (values, idx, pts) in getattr(self, "variable", {}).items():
            col = self.paths.index(key)
            if len(values) == 1 or len(pts) == 0:
                out[:, col] = values[0]
                continue
            k = np.searchsorted(pts, s, side="right")              # first mesh point > s (std::upper_bound)
            v = np.empty(len(s))
            lo, hi = k == 0, k == len(pts)
            v[lo], v[hi] = values[0], values[-1]
            mid = ~(lo | hi)
            km = k[mid]
            xi = (s[mid] - pts[km - 1]) / (pts[km] - pts[km - 1])
            v[mid] = values[idx[km - 1]] + xi * (values[idx[km]] - values[idx[km - 1]])
            out[:, col] = v
        return out

    def set_parameter(self, path, value):
        """Mirror of vehicle_set_parameter (fastestlapc.cpp): `path` is "vehicle/..." as in the reference."""
        key = path[len("vehicle/"):] if path.startswith("vehicle/") else path
        if key in WIND_PATHS:
            self.wind[WIND_PATHS.index(key)] = float(value)
            return
        if key not in self.paths:
            raise KeyError('Parameter "%s" was not found' % path)
        self.params[self.paths.index(key)] = float(value)

    def get_parameter(self, path):
        key = path[len("vehicle/"):] if path.startswith("vehicle/") else path
        if key in WIND_PATHS:
            return float(self.wind[WIND_PATHS.index(key)])
        return float(self.params[self.paths.index(key)])

    def to_xml(self):
        """The vehicle as a database file (vehicle_save_as_xml of the reference): <vehicle type=...> with one leaf per
        parameter path; `vehicle_from_xml` reads it back to the same parameter vector."""
        root = ET.Element("vehicle", type="f1-3dof" if self.model == MODEL_F1 else "kart-6dof")
        extra = [(pth, w) for pth, w in zip(WIND_PATHS, self.wind) if w != 0.0]
        for path, value in list(zip(self.paths, self.params)) + extra:
            node = root
            for part in path.split("/"):
                nxt = node.find(part)
                node = ET.SubElement(node, part) if nxt is None else nxt
            node.text = repr(float(value))
        ET.indent(root, space="    ")
        return ET.tostring(root, encoding="unicode")


def vehicle_from_xml(source, name=""):
    """`source` is a file name or an XML string."""
    text = source if source.lstrip().startswith("<") else open(source).read()
    root = ET.fromstring(text)
    vtype = root.attrib.get("type")
    if vtype == "f1-3dof":
        model, paths = MODEL_F1, F1_PARAM_PATHS
    elif vtype == "kart-6dof":
        model, paths = MODEL_KART, KART_PARAM_PATHS
    else:
        raise ValueError('vehicle type "%s" not recognized' % vtype)
    values = []
    for p in paths:
        v = _lookup(root, p)
        if v is None:
            raise ValueError('vehicle parameter "%s" is missing' % p)
        values.append(v)
    veh = Vehicle(model, values, name)
    for k, pth in enumerate(WIND_PATHS):
        w = _lookup(root, pth)
        if w is not None:
            veh.wind[k] = w
    # the engine as a power curve over the engine speed (rpm-data / power-data / gear-ratio, src/core/actuators/engine.hpp:17-28)
    # interpolates with lion's sPolynomial, which is not part of the reference tree: refused rather than approximated
    for axle in ("rear-axle", "front-axle"):
        if root.find(axle + "/engine/rpm-data") is not None or root.find(axle + "/engine/power-data") is not None:
            raise ValueError('engine power curves ("%s/engine/rpm-data") are not supported: only engine/maximum-power' % axle)
    return veh
