"""Host-side description of one collocation NLP (what `Optimal_laptime::compute_direct / compute_derivative` set up,
src/core/applications/optimal_laptime.hpp:276-730, 734-1220) in the plain arrays the C-ABI takes.

Nothing here evaluates the model: this module only gathers per-node track constants, the vehicle parameter vector,
the mesh and the (fixed) first node of open problems.
"""
import numpy as np

from .vehicle import MODEL_F1, MODEL_KART

FORM_DIRECT, FORM_DERIVATIVE = 0, 1

N_INPUTS = {MODEL_F1: 14, MODEL_KART: 13}
N_CONTROLS = {MODEL_F1: 4, MODEL_KART: 2}
FULL_MESH_CONTROLS = {MODEL_F1: (0, 2), MODEL_KART: (0, 1)}     # steering + throttle (fastestlapc.cpp:1412-1440)
# C-API default dissipations (fastestlapc.cpp:1412-1440)
DEFAULT_DISSIPATION = {MODEL_F1: (50.0, 20.0 * 8.0e-4), MODEL_KART: (1.0e-2, 200 * 200.0 * 1.0e-10)}


class LapProblem:
    _ARRAYS = ("s", "track_nodes", "wl", "wr", "params", "dissipation", "q0", "u0", "du0", "ctrl_default")
    _SCALARS = ("model", "form", "closed", "elevation", "track_length", "sigma", "kart_direct_torque")

    def __init__(self, model, form, closed, s, track_length, track_nodes, wl, wr, params, dissipation=None, sigma=0.5,
                 elevation=False, q0=None, u0=None, du0=None, ctrl_default=None, kart_direct_torque=False, node_params=None, wind=(0.0, 0.0)):
        self.model, self.form, self.closed, self.elevation = int(model), int(form), bool(closed), bool(elevation)
        # parameters that vary with the arclength (Model_parameter, src/core/foundation/model_parameter.h:51-77): one parameter
        # vector per mesh node, [n_points][n_params]; `params` is then node 0's
        self.node_params = None if node_params is None else np.ascontiguousarray(node_params, dtype=np.float64)
        # (northward, eastward) wind velocity [m/s]: vehicle/chassis/aerodynamics/wind_velocity/... (chassis.h:276-277)
        self.wind = (float(wind[0]), float(wind[1]))
        self.kart_direct_torque = bool(kart_direct_torque)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.s = f64(s)
        self.track_length = float(track_length)
        self.track_nodes = f64(track_nodes).reshape(len(self.s), 6)
        self.wl, self.wr = f64(wl), f64(wr)
        self.params = f64(params)
        self.dissipation = f64(DEFAULT_DISSIPATION[self.model] if dissipation is None else dissipation)
        self.sigma = float(sigma)
        if ctrl_default is None:
            # not-optimised controls keep their default: F1 boost = 0, brake-bias = chassis/brake_bias
            ctrl_default = [0.0, 0.0, 0.0, self.params[18]] if self.model == MODEL_F1 else [0.0, 0.0]
        self.ctrl_default = f64(ctrl_default)
        n_in, n_u = N_INPUTS[self.model], N_CONTROLS[self.model]
        self.q0 = f64(np.zeros(n_in) if q0 is None else q0)
        self.u0 = f64(np.zeros(n_u) if u0 is None else u0)
        self.du0 = f64(np.zeros(2) if du0 is None else du0)
        if not self.closed and q0 is None:
            raise ValueError("open problems need the inputs of the first node")
        if len(self.s) < 2:
            raise ValueError("provide at least two values of arclength")
        if np.any(np.diff(self.s) <= 0.0):
            raise ValueError("arclength must be strictly increasing")

    # sizes, as optimal_laptime_data.h:19-30
    @property
    def n_points(self):
        return len(self.s)

    @property
    def n_elements(self):
        return self.n_points if self.closed else self.n_points - 1

    @property
    def nv(self):
        return N_INPUTS[self.model] - 1 + (2 if self.form == FORM_DIRECT else 4)

    @property
    def ncpe(self):
        return N_INPUTS[self.model] - 1 + 6 + (0 if self.form == FORM_DIRECT else 2)

    @property
    def n(self):
        return self.n_elements * self.nv

    @property
    def m(self):
        return self.n_elements * self.ncpe

    @classmethod
    def from_vehicle_track(cls, vehicle, track, s, form=None, closed=None, dissipation=None, sigma=0.5, q0=None, u0=None, du0=None,
                           kart_direct_torque=False):
        if form is None:
            form = FORM_DIRECT if vehicle.model == MODEL_F1 else FORM_DERIVATIVE     # fastestlapc.cpp:1412-1440
        closed = track.closed if closed is None else closed
        s = np.asarray(s, dtype=np.float64)
        if closed and s[-1] >= track.track_length:
            raise ValueError("closed meshes must not repeat the first point")
        elevation = track.has_elevation()
        if elevation and vehicle.model == MODEL_KART:
            raise ValueError("lot2016kart does not support tracks with elevation")      # lot2016kart.h:63-68
        # parameters declared as functions of the arclength are evaluated at the mesh nodes (dynamic_model_car.hpp:62-63)
        node_params = vehicle.params_at(s) if getattr(vehicle, "has_variable_parameters", lambda: False)() else None
        return cls(vehicle.model, form, closed, s, track.track_length, track.node_data(s), track.left_track_limit(s),
                   track.right_track_limit(s), vehicle.params if node_params is None else node_params[0], dissipation, sigma, elevation, q0, u0, du0,
                   None, kart_direct_torque, node_params=node_params, wind=getattr(vehicle, "wind", (0.0, 0.0)))

    def coarsened(self, every):
        """The same lap on every `every`-th node of the mesh (closed laps): the coarse level of a mesh sequence."""
        if not self.closed:
            raise ValueError("mesh sequencing is for closed laps")
        k = slice(None, None, int(every))
        return LapProblem(self.model, self.form, True, self.s[k], self.track_length, self.track_nodes[k], self.wl[k], self.wr[k], self.params,
                          self.dissipation, self.sigma, self.elevation, None, None, None, self.ctrl_default, self.kart_direct_torque,
                          node_params=None if self.node_params is None else self.node_params[k], wind=self.wind)

    def save_npz(self, path, **extra):
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d.update({k: np.float64(getattr(self, k)) for k in self._SCALARS})
        d.update(extra)
        np.savez_compressed(path, **d)

    @classmethod
    def load_npz(cls, path):
        d = np.load(path)
        p = cls(int(d["model"]), int(d["form"]), bool(d["closed"]), d["s"], float(d["track_length"]), d["track_nodes"], d["wl"], d["wr"],
                d["params"], d["dissipation"], float(d["sigma"]), bool(d["elevation"]), d["q0"], d["u0"], d["du0"], d["ctrl_default"],
                bool(d["kart_direct_torque"]))
        extra = {k: d[k] for k in d.files if k not in cls._ARRAYS and k not in cls._SCALARS}
        return p, extra
