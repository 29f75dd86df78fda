"""Host-side description of one collocation NLP (what `Optimal_laptime::compute_direct / compute_derivative` set up,
src/core/applications/optimal_laptime.hpp:276-730, 734-1220) in the plain arrays the C-ABI takes.

Nothing here evaluates the model: this module only gathers per-node track constants, the vehicle parameter vector,
the mesh and the (fixed) first node of open problems.
"""
import numpy as np

from .vehicle import MODEL_F1, MODEL_KART

FORM_DIRECT, FORM_DERIVATIVE = 0, 1
AUG_NONE, AUG_BRAKE_BIAS, AUG_INTEGRAL = 0, 1, 2
# integral quantities of the F1 (limebeer2014f1.h:284-289), in the order of the weights the C-ABI takes
INTEGRAL_QUANTITIES = ("engine-energy", "tire-fl-energy", "tire-fr-energy", "tire-rl-energy", "tire-rr-energy")

N_INPUTS = {MODEL_F1: 14, MODEL_KART: 13}
N_CONTROLS = {MODEL_F1: 4, MODEL_KART: 2}
FULL_MESH_CONTROLS = {MODEL_F1: (0, 2), MODEL_KART: (0, 1)}     # steering + throttle (fastestlapc.cpp:1412-1440)
# C-API default dissipations (fastestlapc.cpp:1412-1440)
DEFAULT_DISSIPATION = {MODEL_F1: (50.0, 20.0 * 8.0e-4), MODEL_KART: (1.0e-2, 200 * 200.0 * 1.0e-10)}


class LapProblem:
    _ARRAYS = ("s", "track_nodes", "wl", "wr", "params", "dissipation", "q0", "u0", "du0", "ctrl_default")
    _SCALARS = ("model", "form", "closed", "elevation", "track_length", "sigma", "kart_direct_torque")

    def __init__(self, model, form, closed, s, track_length, track_nodes, wl, wr, params, dissipation=None, sigma=0.5,
                 elevation=False, q0=None, u0=None, du0=None, ctrl_default=None, kart_direct_torque=False, node_params=None, wind=(0.0, 0.0), aug=None):
        self.model, self.form, self.closed, self.elevation = int(model), int(form), bool(closed), bool(elevation)
        # one quantity that couples the whole lap in the reference (a hypermesh / constant control, a restricted integral quantity),
        # carried node-wise: dict(kind, jump [n_points] | None, bounds (lo, up), weights [5]); see codegen/nlpnode.py
        self.aug = aug
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
    def n_aug(self):
        return 1 if self.aug else 0

    @property
    def nv(self):
        return N_INPUTS[self.model] - 1 + (2 if self.form == FORM_DIRECT else 4) + 2 * self.n_aug

    @property
    def ncpe(self):
        return N_INPUTS[self.model] - 1 + 6 + (0 if self.form == FORM_DIRECT else 2) + self.n_aug

    @property
    def ctrl_pos(self):
        """Positions of the two full-mesh controls within a node's variables."""
        k = N_INPUTS[self.model] - 1 + 2 * self.n_aug
        return (k, k + 1)

    @property
    def aug_pos(self):
        """Position of the augmented state `a` within a node's variables (the free jump `q` follows), or None."""
        return N_INPUTS[self.model] - 1 if self.aug else None

    def _copy(self, **changes):
        kw = dict(model=self.model, form=self.form, closed=self.closed, s=self.s, track_length=self.track_length, track_nodes=self.track_nodes,
                  wl=self.wl, wr=self.wr, params=self.params, dissipation=self.dissipation, sigma=self.sigma, elevation=self.elevation,
                  q0=self.q0, u0=self.u0, du0=self.du0, ctrl_default=self.ctrl_default, kart_direct_torque=self.kart_direct_torque,
                  node_params=self.node_params, wind=self.wind, aug=self.aug)
        kw.update(changes)
        return LapProblem(**kw)

    def _check_aug(self):
        if self.model != MODEL_F1 or self.form != FORM_DIRECT or self.elevation or not self.closed:
            raise ValueError("hypermesh / constant controls and integral constraints are offered for the F1, direct form, on flat closed laps")
        if self.aug:
            raise ValueError("one hypermesh / constant control or one integral constraint per problem")

    def with_hypermesh_brake_bias(self, s_hypermesh, bounds=(0.0, 1.0)):
        """The brake bias optimised on a hypermesh: constant over [s_hypermesh[k], s_hypermesh[k+1]), the value of node i being the
        one of the last hypermesh point <= s_i (Control_variables::control_array_at_s,
        detail/optimal_laptime_control_variables.h:190-222).  One point (s = 0) = a control optimised as a constant."""
        self._check_aug()
        sh = np.asarray(s_hypermesh, dtype=np.float64)
        if sh.size < 1 or abs(sh[0]) > 1.0e-12 or np.any(np.diff(sh) <= 0.0) or sh[-1] >= self.track_length:
            raise ValueError("the hypermesh starts at s = 0, increases strictly and ends before the track length")
        k = np.searchsorted(sh, self.s, side="right") - 1              # stretch of every node
        jump = np.zeros(self.n_points)
        jump[k != np.roll(k, 1)] = 1.0                                 # the element ending at the node crosses into a new stretch
        return self._copy(aug=dict(kind=AUG_BRAKE_BIAS, jump=jump, bounds=(float(bounds[0]), float(bounds[1])), weights=np.zeros(5),
                                   s_hypermesh=sh, stretch=k))

    def with_integral_constraint(self, name, lower, upper):
        """One restricted integral quantity, lower <= integral of q dt <= upper (Options::integral_quantities,
        optimal_laptime.hpp:142-160; limebeer2014f1.h:284-306)."""
        self._check_aug()
        if name not in INTEGRAL_QUANTITIES:
            raise ValueError("integral quantity '%s' is not offered (one of %s)" % (name, ", ".join(INTEGRAL_QUANTITIES)))
        w = np.zeros(5)
        w[INTEGRAL_QUANTITIES.index(name)] = 1.0
        return self._copy(aug=dict(kind=AUG_INTEGRAL, jump=None, bounds=(float(lower), float(upper)), weights=w, name=name))

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
                          node_params=None if self.node_params is None else self.node_params[k], wind=self.wind,
                          aug=None if not self.aug else dict(self.aug, jump=None if self.aug["jump"] is None else self._coarse_jump(k)))

    def _coarse_jump(self, k):
        st = self.aug["stretch"][k]
        jump = np.zeros(len(st))
        jump[st != np.roll(st, 1)] = 1.0
        return jump

    def save_npz(self, path, **extra):
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d.update({k: np.float64(getattr(self, k)) for k in self._SCALARS})
        if self.aug:
            d.update(aug_kind=np.float64(self.aug["kind"]), aug_bounds=np.asarray(self.aug["bounds"], dtype=np.float64),
                     aug_weights=np.asarray(self.aug["weights"], dtype=np.float64))
            if self.aug["kind"] == AUG_BRAKE_BIAS:
                d.update(aug_s_hypermesh=self.aug["s_hypermesh"])
        d.update(extra)
        np.savez_compressed(path, **d)

    @classmethod
    def load_npz(cls, path):
        d = np.load(path)
        p = cls(int(d["model"]), int(d["form"]), bool(d["closed"]), d["s"], float(d["track_length"]), d["track_nodes"], d["wl"], d["wr"],
                d["params"], d["dissipation"], float(d["sigma"]), bool(d["elevation"]), d["q0"], d["u0"], d["du0"], d["ctrl_default"],
                bool(d["kart_direct_torque"]))
        aug_keys = ("aug_kind", "aug_bounds", "aug_weights", "aug_s_hypermesh")
        if "aug_kind" in d.files:
            if int(d["aug_kind"]) == AUG_BRAKE_BIAS:
                p = p.with_hypermesh_brake_bias(d["aug_s_hypermesh"], tuple(d["aug_bounds"]))
            else:
                p = p.with_integral_constraint(INTEGRAL_QUANTITIES[int(np.argmax(d["aug_weights"]))], float(d["aug_bounds"][0]), float(d["aug_bounds"][1]))
        extra = {k: d[k] for k in d.files if k not in cls._ARRAYS and k not in cls._SCALARS and k not in aug_keys}
        return p, extra
