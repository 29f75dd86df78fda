"""ctypes front-end of libfastestlap_b200.so (include/fastestlap_b200.h).

Host-side mirror of the reference's NLP hand-off `ipopt_cppad_solve(options, x0, xl, xu, gl, gu, fg, result)`
(src/core/applications/optimal_laptime.hpp:546-551): the object below is what the reference's `FG` functor plus its CppAD
tape are to Ipopt -- f, g, grad f, Jacobian and Lagrangian-Hessian values in a fixed triplet order.
There is no CPU path: construction raises if the CUDA library is missing or no device is usable.
"""
import ctypes as C
import os
import numpy as np

from .problem import LapProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FL_LIB", os.path.join(_HERE, "libfastestlap_b200.so"))
_lib = None

EVAL_FG, EVAL_JAC, EVAL_HESS = 1, 2, 4
EVAL_ALL = 7
EVAL_NO_VALUES = 8     # measurement only: the derivative kernel alone, on the checkpoints of the previous call

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Desc(C.Structure):
    _fields_ = [("model", C.c_int), ("form", C.c_int), ("closed", C.c_int), ("elevation", C.c_int), ("kart_direct_torque", C.c_int),
                ("n_points", C.c_int), ("s", _dp), ("track_length", C.c_double), ("track_nodes", _dp), ("wl", _dp), ("wr", _dp),
                ("batch", C.c_int), ("vehicle_params", _dp), ("sigma", C.c_double), ("dissipation", C.c_double * 2),
                ("q0", _dp), ("u0", _dp), ("du0", _dp), ("ctrl_default", _dp), ("device", C.c_int), ("node_params", _dp), ("wind", C.c_double * 2),
                ("aug_kind", C.c_int), ("aug_jump", _dp), ("aug_bounds", C.c_double * 2), ("aug_weights", C.c_double * 5)]


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -m fastest_lap_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.fl_nlp_create.restype = C.c_void_p
    lib.fl_nlp_create.argtypes = [C.POINTER(_Desc)]
    lib.fl_nlp_destroy.argtypes = [C.c_void_p]
    lib.fl_last_error.restype = C.c_char_p
    lib.fl_nlp_variant.restype = C.c_char_p
    lib.fl_nlp_variant.argtypes = [C.c_void_p]
    lib.fl_nlp_sizes.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    lib.fl_nlp_batch.argtypes = [C.c_void_p]
    lib.fl_nlp_bounds.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.fl_problem_bounds.argtypes = [C.POINTER(_Desc), _dp, _dp, _dp, _dp]
    lib.fl_measure_fp64_peak.restype = C.c_double
    lib.fl_measure_fp64_peak.argtypes = [C.c_void_p]
    lib.fl_math_eval.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_long]
    lib.fl_nlp_structure.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    lib.fl_nlp_ops_per_node.argtypes = [C.c_void_p, _ip]
    lib.fl_nlp_workspace_slots.argtypes = [C.c_void_p, _ip, _ip]
    lib.fl_eval_all.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp]
    lib.fl_nlp_n_outputs.argtypes = [C.c_void_p]
    lib.fl_nlp_output_name.restype = C.c_char_p
    lib.fl_nlp_output_name.argtypes = [C.c_void_p, C.c_int]
    lib.fl_eval_outputs.argtypes = [C.c_void_p, _dp, _dp]
    # the signatures of Ipopt's Eval_*_CB (bool in, bool out)
    lib.fl_eval_f.argtypes = [C.c_int, _dp, C.c_bool, _dp, C.c_void_p]
    lib.fl_eval_grad_f.argtypes = [C.c_int, _dp, C.c_bool, _dp, C.c_void_p]
    lib.fl_eval_g.argtypes = [C.c_int, _dp, C.c_bool, C.c_int, _dp, C.c_void_p]
    lib.fl_eval_jac_g.argtypes = [C.c_int, _dp, C.c_bool, C.c_int, C.c_int, _ip, _ip, _dp, C.c_void_p]
    lib.fl_eval_h.argtypes = [C.c_int, _dp, C.c_bool, C.c_double, C.c_int, _dp, C.c_bool, C.c_int, _ip, _ip, _dp, C.c_void_p]
    for fn in (lib.fl_eval_f, lib.fl_eval_grad_f, lib.fl_eval_g, lib.fl_eval_jac_g, lib.fl_eval_h):
        fn.restype = C.c_bool
    for name in ("fl_dev_x", "fl_dev_lambda", "fl_dev_f", "fl_dev_g", "fl_dev_grad", "fl_dev_jac", "fl_dev_hess", "fl_nlp_stream"):
        getattr(lib, name).restype = C.c_void_p
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.fl_eval_device.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.fl_nlp_sync.argtypes = [C.c_void_p]
    lib.fl_time_device.restype = C.c_double
    lib.fl_time_device.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    lib.fl_time_device_ring.restype = C.c_double
    lib.fl_time_device_ring.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, C.c_int]
    lib.fl_kernel_launches.restype = C.c_long
    lib.fl_kernel_launches.argtypes = [C.c_void_p]
    lib.fl_host_alloc.restype = C.c_void_p
    lib.fl_host_alloc.argtypes = [C.c_size_t]
    lib.fl_host_free.argtypes = [C.c_void_p]
    lib.fl_flush_l2.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def pinned_empty(count):
    """float64 numpy array over page-locked host memory (cudaHostAlloc), for transfers at full PCIe rate."""
    lib = load_library()
    p = lib.fl_host_alloc(C.c_size_t(max(int(count), 1) * 8))
    if not p:
        raise MemoryError("cudaHostAlloc failed: %s" % lib.fl_last_error().decode())
    buf = (C.c_double * max(int(count), 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=np.float64, count=int(count))
    return arr, p


def _describe(lib, p, vehicle_params=None, device=0):
    """fl_problem_desc of a LapProblem; returns (desc, arrays to keep alive, batch)."""
    params = p.params if vehicle_params is None else vehicle_params
    params = np.ascontiguousarray(params, dtype=np.float64)
    if params.ndim == 1:
        params = params[None, :]
    nvp = lib.fl_n_vehicle_params(p.model)
    if params.shape[1] != nvp:
        raise ValueError("vehicle parameter vectors must have %d entries" % nvp)
    keep = [params, p.s, p.track_nodes, p.wl, p.wr, p.q0, p.u0, p.du0, p.ctrl_default]
    d = _Desc()
    d.model, d.form, d.closed, d.elevation, d.kart_direct_torque = p.model, p.form, int(p.closed), int(p.elevation), int(p.kart_direct_torque)
    d.n_points, d.s, d.track_length, d.track_nodes, d.wl, d.wr = p.n_points, _ptr(p.s), p.track_length, _ptr(p.track_nodes), _ptr(p.wl), _ptr(p.wr)
    d.batch, d.vehicle_params, d.sigma = params.shape[0], _ptr(params), p.sigma
    d.dissipation[0], d.dissipation[1] = float(p.dissipation[0]), float(p.dissipation[1])
    # (ctrl_default stays NULL: the controls that are not optimised are vehicle parameters to the library, see the header)
    d.q0, d.u0, d.du0, d.ctrl_default, d.device = _ptr(p.q0), _ptr(p.u0), _ptr(p.du0), None, int(device)
    wind = getattr(p, "wind", (0.0, 0.0))
    d.wind[0], d.wind[1] = float(wind[0]), float(wind[1])          # northward, eastward [m/s]
    aug = getattr(p, "aug", None)
    if aug:
        d.aug_kind = int(aug["kind"])
        if aug["jump"] is not None:
            jump = np.ascontiguousarray(aug["jump"], dtype=np.float64)
            keep.append(jump)
            d.aug_jump = _ptr(jump)
        d.aug_bounds[0], d.aug_bounds[1] = float(aug["bounds"][0]), float(aug["bounds"][1])
        for k in range(5):
            d.aug_weights[k] = float(aug["weights"][k])
    node_params = getattr(p, "node_params", None)
    if node_params is not None:
        # parameters that vary along the lap (Model_parameter, model_parameter.h:51-77): [n_points][nvp], or [batch][n_points][nvp]
        if vehicle_params is not None:
            raise ValueError("a problem with parameters per node takes no separate vehicle_params")
        npar = np.ascontiguousarray(node_params, dtype=np.float64)
        if npar.ndim == 2:
            npar = npar[None, :, :]
        if npar.shape[1:] != (p.n_points, nvp):
            raise ValueError("node_params must be [n_points][%d]" % nvp)
        keep.append(npar)
        d.batch, d.node_params = npar.shape[0], _ptr(npar)
        return d, keep, npar.shape[0]
    return d, keep, params.shape[0]


def math_eval(fn, x, device=0):
    """The node programs' elementary function `fn` (0 reciprocal, 1 sqrt, 2 sin, 3 cos, 4 atan, 5 tanh, 6 exp) on the device."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    if not lib.fl_math_eval(int(device), int(fn), _ptr(x), _ptr(y), x.size):
        raise RuntimeError("fl_math_eval: " + lib.fl_last_error().decode())
    return y


def time_device_ring(ring, what, steps, obj_factor=1.0):
    """Device milliseconds of `steps` evaluations issued back to back, step k on ring[k % len(ring)] (fl_time_device_ring)."""
    lib = load_library()
    hs = (C.c_void_p * len(ring))(*[h._h for h in ring])
    ms = lib.fl_time_device_ring(hs, len(ring), int(what), float(obj_factor), int(steps))
    if ms < 0:
        raise RuntimeError("fl_time_device_ring: " + lib.fl_last_error().decode())
    return float(ms)


def problem_bounds(problem):
    """Variable and constraint bounds of a lap (get_bounds_info; optimal_laptime.hpp:326-483): host-only, no device needed."""
    lib = load_library()
    d, keep, _ = _describe(lib, problem)
    xl, xu, gl, gu = np.zeros(problem.n), np.zeros(problem.n), np.zeros(problem.m), np.zeros(problem.m)
    if not lib.fl_problem_bounds(C.byref(d), _ptr(xl), _ptr(xu), _ptr(gl), _ptr(gu)):
        raise RuntimeError("fl_problem_bounds: " + lib.fl_last_error().decode())
    return xl, xu, gl, gu


class CollocationNLP:
    """One batch of collocation NLPs sharing mesh and layout, resident on one GPU."""

    def __init__(self, problem, vehicle_params=None, device=0):
        assert isinstance(problem, LapProblem)
        lib = self._lib = load_library()
        p = self.problem = problem
        d, self._keep, batch = _describe(lib, p, vehicle_params, device)
        params = self._keep[0]
        self._h = lib.fl_nlp_create(C.byref(d))
        if not self._h:
            raise RuntimeError("fl_nlp_create: " + lib.fl_last_error().decode())
        n, m, nj, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib.fl_nlp_sizes(self._h, C.byref(n), C.byref(m), C.byref(nj), C.byref(nh))
        self.n, self.m, self.nnz_jac, self.nnz_hess, self.batch = n.value, m.value, nj.value, nh.value, int(batch)
        self.variant = lib.fl_nlp_variant(self._h).decode()
        ops = (C.c_int * 4)()
        lib.fl_nlp_ops_per_node(self._h, ops)
        self.ops_per_node = dict(values=ops[0], jac=ops[1], hess=ops[2], all=ops[3])
        nval, nck = C.c_int(), C.c_int()
        lib.fl_nlp_workspace_slots(self._h, C.byref(nval), C.byref(nck))
        self.n_values, self.n_checkpoints = nval.value, nck.value

    def close(self):
        if getattr(self, "_h", None):
            # solvers bound to this handle hold a pointer to it: they are destroyed first (also when the garbage collector
            # finalises an abandoned NLP before its solver, e.g. after an exception)
            for ref in list(getattr(self, "_dependents", ())):
                dep = ref()
                if dep is not None:
                    dep.close()
            self._dependents = []
            self._lib.fl_nlp_destroy(self._h)
            self._h = None

    def _register_dependent(self, obj):
        import weakref
        if not hasattr(self, "_dependents"):
            self._dependents = []
        self._dependents.append(weakref.ref(obj))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, ok, what):
        if not ok:
            raise RuntimeError("%s: %s" % (what, self._lib.fl_last_error().decode()))

    def structure(self):
        jr, jc = np.zeros(self.nnz_jac, np.int32), np.zeros(self.nnz_jac, np.int32)
        hr, hc = np.zeros(self.nnz_hess, np.int32), np.zeros(self.nnz_hess, np.int32)
        ip = lambda a: a.ctypes.data_as(_ip)
        self._check(self._lib.fl_nlp_structure(self._h, ip(jr), ip(jc), ip(hr), ip(hc)), "fl_nlp_structure")
        return jr, jc, hr, hc

    def bounds(self):
        xl, xu, gl, gu = np.zeros(self.n), np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        self._check(self._lib.fl_nlp_bounds(self._h, _ptr(xl), _ptr(xu), _ptr(gl), _ptr(gu)), "fl_nlp_bounds")
        return xl, xu, gl, gu

    def eval_all(self, x, lam=None, obj_factor=1.0, want=("f", "g", "grad", "jac", "hess"), out=None):
        """One call through the C-ABI with host buffers: x [batch][n], lam [batch][m] -> dict of host arrays."""
        B = self.batch
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == B * self.n
        if lam is not None:
            lam = np.ascontiguousarray(lam, dtype=np.float64)
            assert lam.size == B * self.m
        shapes = dict(f=(B,), g=(B, self.m), grad=(B, self.n), jac=(B, self.nnz_jac), hess=(B, self.nnz_hess))
        out = {} if out is None else out
        for k in want:
            if k not in out:
                out[k] = np.empty(shapes[k])
        if "hess" in want and lam is None:
            raise ValueError("the Hessian needs multipliers")
        self._check(self._lib.fl_eval_all(self._h, _ptr(x), _ptr(lam), float(obj_factor), *[_ptr(out.get(k)) for k in ("f", "g", "grad", "jac", "hess")]),
                    "fl_eval_all")
        return out

    # Ipopt-style callbacks (problem 0)
    @property
    def output_names(self):
        """Names of the model's output channels (the reference's get_outputs_map keys this library evaluates)."""
        return [self._lib.fl_nlp_output_name(self._h, k).decode() for k in range(self._lib.fl_nlp_n_outputs(self._h))]

    def eval_outputs(self, x):
        """Output channels along the lap at x ([n] or [batch][n]): dict name -> [batch][n_points]."""
        names = self.output_names
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(self.batch, self.n))
        out = np.empty((self.batch, len(names), self.problem.n_points))
        self._check(self._lib.fl_eval_outputs(self._h, _ptr(x), _ptr(out)), "fl_eval_outputs")
        return {nm: out[:, k, :] for k, nm in enumerate(names)}

    def eval_f(self, x, new_x=True):
        v = C.c_double()
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self._lib.fl_eval_f(self.n, _ptr(x), bool(new_x), C.byref(v), self._h), "fl_eval_f")
        return v.value

    def eval_g(self, x, new_x=True):
        g = np.empty(self.m)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self._lib.fl_eval_g(self.n, _ptr(x), bool(new_x), self.m, _ptr(g), self._h), "fl_eval_g")
        return g

    def eval_grad_f(self, x, new_x=True):
        g = np.empty(self.n)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self._lib.fl_eval_grad_f(self.n, _ptr(x), bool(new_x), _ptr(g), self._h), "fl_eval_grad_f")
        return g

    def eval_jac_g(self, x, new_x=True):
        v = np.empty(self.nnz_jac)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self._lib.fl_eval_jac_g(self.n, _ptr(x), bool(new_x), self.m, self.nnz_jac, None, None, _ptr(v), self._h), "fl_eval_jac_g")
        return v

    def eval_h(self, x, lam, obj_factor=1.0, new_x=True):
        v = np.empty(self.nnz_hess)
        x = np.ascontiguousarray(x, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        self._check(self._lib.fl_eval_h(self.n, _ptr(x), bool(new_x), float(obj_factor), self.m, _ptr(lam), True, self.nnz_hess, None, None, _ptr(v), self._h),
                    "fl_eval_h")
        return v

    # device-resident path
    def eval_device(self, what=EVAL_ALL, obj_factor=1.0):
        self._check(self._lib.fl_eval_device(self._h, int(what), float(obj_factor)), "fl_eval_device")

    def sync(self):
        self._check(self._lib.fl_nlp_sync(self._h), "fl_nlp_sync")

    def time_device(self, what=EVAL_ALL, obj_factor=1.0, reps=1):
        ms = self._lib.fl_time_device(self._h, int(what), float(obj_factor), int(reps))
        if ms < 0:
            raise RuntimeError("fl_time_device: " + self._lib.fl_last_error().decode())
        return ms

    def measure_fp64_peak(self):
        """DFMA/s of the device's FP64 pipe (micro-kernel measurement)."""
        return float(self._lib.fl_measure_fp64_peak(self._h))

    def flush_l2(self):
        self._check(self._lib.fl_flush_l2(self._h), "fl_flush_l2")

    @property
    def kernel_launches(self):
        return int(self._lib.fl_kernel_launches(self._h))


class _SteadyDesc(C.Structure):
    _fields_ = [("model", C.c_int), ("batch", C.c_int), ("vehicle_params", _dp), ("spec", _dp), ("device", C.c_int)]


class SteadyStateNLP(CollocationNLP):
    """A batch of steady-state (G-G diagram) NLPs, x = [w_axle, z, phi, mu, psi, delta, ax, ay]: the reference's
    Steady_state functors Solve / Max_lat_acc / Max_lon_acc / Min_lon_acc (src/core/applications/steady_state.h:99-284) over
    lot2016kart::steady_state_equations (src/core/vehicles/lot2016kart.h:124-182).  spec rows: (v, ax, ay, free_ax, free_ay,
    cax, cay).  Everything a CollocationNLP offers (callbacks, eval_all, LapSolver) works on it."""

    def __init__(self, model, vehicle_params, spec, device=0):
        lib = self._lib = load_library()
        lib.fl_steady_create.restype = C.c_void_p
        lib.fl_steady_create.argtypes = [C.POINTER(_SteadyDesc)]
        spec = np.ascontiguousarray(spec, dtype=np.float64).reshape(-1, 7)
        B = spec.shape[0]
        params = np.ascontiguousarray(vehicle_params, dtype=np.float64)
        if params.ndim == 1:
            params = np.ascontiguousarray(np.tile(params[None, :], (B, 1)))
        if params.shape != (B, lib.fl_n_vehicle_params(model)):
            raise ValueError("vehicle parameters must be [batch][%d]" % lib.fl_n_vehicle_params(model))
        self.problem = None
        self._keep = [params, spec]
        d = _SteadyDesc(int(model), B, _ptr(params), _ptr(spec), int(device))
        self._h = lib.fl_steady_create(C.byref(d))
        if not self._h:
            raise RuntimeError("fl_steady_create: " + lib.fl_last_error().decode())
        n, m, nj, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib.fl_nlp_sizes(self._h, C.byref(n), C.byref(m), C.byref(nj), C.byref(nh))
        self.n, self.m, self.nnz_jac, self.nnz_hess, self.batch = n.value, m.value, nj.value, nh.value, B
        self.variant = lib.fl_nlp_variant(self._h).decode()
        ops = (C.c_int * 4)()
        lib.fl_nlp_ops_per_node(self._h, ops)
        self.ops_per_node = dict(values=ops[0], jac=ops[1], hess=ops[2], all=ops[3])
        nval, nck = C.c_int(), C.c_int()
        lib.fl_nlp_workspace_slots(self._h, C.byref(nval), C.byref(nck))
        self.n_values, self.n_checkpoints = nval.value, nck.value
