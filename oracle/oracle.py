"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The oracle restates, on the CPU, what the reference's CppAD tape + FG functor compute
(/root/reference/src/core/applications/optimal_laptime.hpp:1226-1657); see oracle/README.md for how it is pinned.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_F1, MODEL_KART = 0, 1
FORM_DIRECT, FORM_DERIVATIVE = 0, 1


class _Desc(C.Structure):
    _fields_ = [("model", C.c_int), ("form", C.c_int), ("closed", C.c_int), ("elevation", C.c_int), ("n_points", C.c_int),
                ("kart_direct_torque", C.c_int),
                ("s", C.POINTER(C.c_double)), ("track_length", C.c_double), ("track", C.POINTER(C.c_double)),
                ("params", C.POINTER(C.c_double)), ("sigma", C.c_double), ("dissipation", C.c_double * 2),
                ("q0", C.POINTER(C.c_double)), ("u0", C.POINTER(C.c_double)), ("du0", C.POINTER(C.c_double)),
                ("ctrl_default", C.POINTER(C.c_double)), ("node_params", C.POINTER(C.c_double)), ("wind", C.c_double * 2),
                ("aug_kind", C.c_int), ("aug_jump", C.POINTER(C.c_double)), ("aug_weights", C.c_double * 5)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    # make decides by the sources' time stamps; a prebuilt library without a compiler at hand is used as it is
    r = subprocess.run(["make"] + (["-B"] if force else []) + ["-C", _HERE, "liboracle.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and not os.path.exists(so):
        raise RuntimeError("building the CPU oracle failed:\n" + r.stdout[-2000:])
    # the extended-precision build (the arbiter of ill-conditioned Hessian entries) travels with it; lib_ld() reports a failure
    subprocess.run(["make", "-C", _HERE, "liboracle_ld.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_time_eval.restype = C.c_double
    return _LIB


_LIB_LD = None


def lib_ld():
    """The same oracle with its jets in x87 extended precision (long double, 64-bit mantissa): oracle/jet.hpp."""
    global _LIB_LD
    if _LIB_LD is None:
        so = os.path.join(_HERE, "liboracle_ld.so")
        r = subprocess.run(["make", "-C", _HERE, "liboracle_ld.so"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 and not os.path.exists(so):
            raise RuntimeError("building the extended-precision oracle failed:\n" + r.stdout[-2000:])
        _LIB_LD = C.CDLL(so)
    return _LIB_LD


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class OracleNLP:
    """The collocation NLP of one lap, evaluated on the CPU."""

    def __init__(self, model, form, closed, s, track_length, track_nodes, params, dissipation, sigma=0.5, elevation=False,
                 ctrl_default=None, q0=None, u0=None, du0=None, kart_direct_torque=False, node_params=None, wind=(0.0, 0.0), extended=False, aug=None):
        self._keep = []
        self._lib = lib_ld() if extended else lib()

        def arr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return a

        self.s = arr(s)
        self.track = arr(track_nodes)
        self.params = arr(params)
        if ctrl_default is None:
            ctrl_default = [0.0, 0.0, 0.0, self.params[18]] if model == MODEL_F1 else [0.0, 0.0]
        d = _Desc()
        d.model, d.form, d.closed, d.elevation, d.n_points = model, form, int(closed), int(elevation), len(self.s)
        d.kart_direct_torque = int(kart_direct_torque)
        d.s, d.track_length, d.track, d.params, d.sigma = _dp(self.s), float(track_length), _dp(self.track), _dp(self.params), float(sigma)
        d.dissipation[0], d.dissipation[1] = float(dissipation[0]), float(dissipation[1])
        d.q0, d.u0, d.du0, d.ctrl_default = _dp(arr(q0)), _dp(arr(u0)), _dp(arr(du0)), _dp(arr(ctrl_default))
        d.node_params = _dp(arr(node_params))          # [n_points][n_params]: parameters that vary along the lap, or None
        d.wind[0], d.wind[1] = float(wind[0]), float(wind[1])
        if aug:
            # a hypermesh / constant brake bias or a restricted integral quantity, node-wise (oracle/nlp.hpp)
            d.aug_kind = int(aug["kind"])
            d.aug_jump = _dp(arr(aug["jump"]))
            for k in range(5):
                d.aug_weights[k] = float(aug["weights"][k])
        self.desc = d
        n, m, nv, ncpe = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        mj, mh = C.c_long(), C.c_long()
        self._lib.orc_sizes(C.byref(d), C.byref(n), C.byref(m), C.byref(nv), C.byref(ncpe), C.byref(mj), C.byref(mh))
        self.n, self.m, self.nv, self.ncpe, self._mj, self._mh = n.value, m.value, nv.value, ncpe.value, mj.value, mh.value

    def eval(self, x, lam=None, obj_factor=1.0, derivs=True, threads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        lam = np.zeros(self.m) if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
        f = C.c_double()
        grad, g = np.zeros(self.n), np.zeros(self.m)
        jr, jc, jv = np.zeros(self._mj, np.int32), np.zeros(self._mj, np.int32), np.zeros(self._mj)
        hr, hc, hv = np.zeros(self._mh, np.int32), np.zeros(self._mh, np.int32), np.zeros(self._mh)
        nj, nh = C.c_long(), C.c_long()
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        self._lib.orc_eval(C.byref(self.desc), _dp(x), _dp(lam), C.c_double(obj_factor), int(derivs), int(threads), C.byref(f), _dp(grad), _dp(g),
                       ip(jr), ip(jc), _dp(jv), C.byref(nj), ip(hr), ip(hc), _dp(hv), C.byref(nh))
        out = {"f": f.value, "g": g}
        if derivs:
            out["grad"] = grad
            out["jac"] = (jr[:nj.value], jc[:nj.value], jv[:nj.value])
            out["hess"] = (hr[:nh.value], hc[:nh.value], hv[:nh.value])
        return out

    def outputs(self, x):
        """The 45 output channels of the model at every node ([45][n_points]), in the order of the product's fl_nlp_output_name."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros((45, self.desc.n_points))
        self._lib.orc_outputs(C.byref(self.desc), _dp(x), _dp(out))
        return out

    def time_eval(self, x, lam, obj_factor=1.0, reps=1, threads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        return float(lib().orc_time_eval(C.byref(self.desc), _dp(x), _dp(lam), C.c_double(obj_factor), int(reps), int(threads)))


def triplets_to_dict(rows, cols, vals, drop_zeros=True):
    """Sum duplicates and key by (row, col)."""
    import collections
    d = collections.defaultdict(float)
    for r, c, v in zip(rows.tolist(), cols.tolist(), vals.tolist()):
        d[(r, c)] += v
    if drop_zeros:
        d = {k: v for k, v in d.items() if v != 0.0}
    return dict(d)


def set_wind(northward=0.0, eastward=0.0):
    """Wind used by the single-node entry points f1_car / kart_car (chassis.h:276-277)."""
    lib().orc_set_wind(C.c_double(northward), C.c_double(eastward))


def f1_car(q, u, params, track6=None):
    track6 = np.zeros(6) if track6 is None else np.ascontiguousarray(track6, dtype=np.float64)
    q, u, params = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, u, params))
    st, dst, ext, diag = np.zeros(14), np.zeros(14), np.zeros(6), np.zeros(24)
    dtds = C.c_double()
    lib().orc_f1_car(_dp(q), _dp(u), _dp(params), _dp(track6), _dp(st), _dp(dst), C.byref(dtds), _dp(ext), _dp(diag))
    return dict(states=st, dstates=dst, dtds=dtds.value, extra=ext, kappa=diag[0:4], lam=diag[4:8], omega_w=diag[8:12],
                Fx=diag[12:16], Fy=diag[16:20], N=diag[20:24])


def kart_car(q, u, params, track6=None, direct_torque=False, curvilinear=True):
    track6 = np.zeros(6) if track6 is None else np.ascontiguousarray(track6, dtype=np.float64)
    q, u, params = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, u, params))
    st, dst, ext, diag = np.zeros(13), np.zeros(13), np.zeros(6), np.zeros(30)
    dtds = C.c_double()
    lib().orc_kart_car(_dp(q), _dp(u), _dp(params), _dp(track6), int(direct_torque), int(curvilinear), _dp(st), _dp(dst), C.byref(dtds), _dp(ext), _dp(diag))
    return dict(states=st, dstates=dst, dtds=dtds.value, extra=ext, kappa=diag[0:4], lam=diag[4:8], N=diag[8:12], Fx=diag[12:16],
                Fy=diag[16:20], w=diag[20:24], force=diag[24:27], torque=diag[27:30])


def kart_steady(x, params, v, ax=0.0, ay=0.0, free_ax=False, free_ay=False, cax=0.0, cay=0.0, lam=None, obj_factor=1.0):
    """Steady-state NLP of the kart at x = [w_axle, z, phi, mu, psi, delta, ax, ay] (oracle/steady.hpp)."""
    x, params = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, params))
    spec = np.array([v, ax, ay, float(free_ax), float(free_ay), cax, cay])
    lam = np.zeros(12) if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
    f = C.c_double()
    c, grad, jac, hess = np.zeros(12), np.zeros(8), np.zeros((12, 8)), np.zeros((8, 8))
    lib().orc_kart_steady(_dp(x), _dp(params), _dp(spec), _dp(lam), C.c_double(obj_factor), C.byref(f), _dp(c), _dp(grad), _dp(jac), _dp(hess))
    return dict(f=f.value, c=c, grad=grad, jac=jac, hess=hess)


def f1_steady(x, params, v, ax=0.0, ay=0.0, free_ax=False, free_ay=False, cax=0.0, cay=0.0, lam=None, obj_factor=1.0):
    """Steady-state NLP of the F1 at x = [kappa x4, psi/10deg, Fz x4, delta/10deg, throttle, ax, ay] (ax, ay in g; oracle/steady.hpp)."""
    x, params = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, params))
    spec = np.array([v, ax, ay, float(free_ax), float(free_ay), cax, cay])
    lam = np.zeros(15) if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
    f = C.c_double()
    c, grad, jac, hess = np.zeros(15), np.zeros(13), np.zeros((15, 13)), np.zeros((13, 13))
    lib().orc_f1_steady(_dp(x), _dp(params), _dp(spec), _dp(lam), C.c_double(obj_factor), C.byref(f), _dp(c), _dp(grad), _dp(jac), _dp(hess))
    return dict(f=f.value, c=c, grad=grad, jac=jac, hess=hess)
