"""Accuracy of the branch-free elementary functions of the node programs (fastest_lap_b200/csrc/fl_math.cuh), host build,
against numpy's libm in units in the last place.  The same functions are exercised on the device by every GPU parity test."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def flm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("flm") / "flm.so")
    flags = ["-O2", "-shared", "-fPIC", "-std=c++17"]
    if "fma" in open("/proc/cpuinfo").read():
        flags.append("-mfma")
    subprocess.check_call(["g++"] + flags + [os.path.join(HERE, "flmath_shim.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.flm_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long]

    def ev(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib.flm_eval(fn, x.ctypes.data, y.ctypes.data, x.size)
        return y
    return ev


def _ulps(y, ref):
    return np.max(np.abs(y - ref) / np.spacing(np.abs(ref)))


def _samples(lo, hi, n=400001, seed=0):
    rng = np.random.default_rng(seed)
    return np.concatenate([np.linspace(lo, hi, n), rng.uniform(lo, hi, n)])


def test_reciprocal_and_square_root(flm):
    x = np.concatenate([_samples(1e-6, 10.0), _samples(10.0, 1e6), -_samples(1e-3, 1e3)])
    assert _ulps(flm(0, x), 1.0 / x) <= 1.0
    x = np.concatenate([_samples(1e-12, 1.0), _samples(1.0, 1e8)])
    assert _ulps(flm(1, x), np.sqrt(x)) <= 1.0
    assert flm(1, np.array([0.0]))[0] == 0.0


def test_sine_and_cosine(flm):
    x = np.concatenate([_samples(-8.0, 8.0), _samples(-1000.0, 1000.0)])
    # near the zeros of the function the error is absolute (2^-53 of the reduced argument's scale), as for any two-constant reduction
    for fn, ref in ((2, np.sin(x)), (3, np.cos(x))):
        err = np.abs(flm(fn, x) - ref)
        assert np.max(err / np.maximum(np.spacing(np.abs(ref)), 2.0 ** -56)) <= 2.0
    assert np.max(np.abs(flm(7, x) - 1.0)) <= 4e-16
    small = _samples(-1e-3, 1e-3)
    assert _ulps(flm(2, small[small != 0]), np.sin(small[small != 0])) <= 1.0


def test_arc_tangent(flm):
    x = np.concatenate([_samples(-3.0, 3.0), _samples(-1e4, 1e4), _samples(-1e-3, 1e-3)])
    x = x[x != 0]
    assert _ulps(flm(4, x), np.arctan(x)) <= 1.5
    for edge in (0.4375, 0.6875, 1.1875, 2.4375):
        e = edge + np.arange(-50, 51) * np.spacing(edge)
        assert _ulps(flm(4, e), np.arctan(e)) <= 1.5
    assert flm(4, np.array([0.0]))[0] == 0.0


def test_exponential_and_hyperbolic_tangent(flm):
    x = np.concatenate([_samples(-5.0, 5.0), _samples(-300.0, 300.0)])
    assert _ulps(flm(6, x), np.exp(x)) <= 1.5
    x = np.concatenate([_samples(-4.0, 4.0), _samples(-60.0, 60.0), _samples(-1e-4, 1e-4)])
    x = x[x != 0]
    assert _ulps(flm(5, x), np.tanh(x)) <= 3.5          # the division after expm1 adds to its 1.5 ulp
    assert flm(5, np.array([0.0]))[0] == 0.0 and flm(5, np.array([1e3]))[0] == 1.0
    for fn in range(7):          # a NaN argument stays a NaN (the solver's non-finite checks rely on it)
        assert np.isnan(flm(fn, np.array([np.nan]))[0]), fn


@pytest.mark.gpu
def test_device_build_has_the_same_accuracy():
    """The device build (seeds from MUFU.RCP64H / MUFU.RSQ64H) against numpy, same bounds as the host build."""
    from fastest_lap_b200.nlp import math_eval
    x = np.concatenate([_samples(1e-6, 10.0), _samples(10.0, 1e6), -_samples(1e-3, 1e3)])
    assert _ulps(math_eval(0, x), 1.0 / x) <= 1.0
    x = np.concatenate([_samples(1e-12, 1.0), _samples(1.0, 1e8)])
    assert _ulps(math_eval(1, x), np.sqrt(x)) <= 1.0
    assert math_eval(1, np.array([0.0]))[0] == 0.0
    x = np.concatenate([_samples(-8.0, 8.0), _samples(-1000.0, 1000.0)])
    for fn, ref in ((2, np.sin(x)), (3, np.cos(x))):
        assert np.max(np.abs(math_eval(fn, x) - ref) / np.maximum(np.spacing(np.abs(ref)), 2.0 ** -56)) <= 2.0
    x = np.concatenate([_samples(-3.0, 3.0), _samples(-1e4, 1e4), _samples(-1e-3, 1e-3)])
    x = x[x != 0]
    assert _ulps(math_eval(4, x), np.arctan(x)) <= 1.5
    x = np.concatenate([_samples(-4.0, 4.0), _samples(-60.0, 60.0), _samples(-1e-4, 1e-4)])
    x = x[x != 0]
    assert _ulps(math_eval(5, x), np.tanh(x)) <= 3.5
    x = np.concatenate([_samples(-5.0, 5.0), _samples(-300.0, 300.0)])
    assert _ulps(math_eval(6, x), np.exp(x)) <= 1.5
