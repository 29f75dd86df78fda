"""The C-ABI library builds, loads and exports every symbol include/fastestlap_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fastest_lap_b200 import build
    return ctypes.CDLL(build.build())


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fastestlap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_callbacks():
    names = declared_functions()
    for n in ("fl_nlp_create", "fl_eval_f", "fl_eval_grad_f", "fl_eval_g", "fl_eval_jac_g", "fl_eval_h", "fl_eval_all", "fl_nlp_structure", "fl_nlp_bounds"):
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    for n in declared_functions():
        assert hasattr(lib, n), n


def test_parameter_vector_sizes_match_the_xml_tables(lib):
    from fastest_lap_b200.vehicle import F1_PARAM_PATHS, KART_PARAM_PATHS
    assert lib.fl_n_vehicle_params(0) == len(F1_PARAM_PATHS)
    assert lib.fl_n_vehicle_params(1) == len(KART_PARAM_PATHS)
    assert lib.fl_n_vehicle_params(7) == -1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fastest_lap_b200.nlp import CollocationNLP
    from tests.helpers import load_golden
    p, _ = load_golden("f1_catalunya_chicane")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        CollocationNLP(p)


def test_callbacks_have_ipopts_types(tmp_path):
    """Compile-only: `Eval_F_CB f = fl_eval_f;` ... against the typedefs of Ipopt 3.14's IpStdCInterface.h (restated in
    tests/IpStdCInterface_3_14.h), in C, with incompatible pointer types as errors."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-Werror=incompatible-pointer-types", "-c", os.path.join(here, "ipopt_callbacks_check.c"),
                        "-o", str(tmp_path / "check.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
