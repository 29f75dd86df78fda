"""Shared test helpers: the CPU oracle behind the same LapProblem description the CUDA path takes."""
import os
import numpy as np

from fastest_lap_b200.problem import LapProblem
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return LapProblem.load_npz(os.path.join(GOLDEN, name + ".npz"))


def oracle_nlp(p):
    return orc.OracleNLP(p.model, p.form, p.closed, p.s, p.track_length, p.track_nodes, p.params, p.dissipation, p.sigma, p.elevation,
                         ctrl_default=p.ctrl_default, q0=p.q0, u0=p.u0, du0=p.du0, kart_direct_torque=p.kart_direct_torque)


def equality_rows(p):
    """Row offsets (within an element) of the equality constraints: trapezoid + algebraic rows, control-rate rows."""
    n_eq = p.ncpe - 6 - (0 if p.form == 0 else 2)
    rows = list(range(n_eq))
    if p.form != 0:
        rows += [p.ncpe - 2, p.ncpe - 1]
    return np.array(rows)


def rel_err(a, b, floor=1.0):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)) if a.size else 0.0
