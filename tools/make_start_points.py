"""Generates fastest_lap_b200/data/f1_steady_state.npz: straight-line steady states (0 g) of limebeer-2014-f1 at a few speeds,
the reference's starting point of an optimal-lap computation (fastestlapc.cpp:1508-1545: Steady_state(car).solve(v, 0, 0)
replicated over the mesh).  Uses the CPU oracle's node model; run in the dev container, the result is committed data."""
import os, sys
import numpy as np
import scipy.optimize as so
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc

params = np.load(os.path.join(ROOT, "fastest_lap_b200", "data", "vehicles.npz"))["limebeer_2014_f1"]


def steady(v, z0):
    def build(z):
        q = np.zeros(14); u = np.zeros(4)
        q[0] = q[1] = z[0]; q[2] = q[3] = z[1]; q[4] = v; q[7] = q[8] = z[2]; q[9] = q[10] = z[3]
        u[2] = z[4]; u[3] = params[18]
        return q, u

    def res(z):
        d = orc.f1_car(*build(z), params)["dstates"]
        return np.array([d[0], d[2], d[4], d[7], d[9]])      # wheels, longitudinal, vertical and pitch balance
    sol = so.fsolve(res, z0, xtol=1e-13)
    assert np.max(np.abs(res(sol))) < 1e-9, (v, res(sol))
    q, u = build(sol)
    return np.concatenate([q[:11], q[12:14], [u[0], u[2]]]), sol      # the 15 variables of a direct-form node


speeds = np.array([50.0, 100.0, 150.0, 200.0, 250.0])
nodes, z = [], np.array([0.0, 0.0, -0.3, -0.2, 0.01])
for v in speeds:
    node, z = steady(v / 3.6, z)
    nodes.append(node)
nodes = np.array(nodes)
np.savez(os.path.join(ROOT, "fastest_lap_b200", "data", "f1_steady_state.npz"), speeds_kmh=speeds, nodes=nodes)
print(nodes)
