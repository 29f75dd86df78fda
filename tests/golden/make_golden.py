"""Generates the committed fixtures of tests/golden/ from the reference's own data files (run in the dev container only).

    python tests/golden/make_golden.py

Reads, with this repository's XML loaders, the vehicle / track databases and the golden solution files that the
reference's GoogleTest suites compare against (src/test/applications/f1_optimal_laptime_test.cpp,
optimal_laptime_test.cpp) and stores, per case, one compressed .npz holding the LapProblem arrays plus the golden NLP
point (x, lambda, zl, zu, lap time, iteration count).  /root/reference is not available on the GPU box, these files are.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from fastest_lap_b200 import vehicle_from_xml, track_from_xml, MODEL_F1, MODEL_KART      # noqa: E402
from fastest_lap_b200.problem import LapProblem, FORM_DIRECT, FORM_DERIVATIVE            # noqa: E402
from fastest_lap_b200.solution import solution_from_xml                                  # noqa: E402
from fastest_lap_b200.track import track_to_npz, TrackByArcs                                          # noqa: E402

REF = "/root/reference/"
DATA = REF + "src/test/applications/data/"


def golden_case(name, vehicle, track, sol, dissipation, **kw):
    form = FORM_DIRECT if sol.is_direct else FORM_DERIVATIVE
    q0 = u0 = du0 = None
    if not sol.is_closed:
        q0 = sol.inputs[0]
        u0 = np.array(LapProblem(vehicle.model, form, True, sol.s, track.track_length, track.node_data(sol.s), track.left_track_limit(sol.s),
                                 track.right_track_limit(sol.s), vehicle.params).ctrl_default)
        fm = sol.full_mesh_controls()
        for c in fm:
            u0[c] = sol.controls[c].values[0]
        if not sol.is_direct:
            du0 = np.array([sol.controls[c].derivatives[0] for c in fm])
    p = LapProblem.from_vehicle_track(vehicle, track, sol.s, form=form, closed=sol.is_closed, dissipation=dissipation, q0=q0, u0=u0, du0=du0, **kw)
    x = sol.pack_x()
    assert x.size == p.n and sol.lam.size == p.m, (x.size, p.n, sol.lam.size, p.m)
    # The flat-track F1 goldens predate a sign change of the roll-moment equation (second algebraic row, offset 10 of 19/21):
    # with the stored multiplier of that row negated, grad f + J^T lambda - zl + zu vanishes to 1.5e-10 on every variable
    # (1.3 otherwise).  The tests of the reference never compare multipliers.  `lam` is the sign-corrected vector, `lam_raw` the file's.
    lam = sol.lam.copy()
    if vehicle.model == MODEL_F1 and not p.elevation:
        lam.reshape(p.n_elements, p.ncpe)[:, 10] *= -1.0
    p.save_npz(os.path.join(HERE, name + ".npz"), x=x, lam=lam, lam_raw=sol.lam, zl=sol.zl, zu=sol.zu, laptime=sol.laptime, iter_count=sol.iter_count,
               time=sol.inputs[:, sol.inputs.shape[1] - 3])
    print(name, "n =", p.n, "m =", p.m, "laptime", sol.laptime)


def integral_case(name, vehicle, track, sol, dissipation, quantity, lower, upper):
    """A golden run with ONE restricted integral quantity (Options::integral_quantities, f1_optimal_laptime_test.cpp:1057-1117): the file
    holds the reference's variables (15 per node) and its multipliers -- 19 per element, then the multiplier of the integral's row.
    Stored as they are (x_ref, lam_ref, nu); tests map them to the node-wise formulation of the integral (tests/helpers.py)."""
    p = LapProblem.from_vehicle_track(vehicle, track, sol.s, form=FORM_DIRECT, closed=True, dissipation=dissipation)
    x = sol.pack_x()
    assert x.size == p.n and sol.lam.size == p.m + 1, (x.size, p.n, sol.lam.size, p.m)
    lam = sol.lam[:p.m].copy()
    lam.reshape(p.n_elements, p.ncpe)[:, 10] *= -1.0          # (the sign change of the roll-moment row, see golden_case)
    pa = p.with_integral_constraint(quantity, lower, upper)
    pa.save_npz(os.path.join(HERE, name + ".npz"), x_ref=x, lam_ref=lam, nu=sol.lam[p.m], zl_ref=sol.zl, zu_ref=sol.zu, laptime=sol.laptime,
                iter_count=sol.iter_count)
    print(name, "n =", pa.n, "m =", pa.m, "laptime", sol.laptime, "multiplier of the integral's row", sol.lam[p.m])


def trajectory_case(name, track, sol):
    """x, y, psi and time the reference stored with a solution on a 3-D track, next to the centreline the track gives at
    the mesh nodes: pins api.trajectory (position = centreline + n * normal vector, road_curvilinear.hpp:17-18)."""
    np.savez_compressed(os.path.join(HERE, name + ".npz"), s=sol.s, track_nodes=track.node_data(sol.s), centreline=track.position(sol.s),
                        track_length=track.track_length, inputs=sol.inputs, x_coord=sol.x_coord, y_coord=sol.y_coord, psi=sol.psi, laptime=sol.laptime)
    print(name, "points", len(sol.s))


def main():
    f1 = vehicle_from_xml(REF + "database/vehicles/f1/limebeer-2014-f1.xml")
    kart = vehicle_from_xml(REF + "database/vehicles/kart/roberto-lot-kart-2016.xml")
    rental = vehicle_from_xml(REF + "database/vehicles/kart/rental-kart.xml")
    catalunya = track_from_xml(REF + "database/tracks/catalunya/catalunya.xml")
    catalunya_3d = track_from_xml(REF + "database/tracks/catalunya/catalunya_3d.xml")
    catalunya_adapted = track_from_xml(REF + "database/tracks/catalunya/catalunya_adapted.xml")
    vendrell = track_from_xml(REF + "database/tracks/vendrell/vendrell.xml")

    # f1_optimal_laptime_test.cpp:168-213 (Ixx = Iyy = 0)
    v = f1.copy()
    v.set_parameter("vehicle/chassis/inertia/Ixx", 0.0)
    v.set_parameter("vehicle/chassis/inertia/Iyy", 0.0)
    golden_case("f1_catalunya_discrete", v, catalunya, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_discrete.xml", MODEL_F1), (50.0, 20.0 * 8.0e-4))
    # f1_optimal_laptime_test.cpp:1296-1340
    golden_case("f1_catalunya_3d", f1, catalunya_3d, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_3d.xml", MODEL_F1), (50.0, 50.0 * 8.0e-4))
    # f1_optimal_laptime_test.cpp:318-395 and :782-852 (open sections, wheel inertia 1.00 / 1.55)
    v = f1.copy()
    v.set_parameter("vehicle/front-axle/inertia", 1.00)
    v.set_parameter("vehicle/rear-axle/inertia", 1.55)
    golden_case("f1_catalunya_chicane", v, catalunya_adapted, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_chicane.xml", MODEL_F1), (50.0, 20.0 * 8.0e-4))
    golden_case("f1_catalunya_chicane_derivative", v, catalunya_adapted,
                solution_from_xml(DATA + "f1_optimal_laptime_catalunya_chicane_derivative.xml", MODEL_F1), (2.0e-1, 2.0e-5))
    # optimal_laptime_test.cpp:516-572
    golden_case("kart_vendrell", rental, vendrell, solution_from_xml(DATA + "vendrell_optimal.xml", MODEL_KART), (1.0e-2, 200 * 200.0 * 1.0e-10))

    # straight-and-corner tracks (Track_by_arcs): f1_optimal_laptime_test.cpp:694-733 (closed, 100 points) and :655-692 (open,
    # 401 points), dissipations 1e2 / 2e-3; optimal_laptime_test.cpp:283-334: the kart with direct torque, derivative form,
    # oval scaled by 0.2, dissipations 1e-3 / 2e-9 -- the golden that pins the kart's direct-torque variants
    oval = TrackByArcs(DATA + "ovaltrack.xml", 1.0, True)
    golden_case("f1_ovaltrack_closed", f1, oval, solution_from_xml(DATA + "f1_ovaltrack_closed.xml", MODEL_F1), (1.0e2, 2.0e-3))
    golden_case("f1_ovaltrack_open", f1, oval, solution_from_xml(DATA + "f1_ovaltrack_open.xml", MODEL_F1), (1.0e2, 2.0e-3))
    oval_kart = TrackByArcs(DATA + "ovaltrack.xml", 0.2, True)
    golden_case("kart_ovaltrack_derivative", kart, oval_kart, solution_from_xml(DATA + "ovaltrack_derivative.xml", MODEL_KART), (1.0e-3, 2.0e-9),
                kart_direct_torque=True)

    # optimal_laptime_test.cpp:337-459: the kart on Catalunya "by arcs" scaled by 0.2, 500 points: direct form with direct
    # torque (pins kart_direct_torque), derivative form with direct torque, derivative form with the throttle law
    cat_arcs = TrackByArcs(DATA + "catalunya.xml", 0.2, True)
    golden_case("kart_catalunya_direct", kart, cat_arcs, solution_from_xml(DATA + "optimal_laptime_catalunya_discrete.xml", MODEL_KART), (5.0, 8.0e-6),
                kart_direct_torque=True)
    golden_case("kart_catalunya_derivative", kart, cat_arcs, solution_from_xml(DATA + "catalunya_derivative.xml", MODEL_KART), (1.0e-2, 1.0e-10),
                kart_direct_torque=True)
    golden_case("kart_catalunya_derivative_throttle", kart, cat_arcs, solution_from_xml(DATA + "catalunya_derivative_throttle.xml", MODEL_KART),
                (1.0e-2, 200 * 200.0 * 1.0e-10))
    # f1_optimal_laptime_test.cpp:215-258: non-uniform ("adapted") mesh of the track file
    golden_case("f1_catalunya_adapted", f1, catalunya_adapted, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_adapted.xml", MODEL_F1), (50.0, 20.0 * 8.0e-4))
    # f1_optimal_laptime_test.cpp:736-778 and :857-906: ferrari-2022-australia on Melbourne, 700 points, both forms
    ferrari = vehicle_from_xml(REF + "database/vehicles/f1/ferrari-2022-australia.xml")
    melbourne = track_from_xml(REF + "database/tracks/melbourne/melbourne.xml")
    golden_case("f1_melbourne_direct", ferrari, melbourne, solution_from_xml(DATA + "f1_optimal_laptime_melbourne_700_direct.xml", MODEL_F1), (50.0, 20.0 * 8.0e-4))
    golden_case("f1_melbourne_derivative", ferrari, melbourne, solution_from_xml(DATA + "f1_optimal_laptime_melbourne_700_derivative.xml", MODEL_F1), (1.0e-1, 1.0e-5))

    # f1_optimal_laptime_test.cpp:1249-1292 and :1343-1388: two more 3-D tracks, each on the mesh of its track file
    # (the stored Laguna Seca run predates the used="1" kerbs of today's track file: with the limits widened by them the
    # optimum is 0.56 s faster, 58.7284 s; without them the stored 59.287501 s is reproduced to 4e-10 from a cold start)
    for nm, trk in (("f1_catalunya_2022_3d", "catalunya_2022/catalunya_2022_3d.xml"), ("f1_laguna_seca_3d", "laguna_seca/laguna_seca_3d.xml")):
        golden_case(nm, f1, track_from_xml(REF + "database/tracks/" + trk, use_kerbs=False), solution_from_xml(DATA + "f1_optimal_laptime_%s.xml" % nm[3:], MODEL_F1), (50.0, 50.0 * 8.0e-4))

    # f1_optimal_laptime_test.cpp:1057-1117: engine energy limited to 24 MJ (smooth throttle coefficient 1e-2, no boost power,
    # throttle dissipation 8e-4)
    v = f1.copy()
    v.set_parameter("vehicle/rear-axle/smooth_throttle_coeff", 1.0e-2)
    v.set_parameter("vehicle/rear-axle/boost/maximum-power", 0.0)
    integral_case("f1_catalunya_engine_energy", v, catalunya, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_engine_energy_limited.xml", MODEL_F1),
                  (50.0, 8.0e-4), "engine-energy", 0.0, 24.0)

    trajectory_case("f1_catalunya_3d_trajectory", catalunya_3d, solution_from_xml(DATA + "f1_optimal_laptime_catalunya_3d.xml", MODEL_F1))

    # input data of the benchmark configurations: tracks and vehicle parameter vectors
    data_dir = os.path.join(HERE, "..", "..", "fastest_lap_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    track_to_npz(catalunya, os.path.join(data_dir, "catalunya.npz"))
    track_to_npz(vendrell, os.path.join(data_dir, "vendrell.npz"))
    np.savez(os.path.join(data_dir, "vehicles.npz"), limebeer_2014_f1=f1.params, roberto_lot_kart_2016=kart.params, rental_kart=rental.params)




def steady_state_golden():
    """src/test/applications/data/steady_state.xml (steady_state_ad_test.cpp / steady_state_test.cpp): converged kart steady
    states -- 0 g, maximum lateral acceleration, maximum / minimum longitudinal acceleration at given lateral ones -- as
    x = [w_axle, z, phi, mu, psi, delta, ax, ay] of the 8-variable steady-state NLP (steady_state.hpp:899)."""
    import xml.etree.ElementTree as ET
    root = ET.fromstring(open(DATA + "steady_state.xml").read())
    names, rows = [], []
    for vel in root:
        v = float(vel.tag.split("_")[1]) / 3.6
        for case in vel:
            q = np.array([float(t) for t in case.find("q").text.split(",")])
            u = np.array([float(t) for t in case.find("u").text.split(",")])
            ax = float(case.find("ax").text) if case.find("ax") is not None else 0.0
            ay = q[3] * v
            names.append(vel.tag + "/" + case.tag)
            rows.append([v, (q[0] * 0.139 - v) / v, q[4], q[5], q[6], q[12], u[0], ax, ay])
    kart = vehicle_from_xml(REF + "database/vehicles/kart/roberto-lot-kart-2016.xml")
    np.savez_compressed(os.path.join(HERE, "kart_steady_state.npz"), names=np.array(names), v=np.array(rows)[:, 0], x=np.array(rows)[:, 1:], params=kart.params)
    print("kart_steady_state", len(names), "points")


def f1_steady_state_golden():
    """src/test/applications/data/f1_steady_state_test.xml (f1_steady_state_test.cpp, tolerance 2e-4): maximum lateral
    acceleration over speed and the G-G diagrams at 100 ... 300 km/h of limebeer-2014-f1."""
    import xml.etree.ElementTree as ET
    root = ET.fromstring(open(DATA + "f1_steady_state_test.xml").read())
    vec = lambda p: np.array([float(t) for t in root.find(p).text.split(",")])
    d = dict(params=vehicle_from_xml(REF + "database/vehicles/f1/limebeer-2014-f1.xml").params,
             v=vec("max_lateral_acceleration/velocity"), ay_max=vec("max_lateral_acceleration/acceleration-y"), ax_at_ay_max=vec("max_lateral_acceleration/acceleration-x"))
    for k in (100, 150, 200, 250, 300):
        for q, nm in (("maximum_x_acceleration", "ax_max"), ("minimum_x_acceleration", "ax_min"), ("y_acceleration", "ay")):
            d["gg%d_%s" % (k, nm)] = vec("gg_diagram_%d/%s" % (k, q))
    np.savez_compressed(os.path.join(HERE, "f1_steady_state.npz"), **d)
    print("f1_steady_state", len(d["v"]), "speeds")


if __name__ == "__main__":
    main()
    steady_state_golden()
    f1_steady_state_golden()
