"""Parity of the CUDA callbacks (through the C-ABI, host buffers) with the CPU oracle.

Tolerance (BASELINE.json north_star): callback values and derivative entries within 1e-10 relative, fp64 -- Jacobian AND
Hessian entries, at golden points, perturbed points, full-size meshes and with N(0,1) multipliers.  "Relative" is taken against
max(|reference entry|, 1) for f, g and grad f, and against max(|reference entry|, 0.1 * largest entry of the same block) for
matrix entries, the block being the node's Hessian block / the element's Jacobian rows: the floor is there because some small
entries are ill-conditioned in fp64 whichever way they are differentiated (e.g. d2/dkappa2 of kappa/rho = lambda_n^2/rho^3
evaluated as 1/rho - kappa_n^2/rho^3 on a straight).  The reference for second derivatives is the oracle's own derivative mode
(dense second-order jets over its restatement of the reference's formulas); where its double evaluation could itself be the
limit, the same formulas in extended precision arbitrate (_hessian_close), and test_ill_conditioned_entries_against_40_digit_truth
puts the rounding of the CUDA code's expression graph next to a 40-digit evaluation.  Entries are compared by (row, col) with
duplicates summed -- never by position.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from tests.helpers import load_golden, oracle_nlp, rel_err

pytestmark = pytest.mark.gpu

CASES = ["f1_catalunya_discrete", "f1_catalunya_3d", "f1_catalunya_chicane", "f1_catalunya_chicane_derivative", "kart_vendrell",
         "f1_ovaltrack_closed", "f1_ovaltrack_open", "kart_ovaltrack_derivative", "kart_catalunya_direct", "kart_catalunya_derivative",
         "kart_catalunya_derivative_throttle", "f1_catalunya_adapted", "f1_melbourne_direct", "f1_melbourne_derivative", "f1_catalunya_2022_3d", "f1_laguna_seca_3d",
         "f1_catalunya_engine_energy", "f1_catalunya_hypermesh"]
RTOL = 1.0e-10


def _csr(rows, cols, vals, shape):
    return sp.coo_matrix((vals, (rows, cols)), shape=shape).tocsr()


def _matrix_close(A, B, what, block, rtol=RTOL):
    """A: candidate, B: reference; both CSR with duplicates already summed; `block` = rows per block."""
    D = (A - B).tocoo()
    if D.nnz == 0:
        return
    Babs = abs(B)
    Bc = Babs.tocoo()
    S = np.zeros(B.shape[0] // block + 1)
    np.maximum.at(S, Bc.row // block, Bc.data)
    ref = np.asarray(Babs[D.row, D.col]).ravel()
    scale = np.maximum(ref, 0.1 * S[D.row // block])
    scale[scale == 0.0] = 1.0
    err = np.abs(D.data) / scale
    k = int(np.argmax(err))
    assert err[k] <= rtol, "%s: entry (%d, %d) differs by %.3e relative (ref %.6e)" % (what, D.row[k], D.col[k], err[k], ref[k])


def _hessian_close(p, H, x, lam, obj, ref_hess, what="hessian"):
    """north_star's 1e-10 on every Hessian entry (relative to max(|entry|, 0.1 x the largest entry of its block)) against the
    oracle.  Should fp64 cancellation in the oracle's own double evaluation ever exceed that on an entry, the same formulas
    evaluated in extended precision (oracle/jet.hpp, liboracle_ld.so) arbitrate: the CUDA value must then be within 1e-10 of THAT."""
    A, B = _csr(*H, (p.n, p.n)), _csr(*ref_hess, (p.n, p.n))
    try:
        _matrix_close(A, B, what, p.nv)
    except AssertionError:
        _matrix_close(A, _csr(*oracle_nlp(p, extended=True).eval(x, lam, obj)["hess"], (p.n, p.n)), what + " (extended-precision oracle)", p.nv)


def _points(p, ex, count=3, seed=0):
    """The golden point plus seeded perturbations of it (x (1 + 0.01 N) + 1e-3 N) with multipliers of the
    size a solver carries: lambda_golden (1 + 0.05 N), obj_factor in (0.1, 1)."""
    rng = np.random.default_rng(seed)
    pts = [(ex["x"], ex["lam"], 1.0)]
    for _ in range(count - 1):
        x = ex["x"] * (1.0 + 0.01 * rng.standard_normal(p.n)) + 1.0e-3 * rng.standard_normal(p.n)
        pts.append((x, ex["lam"] * (1.0 + 0.05 * rng.standard_normal(p.m)), float(rng.uniform(0.1, 1.0))))
    return pts


def _stress_point(p, ex, seed=11):
    """Multipliers ~ N(0,1) on every row: O(1e6) second-derivative terms of the tyre model then cancel to O(10) in the
    (steering, steering) entry, so fp64 -- oracle and CUDA alike -- keeps about 8 digits there."""
    rng = np.random.default_rng(seed)
    x = ex["x"] * (1.0 + 0.01 * rng.standard_normal(p.n)) + 1.0e-3 * rng.standard_normal(p.n)
    return x, rng.standard_normal(p.m), float(rng.uniform(0.1, 1.0))


@pytest.fixture(scope="module", params=CASES)
def case(request):
    from fastest_lap_b200.nlp import CollocationNLP
    p, ex = load_golden(request.param)
    nlp = CollocationNLP(p)
    yield p, ex, nlp, oracle_nlp(p)
    nlp.close()


def test_sizes_match_reference_layout(case):
    p, ex, nlp, o = case
    assert (nlp.n, nlp.m) == (p.n, p.m) == (o.n, o.m)
    jr, jc, hr, hc = nlp.structure()
    assert jr.min() >= 0 and jr.max() < p.m and jc.min() >= 0 and jc.max() < p.n
    assert np.all(hr >= hc) and hr.max() < p.n            # lower triangle


def test_callbacks_match_oracle(case):
    p, ex, nlp, o = case
    jr, jc, hr, hc = nlp.structure()
    for x, lam, obj in _points(p, ex):
        ref = o.eval(x, lam, obj)
        out = nlp.eval_all(x, lam, obj)
        assert abs(out["f"][0] - ref["f"]) <= RTOL * max(1.0, abs(ref["f"]))
        assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(np.abs(ref["g"]), 1.0)) <= RTOL
        assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(np.abs(ref["grad"]), 1.0)) <= RTOL
        _matrix_close(_csr(jr, jc, out["jac"][0], (p.m, p.n)), _csr(*ref["jac"], (p.m, p.n)), "jacobian", p.ncpe)
        _hessian_close(p, (hr, hc, out["hess"][0]), x, lam, obj, ref["hess"])


def test_output_channels_match_oracle(case):
    """The model's output channels along the lap (get_outputs_map of the reference: tyre slips, forces, dissipation, wheel speeds,
    contact-point velocities: tire.h:170-196; chassis accelerations, resultant force, aerodynamic forces: chassis.h:282-308,
    chassis.hpp:228-244) at the golden point and at a perturbed one, against the oracle's own evaluation: 1e-10 relative to
    max(|value|, largest value of the channel along the lap) -- with a floor of 1 N / 1 W for forces and powers: the vertical resultant
    force of a converged flat lap is the residual of an equation, 1e-12 N of rounding on either side."""
    p, ex, nlp, o = case
    names = nlp.output_names
    assert len(names) == 45 and names[0] == "front-axle.left-tire.lambda" and names[36] == "chassis.acceleration.x"
    rng = np.random.default_rng(4)
    for x in (ex["x"], ex["x"] * (1.0 + 0.01 * rng.standard_normal(p.n))):
        out = nlp.eval_outputs(x)
        ref = o.outputs(x)
        for k, nm in enumerate(names):
            floor = 1.0 if (".force." in nm or "aerodynamics" in nm or "dissipation" in nm) else 1e-3      # (slip of the kart's front tyres: exactly 0 here, 2e-16 of rounding in (omega R0 - v) / v there)
            scale = np.maximum(np.abs(ref[k]), max(np.abs(ref[k]).max(), floor))
            assert np.max(np.abs(out[nm][0] - ref[k]) / scale) <= RTOL, nm


def test_stress_point_random_multipliers(case):
    """N(0,1) multipliers: values, gradient, Jacobian and Hessian entries all within 1e-10 of the oracle
    (measured: <= 2.1e-11 against the oracle evaluated in extended precision, tools/hess_arbiter.py)."""
    p, ex, nlp, o = case
    jr, jc, hr, hc = nlp.structure()
    x, lam, obj = _stress_point(p, ex)
    ref = o.eval(x, lam, obj)
    out = nlp.eval_all(x, lam, obj)
    assert abs(out["f"][0] - ref["f"]) <= RTOL * max(1.0, abs(ref["f"]))
    assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(np.abs(ref["g"]), 1.0)) <= RTOL
    assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(np.abs(ref["grad"]), 1.0)) <= RTOL
    _matrix_close(_csr(jr, jc, out["jac"][0], (p.m, p.n)), _csr(*ref["jac"], (p.m, p.n)), "jacobian", p.ncpe)
    _hessian_close(p, (hr, hc, out["hess"][0]), x, lam, obj, ref["hess"])


def test_ill_conditioned_entries_against_40_digit_truth(case):
    """Where CUDA and oracle differ by more than 1e-10 of the entry, neither is wrong: both are fp64 evaluations of an
    ill-conditioned expression.  Against a 40-digit evaluation of the same node, at the stress point, the CUDA Hessian
    entries must be within 1e-8 relative and no further from the truth than 50x the oracle's own error (+1e-12 of the block
    scale); Jacobian entries within 1e-12 of the block scale; gradient entries within 1e-12."""
    pytest.importorskip("mpmath")
    from tests.helpers import truth_node
    p, ex, nlp, o = case
    if getattr(p, "aug", None):
        pytest.skip("the 40-digit evaluation covers the reference's own node layout")
    x, lam, obj = _stress_point(p, ex)
    out = nlp.eval_all(x, lam, obj)
    ref = o.eval(x, lam, obj)
    jr, jc, hr, hc = nlp.structure()
    H = _csr(hr, hc, out["hess"][0], (p.n, p.n))
    Ho = _csr(*ref["hess"], (p.n, p.n))
    J = _csr(jr, jc, out["jac"][0], (p.m, p.n))
    Jo = _csr(*ref["jac"], (p.m, p.n))
    node0 = 0 if p.closed else 1
    rng = np.random.default_rng(5)
    for node in rng.choice(np.arange(node0, p.n_points), 6, replace=False):
        t = truth_node(p, x, lam, obj, int(node))
        v0 = (node - node0) * p.nv
        S = max(abs(v) for v in t["hess"].values())
        for (r, c), tv in t["hess"].items():
            e_gpu, e_orc = abs(H[v0 + r, v0 + c] - tv), abs(Ho[v0 + r, v0 + c] - tv)
            assert e_gpu <= 1e-8 * max(abs(tv), 0.1 * S) and e_gpu <= 50 * e_orc + 1e-12 * S, (node, r, c, e_gpu, e_orc, tv)
        e_in = (node - 1) % p.n_points
        S = max(abs(v) for v in t["jacA"].values())
        for (r, c), tv in t["jacA"].items():
            e_gpu, e_orc = abs(J[e_in * p.ncpe + r, v0 + c] - tv), abs(Jo[e_in * p.ncpe + r, v0 + c] - tv)
            assert e_gpu <= 1e-12 * S and e_gpu <= 10 * e_orc + 1e-13 * S, (node, r, c, e_gpu, e_orc, tv)
        for (k, _), tv in t["grad"].items():
            assert abs(out["grad"][0][v0 + k] - tv) <= 1e-12 * max(1.0, abs(tv))


def test_ipopt_style_callbacks_agree_with_fused_call(case):
    p, ex, nlp, o = case
    x, lam, obj = _points(p, ex, 2, seed=3)[1]
    out = nlp.eval_all(x, lam, obj)
    assert nlp.eval_f(x, True) == out["f"][0]
    assert np.array_equal(nlp.eval_g(x, False), out["g"][0])
    # the jac-only / hess-only kernels run differently partitioned code than the fused one: same values up to rounding
    np.testing.assert_allclose(nlp.eval_grad_f(x, False), out["grad"][0], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(nlp.eval_jac_g(x, False), out["jac"][0], rtol=1e-12, atol=1e-12)
    h = nlp.eval_h(x, lam, obj, False)
    np.testing.assert_allclose(h, out["hess"][0], rtol=1e-9, atol=1e-10)


def test_golden_point_is_a_kkt_point_on_the_gpu(case):
    p, ex, nlp, o = case
    if getattr(p, "aug", None) and p.aug["kind"] == 1:
        pytest.skip("the hypermesh case is a feasible point of its problem, not a converged run")
    out = nlp.eval_all(ex["x"], ex["lam"], 1.0)
    jr, jc, hr, hc = nlp.structure()
    JTl = np.zeros(p.n)
    np.add.at(JTl, jc, out["jac"][0] * ex["lam"][jr])
    assert np.abs(out["grad"][0] + JTl - ex["zl"] + ex["zu"]).max() < 5.0e-10


def test_batch_of_parameter_sets_matches_single_problems():
    from fastest_lap_b200.nlp import CollocationNLP
    p, ex = load_golden("f1_catalunya_chicane")
    rng = np.random.default_rng(2)
    B = 5
    params = np.tile(p.params, (B, 1))
    params[:, 0] = rng.uniform(620, 760, B)        # mass
    params[:, 6] = rng.uniform(0.8, 1.1, B)        # cd
    params[:, 7] = rng.uniform(2.6, 3.4, B)        # cl
    params[:, 30] = rng.uniform(650, 800, B)       # engine power
    xs = np.stack([ex["x"] * (1 + 0.01 * rng.standard_normal(p.n)) for _ in range(B)])
    lams = rng.standard_normal((B, p.m))
    nlp = CollocationNLP(p, params)
    out = nlp.eval_all(xs, lams, 0.5)
    for b in (0, B - 1):
        pb = type(p)(p.model, p.form, p.closed, p.s, p.track_length, p.track_nodes, p.wl, p.wr, params[b], p.dissipation, p.sigma, p.elevation,
                     p.q0, p.u0, p.du0, None, p.kart_direct_torque)
        single = CollocationNLP(pb)
        ob = single.eval_all(xs[b], lams[b], 0.5)
        for k in ("f", "g", "grad", "jac", "hess"):
            # the batch reads its per-problem parameters from global memory, the single problem from the constant bank:
            # two compilations of the same straight-line code, equal up to the rounding of fused multiply-adds
            assert rel_err(out[k][b], ob[k][0]) <= 1e-12, k
        single.close()
    nlp.close()


def test_empty_and_invalid_inputs_are_rejected():
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.problem import LapProblem
    p, ex = load_golden("f1_catalunya_chicane")
    with pytest.raises(ValueError):
        LapProblem(p.model, p.form, True, p.s[:1], p.track_length, p.track_nodes[:1], p.wl[:1], p.wr[:1], p.params)
    bad = type(p)(p.model, p.form, p.closed, p.s, p.track_length, p.track_nodes, p.wl, p.wr, p.params, p.dissipation, p.sigma, p.elevation, p.q0, p.u0, p.du0)
    nlp = CollocationNLP(bad)
    with pytest.raises(AssertionError):
        nlp.eval_all(ex["x"][:-1], ex["lam"])
    nlp.close()


# ---- BASELINE.json's full sizes --------------------------------------------------------------------------------------
def test_full_size_f1_8000_nodes_against_oracle():
    """configs[2]: F1, Catalunya, 8000 nodes (n = 120 000, m = 152 000): every callback against the oracle at the golden
    solution interpolated to the mesh, and a size-independent property: refining the mesh leaves the lap time (the
    objective without the control penalty) of the interpolated solution within the trapezoid error of the 500-node one."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP
    p = workloads.f1_catalunya(8000)
    gp, ex = load_golden("f1_catalunya_discrete")
    x = workloads.resample_point(p, gp.s, ex["x"], p.nv)
    lam = np.random.default_rng(1).standard_normal(p.m)
    nlp = CollocationNLP(p)
    assert (nlp.n, nlp.m) == (120000, 152000)
    out = nlp.eval_all(x, lam, 1.0)
    ref = oracle_nlp(p).eval(x, lam, 1.0)
    jr, jc, hr, hc = nlp.structure()
    assert abs(out["f"][0] - ref["f"]) <= RTOL * abs(ref["f"])
    assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(np.abs(ref["g"]), 1.0)) <= RTOL
    assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(np.abs(ref["grad"]), 1.0)) <= RTOL
    _matrix_close(_csr(jr, jc, out["jac"][0], (p.m, p.n)), _csr(*ref["jac"], (p.m, p.n)), "jacobian", p.ncpe)
    _hessian_close(p, (hr, hc, out["hess"][0]), x, lam, 1.0, ref["hess"])
    from fastest_lap_b200.ipm import lap_time
    assert abs(lap_time(p, x, out["f"][0]) - float(ex["laptime"])) <= 5.0e-2       # 77.82 s vs 77.79 s: interpolation of a 500-node solution
    nlp.close()


def test_full_size_kart_2000_nodes_against_oracle():
    """configs[3]: kart, Vendrell, 2000 nodes, derivative form (n = 32 000, m = 40 000)."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP
    p = workloads.kart_vendrell(2000)
    gp, ex = load_golden("kart_vendrell")
    x = workloads.resample_point(p, gp.s, ex["x"], p.nv)
    lam = 0.1 * np.random.default_rng(1).standard_normal(p.m)
    nlp = CollocationNLP(p)
    assert (nlp.n, nlp.m) == (32000, 40000)
    out = nlp.eval_all(x, lam, 1.0)
    ref = oracle_nlp(p).eval(x, lam, 1.0)
    jr, jc, hr, hc = nlp.structure()
    assert abs(out["f"][0] - ref["f"]) <= RTOL * abs(ref["f"])
    assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(np.abs(ref["g"]), 1.0)) <= RTOL
    assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(np.abs(ref["grad"]), 1.0)) <= RTOL
    _matrix_close(_csr(jr, jc, out["jac"][0], (p.m, p.n)), _csr(*ref["jac"], (p.m, p.n)), "jacobian", p.ncpe)
    _hessian_close(p, (hr, hc, out["hess"][0]), x, lam, 1.0, ref["hess"])
    nlp.close()


def test_batch_linearity_in_the_multipliers():
    """configs[4]-style batch (64 laps with perturbed vehicles): the Lagrangian Hessian is linear in (obj_factor, lambda):
    H(a, lam1 + lam2) = H(a, lam1) + H(0, lam2), lap by lap -- a size-independent check of the batched derivative kernel."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.vehicle import MODEL_F1
    p, ex = load_golden("f1_catalunya_discrete")
    B = 64
    rng = np.random.default_rng(5)
    nlp = CollocationNLP(p, vehicle_params=workloads.sweep_parameters(MODEL_F1, B))
    x = ex["x"][None, :] * (1 + 1e-3 * rng.standard_normal((B, 1)))
    l1, l2 = rng.standard_normal((B, p.m)), rng.standard_normal((B, p.m))
    h12 = nlp.eval_all(x, l1 + l2, 0.7, want=("hess",))["hess"].copy()
    h1 = nlp.eval_all(x, l1, 0.7, want=("hess",))["hess"].copy()
    h2 = nlp.eval_all(x, l2, 0.0, want=("hess",))["hess"].copy()
    scale = np.maximum(np.abs(h1) + np.abs(h2), 1.0)
    assert np.max(np.abs(h12 - h1 - h2) / scale) <= 1e-9       # entries are sums of ~40 products of mixed sign: fp64 cancellation
    nlp.close()


# ---- the two generated variants without a golden of their own ---------------------------------------------------------
@pytest.mark.parametrize("name, form", [("kart_vendrell", 0), ("kart_catalunya_direct", 1), ("f1_catalunya_3d", 1), ("f1_laguna_seca_3d", 1)])
def test_variants_without_a_golden_match_the_oracle(name, form):
    """`kart_direct` (throttle law, direct form), `f1_derivative_3d` and, for good measure, the other form of two more
    goldens: the problem of a golden rebuilt in the OTHER form (control rates dropped from / appended to the stored point),
    every callback against the oracle at perturbed points.  The oracle's form handling is the one its pinned cases use."""
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.problem import FORM_DIRECT
    p, ex = load_golden(name)
    assert p.form != form
    n_state = p.nv - (2 if p.form == FORM_DIRECT else 4)
    X = ex["x"].reshape(-1, p.nv)
    X = X[:, :n_state + 2] if form == FORM_DIRECT else np.hstack([X, np.zeros((X.shape[0], 2))])
    diss = (5.0, 8.0e-6) if (form == FORM_DIRECT and p.model == 1) else ((50.0, 1.6e-2) if form == FORM_DIRECT else (1.0e-1, 1.0e-5))
    q = type(p)(p.model, form, p.closed, p.s, p.track_length, p.track_nodes, p.wl, p.wr, p.params, diss, p.sigma, p.elevation,
                p.q0, p.u0, p.du0, p.ctrl_default, p.kart_direct_torque)
    nlp, o = CollocationNLP(q), oracle_nlp(q)
    expected = {("kart_vendrell", 0): "kart_direct", ("kart_catalunya_direct", 1): "kart_derivative_torque",
                ("f1_catalunya_3d", 1): "f1_derivative_3d", ("f1_laguna_seca_3d", 1): "f1_derivative_3d"}[(name, form)]
    assert nlp.variant == expected
    jr, jc, hr, hc = nlp.structure()
    rng = np.random.default_rng(11)
    x0 = np.ascontiguousarray(X).ravel()
    for _ in range(3):
        x = x0 * (1.0 + 0.01 * rng.standard_normal(q.n)) + 1.0e-4 * rng.standard_normal(q.n)
        lam, obj = rng.standard_normal(q.m) * 0.1, float(rng.uniform(0.1, 1.0))
        ref, out = o.eval(x, lam, obj), nlp.eval_all(x, lam, obj)
        assert abs(out["f"][0] - ref["f"]) <= RTOL * max(1.0, abs(ref["f"]))
        assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(np.abs(ref["g"]), 1.0)) <= RTOL
        assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(np.abs(ref["grad"]), 1.0)) <= RTOL
        _matrix_close(_csr(jr, jc, out["jac"][0], (q.m, q.n)), _csr(*ref["jac"], (q.m, q.n)), "jacobian", q.ncpe)
        _hessian_close(q, (hr, hc, out["hess"][0]), x, lam, obj, ref["hess"])
    nlp.close()


def test_parameters_that_vary_along_the_lap():
    """Model_parameter (src/core/foundation/model_parameter.h:51-77; evaluated per node, dynamic_model_car.hpp:62-63): mass, lift
    and engine power declared as piecewise-linear functions of the arclength.  The kernels read one parameter vector per node
    (fl_problem_desc::node_params): f, g, grad f, Jacobian and Hessian equal the oracle's, which evaluates every node with
    that node's parameters; with the same vector at every node the result is bit-identical to the constant-parameter path."""
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.problem import LapProblem
    from fastest_lap_b200.vehicle import Vehicle, MODEL_F1
    p0, ex = load_golden("f1_catalunya_discrete")
    veh = Vehicle(MODEL_F1, p0.params.copy())
    L = p0.track_length
    veh.declare_variable_parameter("vehicle/chassis/mass", [660.0, 700.0, 640.0], [0, 1, 2, 0], [0.0, 0.3 * L, 0.6 * L, L])
    veh.declare_variable_parameter("vehicle/chassis/aerodynamics/cl", [3.0, 2.6], [0, 1], [0.2 * L, 0.8 * L])
    veh.declare_variable_parameter("vehicle/rear-axle/engine/maximum-power", [735.0, 600.0], [0, 1], [0.5 * L, 0.51 * L])
    node_params = veh.params_at(p0.s)
    assert np.ptp(node_params[:, 0]) > 50.0
    p = LapProblem(p0.model, p0.form, p0.closed, p0.s, p0.track_length, p0.track_nodes, p0.wl, p0.wr, node_params[0], p0.dissipation, p0.sigma,
                   p0.elevation, node_params=node_params)
    rng = np.random.default_rng(5)
    x = ex["x"] * (1 + 1e-3 * rng.standard_normal(p.n))
    lam = rng.standard_normal(p.m)
    nlp = CollocationNLP(p)
    out = nlp.eval_all(x, lam, 0.9)
    ref = oracle_nlp(p).eval(x, lam, 0.9)
    jr, jc, hr, hc = nlp.structure()
    assert abs(out["f"][0] - ref["f"]) <= 1e-10 * abs(ref["f"])
    assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(1.0, np.abs(ref["g"]))) <= 1e-10
    assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(1.0, np.abs(ref["grad"]))) <= 1e-10
    import scipy.sparse as sp
    J = sp.coo_matrix((out["jac"][0], (jr, jc)), shape=(p.m, p.n)).tocsr()
    Jr = sp.coo_matrix((ref["jac"][2], (ref["jac"][0], ref["jac"][1])), shape=(p.m, p.n)).tocsr()
    assert abs(J - Jr).max() <= 1e-10 * max(1.0, abs(Jr).max())
    H = sp.coo_matrix((out["hess"][0], (hr, hc)), shape=(p.n, p.n)).tocsr()
    Hr = sp.coo_matrix((ref["hess"][2], (ref["hess"][0], ref["hess"][1])), shape=(p.n, p.n)).tocsr()
    assert abs(H - Hr).max() <= 1e-10 * max(1.0, abs(Hr).max())
    # the varying mass is felt: the constant-parameter problem gives other constraint values
    plain = CollocationNLP(p0)
    o0 = plain.eval_all(x, lam, 0.9)
    assert np.max(np.abs(o0["g"][0] - out["g"][0])) > 1e-3
    # ... and the same vector at every node is the constant-parameter problem, bit for bit
    pc = LapProblem(p0.model, p0.form, p0.closed, p0.s, p0.track_length, p0.track_nodes, p0.wl, p0.wr, p0.params, p0.dissipation, p0.sigma,
                    p0.elevation, node_params=np.tile(p0.params, (p0.n_points, 1)))
    same = CollocationNLP(pc)
    o1 = same.eval_all(x, lam, 0.9)
    for k in ("f", "g", "grad", "jac", "hess"):
        assert np.array_equal(o0[k], o1[k]), k
    nlp.close(); plain.close(); same.close()


def test_lap_with_a_mass_that_varies_along_the_lap():
    """A lap solved with the mass declared per node (fuel burnt along the lap: 700 -> 640 kg): it converges, the CPU oracle finds
    the KKT conditions satisfied at the returned point, and the lap time lies between the laps of the two constant masses."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.ipm import LapSolver, lap_time
    from fastest_lap_b200.problem import LapProblem
    from fastest_lap_b200.vehicle import Vehicle, MODEL_F1
    from tests.helpers import oracle_kkt
    p0 = workloads.f1_catalunya(500)
    times = {}
    for name in ("heavy", "light", "burning"):
        veh = Vehicle(MODEL_F1, p0.params.copy())
        if name == "burning":
            veh.declare_variable_parameter("vehicle/chassis/mass", [700.0, 640.0], [0, 1], [0.0, p0.track_length])
        else:
            veh.set_parameter("vehicle/chassis/mass", 700.0 if name == "heavy" else 640.0)
        node_params = veh.params_at(p0.s) if veh.has_variable_parameters() else None
        p = LapProblem(p0.model, p0.form, p0.closed, p0.s, p0.track_length, p0.track_nodes, p0.wl, p0.wr, veh.params, p0.dissipation, p0.sigma,
                       p0.elevation, node_params=node_params)
        nlp = CollocationNLP(p)
        solver = LapSolver(nlp)
        out = solver.solve(workloads.f1_start_point(p, 50.0))
        assert out["status"][0] == 1, (name, out["status"])
        times[name] = lap_time(p, out["x"][0], out["obj"][0])
        if name == "burning":
            stat, veq, vb, comp = oracle_kkt(p, out)
            assert stat <= 1e-8 and veq <= 1e-9, (stat, veq)
        solver.close(); nlp.close()
    print("lap times: 700 kg %.4f s, 700 -> 640 kg along the lap %.4f s, 640 kg %.4f s" % (times["heavy"], times["burning"], times["light"]))
    assert times["light"] < times["burning"] < times["heavy"]


@pytest.mark.parametrize("case", ["f1_catalunya_discrete", "f1_catalunya_3d", "kart_vendrell", "f1_catalunya_chicane_derivative"])
def test_wind(case):
    """vehicle/chassis/aerodynamics/wind_velocity/{northward, eastward} (chassis.h:276-277; chassis.hpp:264-288): callbacks with a
    wind of (12, -7) m/s equal the oracle's (pinned on the reference's closed-form test in tests/test_oracle_golden.py) on flat
    and 3-D tracks, both models and forms; and they differ from the calm ones."""
    from fastest_lap_b200.nlp import CollocationNLP
    import copy, scipy.sparse as sp
    p, ex = load_golden(case)
    pw = copy.copy(p)
    pw.wind = (12.0, -7.0)
    rng = np.random.default_rng(11)
    x = ex["x"] * (1 + 1e-3 * rng.standard_normal(p.n))
    lam = rng.standard_normal(p.m)
    nlp, calm = CollocationNLP(pw), CollocationNLP(p)
    out, out0 = nlp.eval_all(x, lam, 1.0), calm.eval_all(x, lam, 1.0)
    ref = oracle_nlp(pw).eval(x, lam, 1.0)
    jr, jc, hr, hc = nlp.structure()
    assert abs(out["f"][0] - ref["f"]) <= 1e-10 * abs(ref["f"])
    assert np.max(np.abs(out["g"][0] - ref["g"]) / np.maximum(1.0, np.abs(ref["g"]))) <= 1e-10
    assert np.max(np.abs(out["grad"][0] - ref["grad"]) / np.maximum(1.0, np.abs(ref["grad"]))) <= 1e-10
    J = sp.coo_matrix((out["jac"][0], (jr, jc)), shape=(p.m, p.n)).tocsr()
    Jr = sp.coo_matrix((ref["jac"][2], (ref["jac"][0], ref["jac"][1])), shape=(p.m, p.n)).tocsr()
    assert abs(J - Jr).max() <= 1e-10 * max(1.0, abs(Jr).max())
    H = sp.coo_matrix((out["hess"][0], (hr, hc)), shape=(p.n, p.n)).tocsr()
    Hr = sp.coo_matrix((ref["hess"][2], (ref["hess"][0], ref["hess"][1])), shape=(p.n, p.n)).tocsr()
    assert abs(H - Hr).max() <= 1e-10 * max(1.0, abs(Hr).max())
    assert np.max(np.abs(out["g"][0] - out0["g"][0])) > 1e-4
    nlp.close(); calm.close()


def test_callbacks_in_ipopts_call_order():
    """The five callbacks driven straight through the C ABI the way Ipopt drives a TNLP: first the two structure calls with
    values == NULL (triplets of the Jacobian and of the Hessian's lower triangle), then, per iterate, eval_f with new_x = true
    followed by eval_grad_f / eval_g / eval_jac_g / eval_h with new_x = false on the same x; then a new iterate announced
    by eval_g (Ipopt may start an iterate with any callback), and a Hessian call with new multipliers on an old x."""
    import ctypes as C
    from fastest_lap_b200.nlp import CollocationNLP, load_library, _ptr
    p, ex = load_golden("f1_catalunya_chicane")
    nlp = CollocationNLP(p)
    lib, h, n, m = load_library(), nlp._h, nlp.n, nlp.m
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    jr, jc = np.empty(nlp.nnz_jac, np.int32), np.empty(nlp.nnz_jac, np.int32)
    hr, hc = np.empty(nlp.nnz_hess, np.int32), np.empty(nlp.nnz_hess, np.int32)
    assert lib.fl_eval_jac_g(n, None, False, m, nlp.nnz_jac, ip(jr), ip(jc), None, h)
    assert lib.fl_eval_h(n, None, False, 1.0, m, None, False, nlp.nnz_hess, ip(hr), ip(hc), None, h)
    sj, sc, sh, shc = nlp.structure()
    assert np.array_equal(jr, sj) and np.array_equal(jc, sc) and np.array_equal(hr, sh) and np.array_equal(hc, shc)
    assert np.all(hr >= hc) and jr.max() == m - 1 and jc.max() == n - 1
    rng = np.random.default_rng(8)
    f, g, gr, jv, hv = C.c_double(), np.empty(m), np.empty(n), np.empty(nlp.nnz_jac), np.empty(nlp.nnz_hess)
    for it in range(3):
        x = np.ascontiguousarray(ex["x"] * (1 + 1e-3 * rng.standard_normal(n)))
        lam = np.ascontiguousarray(rng.standard_normal(m))
        ref = nlp.eval_all(x, lam, 0.5 + it)
        if it < 2:
            assert lib.fl_eval_f(n, _ptr(x), True, C.byref(f), h)
            assert lib.fl_eval_grad_f(n, _ptr(x), False, _ptr(gr), h)
            assert lib.fl_eval_g(n, _ptr(x), False, m, _ptr(g), h)
        else:
            assert lib.fl_eval_g(n, _ptr(x), True, m, _ptr(g), h)
            assert lib.fl_eval_f(n, _ptr(x), False, C.byref(f), h)
            assert lib.fl_eval_grad_f(n, _ptr(x), False, _ptr(gr), h)
        assert lib.fl_eval_jac_g(n, _ptr(x), False, m, nlp.nnz_jac, None, None, _ptr(jv), h)
        assert lib.fl_eval_h(n, _ptr(x), False, 0.5 + it, m, _ptr(lam), True, nlp.nnz_hess, None, None, _ptr(hv), h)
        assert f.value == ref["f"][0] and np.array_equal(g, ref["g"][0])
        np.testing.assert_allclose(gr, ref["grad"][0], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(jv, ref["jac"][0], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(hv, ref["hess"][0], rtol=1e-9, atol=1e-10)
        # new multipliers on the old x
        lam2 = np.ascontiguousarray(rng.standard_normal(m))
        assert lib.fl_eval_h(n, _ptr(x), False, 1.0, m, _ptr(lam2), True, nlp.nnz_hess, None, None, _ptr(hv), h)
        np.testing.assert_allclose(hv, nlp.eval_all(x, lam2, 1.0)["hess"][0], rtol=1e-9, atol=1e-10)
    # a wrong size is refused, not read past
    assert not lib.fl_eval_g(n, _ptr(x), False, m + 1, _ptr(g), h)
    nlp.close()
