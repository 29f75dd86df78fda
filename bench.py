#!/usr/bin/env python
"""Benchmark of the hot path: collocation node-evals/s (g + Jacobian + Hessian) on the F1 3DOF Catalunya lap refined to
8000 nodes (BASELINE.json configs[2], the configuration the north_star target is quoted on).

A step = one full callback round (f, g, grad f, Jacobian values, Lagrangian-Hessian values) over every node of the lap.
  value   node-evals/s with x and lambda resident in HBM (CUDA events on the library's stream, L2 flushed between steps)
  e2e     the same through the C-ABI call with HOST buffers: x, lambda host->device and f, g, grad, jac, hess device->host
          inside the timed region (pinned memory)
  --impl reference   the CPU oracle port (oracle/, OpenMP over nodes, all host cores) on the same workload
Under torchrun every rank evaluates its own lap on its own GPU (weak scaling, no data-path collective; one NCCL
all_gather of the objective values at the end); time = max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collocation node-evals/s (g+Jac+Hess), F1 3DOF Catalunya 8000 nodes"
UNIT = "node-evals/s"
N_NODES = 8000


def workload(n_nodes=N_NODES):
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.problem import LapProblem
    p = workloads.f1_catalunya(n_nodes)
    gp, ex = LapProblem.load_npz(os.path.join(ROOT, "tests", "golden", "f1_catalunya_discrete.npz"))
    x = workloads.resample_point(p, gp.s, ex["x"], p.nv)
    lam = np.random.default_rng(1).standard_normal(p.m)
    return p, x, lam


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def host_threads():
    """Host cores this process may use -- not OMP_NUM_THREADS, which torchrun sets to 1 for every rank."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def headline_config(p, nlp_sizes=None):
    """The `config` object of the JSON line: the same keys and values in the b200 arm and in the reference arm."""
    return {"workload": "limebeer-2014-f1 3DOF, Catalunya, closed, direct form, %d trapezoidal nodes, one full callback round per step" % p.n_points,
            "nodes": p.n_points, "n": p.n, "m": p.m,
            "point": "golden 500-node solution interpolated to the mesh, lambda ~ N(0,1) seed 1, obj_factor 1"}


def cpu_baseline(p, x, lam, budget_s=12.0, threads=0):
    """Times the CPU oracle port on a bounded sample: whole callback rounds of the same lap, about `budget_s` seconds."""
    from tests.helpers import oracle_nlp
    o = oracle_nlp(p)
    cores = host_threads() if threads == 0 else threads
    t1 = o.time_eval(x, lam, 1.0, reps=1, threads=cores)
    reps = max(1, min(200, int(budget_s / max(t1, 1e-6))))
    t = o.time_eval(x, lam, 1.0, reps=reps, threads=cores)
    return {"value": p.n_points / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d full callback rounds (f, g, grad, Jacobian, Hessian) of the %d-node lap, C++ second-order jets (-O3), OpenMP over element chunks" % (reps, p.n_points)}, t


def run_reference(args):
    """Reference arm: the CPU statement of the path (oracle port; the reference's own stack -- lion, CppAD, Ipopt, MUMPS -- cannot
    be built offline) on all host cores, same metric / config; W warm-up and K timed steps as asked.  A step is a full callback
    round of the 8000-node lap while K rounds fit a couple of minutes, else a round of a shorter closed lap (the per-node work
    does not depend on the mesh; the sample is named)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p, x, lam = workload(args.nodes)
    from tests.helpers import oracle_nlp
    cores = host_threads()
    o = oracle_nlp(p)
    t1 = o.time_eval(x, lam, 1.0, reps=1, threads=cores)
    ps, xs, ls, os_ = p, x, lam, o
    if t1 * (args.steps + args.warmup) > 150.0:
        n_s = max(64, int(p.n_points * 150.0 / (t1 * (args.steps + args.warmup))))
        ps, xs, ls = workload(n_s)
        os_ = oracle_nlp(ps)
    for _ in range(args.warmup):
        os_.time_eval(xs, ls, 1.0, reps=1, threads=cores)
    t = os_.time_eval(xs, ls, 1.0, reps=args.steps, threads=cores)
    v = ps.n_points / t
    sample = "%d full callback rounds of %s on the CPU oracle port, %d threads (the reference itself cannot be built offline)" % (
        args.steps, "the %d-node lap" % p.n_points if ps is p else "a %d-node closed lap of the same track (bounded sample of the %d-node workload)" % (ps.n_points, p.n_points), cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3 * (p.n_points / ps.n_points), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": headline_config(p),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


N_LAPS = 4096


def batched_leg(args, torch, dist, world, rank, local, peak):
    """BASELINE.json configs[4]: 4096 independent F1 laps of 500 nodes with perturbed vehicle parameters, sharded
    contiguously over the ranks (no data-path collective; one gather of the objective values at the end).  A step = one
    full callback round of every lap.  The batch's buffers (GBs) exceed the L2 many times over: no flush needed."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP, EVAL_ALL, EVAL_JAC, EVAL_HESS, EVAL_FG, EVAL_NO_VALUES, pinned_empty
    from fastest_lap_b200.problem import LapProblem
    from fastest_lap_b200.sharding import shard_range, gather_results
    from fastest_lap_b200.vehicle import MODEL_F1
    p, ex = LapProblem.load_npz(os.path.join(ROOT, "tests", "golden", "f1_catalunya_discrete.npz"))
    lo, hi = shard_range(args.laps, world, rank)
    B = hi - lo
    params = workloads.sweep_parameters(MODEL_F1, args.laps)[lo:hi]
    nlp = CollocationNLP(p, vehicle_params=params, device=local)
    rng = np.random.default_rng(1000 + lo)
    hx, _ = pinned_empty(B * nlp.n); hl, _ = pinned_empty(B * nlp.m)
    X = hx.reshape(B, nlp.n); Lm = hl.reshape(B, nlp.m)
    X[:] = ex["x"][None, :] * (1.0 + 1.0e-3 * rng.standard_normal((B, 1)))
    Lm[:] = np.random.default_rng(1).standard_normal(nlp.m)[None, :]
    out = {}
    for k, cnt in (("f", B), ("g", B * nlp.m), ("grad", B * nlp.n), ("jac", B * nlp.nnz_jac), ("hess", B * nlp.nnz_hess)):
        out[k], _ = pinned_empty(cnt)
    nlp.eval_all(hx, hl, 1.0, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        nlp.sync()

    K = max(1, min(args.steps, 20))
    for _ in range(3):
        nlp.time_device(EVAL_ALL, 1.0, 1)
    barrier()
    l0 = nlp.kernel_launches
    dev_ms = sum(nlp.time_device(EVAL_ALL, 1.0, 1) for _ in range(K))
    barrier()
    launches = nlp.kernel_launches - l0
    val_ms = sum(nlp.time_device(0, 1.0, 1) for _ in range(K))
    ker_ms = sum(nlp.time_device(EVAL_JAC | EVAL_HESS | EVAL_NO_VALUES, 1.0, 1) for _ in range(K))     # the derivative kernel alone, its own duration
    Ke = 2
    nlp.eval_all(hx, hl, 1.0, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        nlp.eval_all(hx, hl, 1.0, out=out)
    barrier()
    e2e_s = (time.perf_counter() - t0) / Ke
    t = torch.tensor([dev_ms / K, e2e_s, ker_ms / K, val_ms / K], dtype=torch.float64, device="cuda" if dist is not None else "cpu")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f_all = gather_results(out["f"].reshape(B, 1).copy(), args.laps)          # the final gather of per-lap results
    dev_ms, e2e_s, ker_ms, val_ms = [float(v) for v in t.tolist()]
    nodes = p.n_points * args.laps
    per_node = (p.nv + p.ncpe + 5) * 8 + p.nv * 8 + (nlp.nnz_jac + nlp.nnz_hess) * 8 / p.n_points
    # per-rank kernel: the slowest rank's launch over the largest shard
    shard_nodes = p.n_points * (-(-args.laps // world))
    achieved = per_node * shard_nodes / (ker_ms * 1e-3) / 1e9
    res = {"workload": "%d independent limebeer-2014-f1 laps, Catalunya, 500 nodes, closed, direct form; mass, cd, cl, power drawn per lap (rng 2); "
                       "contiguous shards over %d GPU(s), no data-path collective, one final all_gather of the %d objective values" % (args.laps, world, args.laps),
           "laps": args.laps, "laps_per_gpu": B, "scaling": "strong", "value": nodes / (dev_ms * 1e-3), "unit": UNIT, "steps": K, "ms_per_step": dev_ms,
           "rounds_per_s": args.laps / (dev_ms * 1e-3), "gpu_launches": int(launches),
           "e2e": {"value": nodes / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "steps": Ke, "h2d_bytes_per_step": int(B * (nlp.n + nlp.m) * 8),
                   "d2h_bytes_per_step": int(B * (1 + nlp.m + nlp.n + nlp.nnz_jac + nlp.nnz_hess) * 8)},
           "kernels_ms": {"node_values": val_ms, "node_derivs+assemble": ker_ms},
           "roofline": {"bound": "hbm", "kernel": "k_node_derivs<%s, all> (per-lap parameters from global memory)" % nlp.variant, "achieved": achieved,
                        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "bytes_per_node_eval": per_node,
                        "fp64_ops_per_node_eval": nlp.ops_per_node["all"]},
           "objective_checksum": float(np.sum(f_all))}
    nlp.close()
    return res


N_SOLVE = 4096


def solve_leg(args, torch, dist, world, rank, local, weak=False):
    """Batched laps/s: `--solve-laps` independent F1 laps (Catalunya, 500 nodes, perturbed vehicles, configs[4] scaled to a
    default bench run) solved to Ipopt's tolerances by the device interior-point solver from the reference's starting point
    (steady state at 50 km/h replicated).  Sharded contiguously over the ranks; one final all_gather of (lap time, status,
    iterations).  Timed end to end through the public API: host start point in, host solution out."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.ipm import LapSolver, lap_time
    from fastest_lap_b200.sharding import shard_range, gather_results
    from fastest_lap_b200.vehicle import MODEL_F1
    p = workloads.f1_catalunya(500)
    n_total = args.solve_laps * (world if weak else 1)
    lo, hi = shard_range(n_total, world, rank)
    B = hi - lo
    params = workloads.sweep_parameters(MODEL_F1, n_total)
    params[0] = p.params                                     # lap 0 drives the reference vehicle (known answer)
    from fastest_lap_b200.ipm import default_options, solve_sweep
    opts = default_options()
    opts.max_iter = args.solve_max_iter       # first pass; laps that need more go to the retry passes (reported below); Ipopt's default is 3000
    # untimed warm-up: three iterations of a 64-lap batch (the smallest the batch schedule serves) load the solver's kernels
    warm = default_options()
    warm.max_iter = 3
    solve_sweep(p, params[:64], options=warm, device=local, start_speeds_kmh=(50.0,))
    # ... and of a 4-lap batch: the kernels of the small-batch schedule, which the retry pass of the timed sweep uses (their first use
    # inside the timed region showed as 0.59 ... 0.96 s of "retries" for the same 0.58 s of device time)
    solve_sweep(p, params[:4], options=warm, device=local, start_speeds_kmh=(50.0,))
    # ... and the device memory of the sweep's own handles: created once and released into the library's cache of freed blocks, from which
    # the timed sweep gets them back (the first cudaMalloc of their 45 GB took 0.15 ... 1.8 s from box to box: a service that runs sweep
    # after sweep pays it once)
    from fastest_lap_b200.nlp import CollocationNLP
    from fastest_lap_b200.ipm import LapSolver, sweep_bounds
    nlp_w = CollocationNLP(p, vehicle_params=params[lo:hi], device=local)
    LapSolver(nlp_w, options=opts, bounds=sweep_bounds(p, params[lo:hi])).close()
    nlp_w.close()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = solve_sweep(p, params[lo:hi], options=opts, device=local)      # start: steady state at 50 km/h (GPU steady-state solve), then retries
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ph = out["phases_s"]
    t = torch.tensor([wall, out["device_ms"], ph["allocate"], ph["start_point"], ph["first_pass"], ph["retries"]], dtype=torch.float64, device="cuda" if dist is not None else "cpu")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_ms = [float(v) for v in t.tolist()[:2]]
    phases = dict(zip(("allocate", "start_point", "first_pass", "retries"), [round(float(v), 3) for v in t.tolist()[2:]]))
    times = np.array([lap_time(p, out["x"][b], out["obj"][b]) for b in range(B)])
    local_res = np.stack([times, out["status"].astype(np.float64), out["iters"].astype(np.float64)], axis=1)
    res = gather_results(local_res, n_total)          # the final gather of per-lap results
    ok = res[:, 1] >= 1
    st = out["stats"]
    golden = 77.791438
    r = {"workload": "%d independent limebeer-2014-f1 laps, Catalunya, 500 nodes, closed, direct form, vehicles drawn per lap (rng 2), start = steady state at 50 km/h; "
                     "tol 1e-10 (the reference's Ipopt options), at most %d iterations per lap; contiguous shards over %d GPU(s)" % (n_total, args.solve_max_iter, world),
         "laps": n_total, "laps_per_gpu": B, "scaling": "weak" if weak else "strong", "value": n_total / wall, "unit": "laps/s", "seconds": wall, "device_ms": dev_ms,
         "phases_s": phases,
         "converged": int(ok.sum()), "not_converged": int((~ok).sum()), "iterations_median": float(np.median(res[:, 2])), "iterations_max": float(res[:, 2].max()),
         "lap0_time_s": float(res[0, 0]), "lap0_reference_s": golden, "lap0_rel_err": abs(float(res[0, 0]) - golden) / golden,
         "lap_time_range_s": [float(res[ok, 0].min()), float(res[ok, 0].max())] if ok.any() else None,
         "rank0_launches": st, "rank0_passes": out["passes"], "reference_cpu": "Ipopt+MUMPS in the reference: 82.7 s for a 697-node lap (docs quickstart.rst:134), about 1 lap/minute/core"}
    return r



def gg_leg(args, torch, dist, world, rank, local):
    """BASELINE.json configs[1]: roberto-lot-2016-kart G-G diagram sweep, 32 speeds x 64 lateral accelerations, maximum and
    minimum longitudinal acceleration at each (plus the 0 g and maximum-lateral NLPs per speed), solved as batches of
    one-node NLPs by the device interior-point solver.  Speeds are sharded over the ranks."""
    from fastest_lap_b200 import gg, workloads
    from fastest_lap_b200.sharding import shard_range, gather_results
    params = workloads._vehicles()["roberto_lot_kart_2016"]
    speeds = np.linspace(50.0, 120.0, args.gg_speeds) / 3.6
    lo, hi = shard_range(args.gg_speeds, world, rank)
    # untimed warm-up: the same sweep once.  The timed sweep then finds the steady-state kernels loaded and the device blocks of its
    # dozen handles in the library's cache (cudaMalloc / cudaFree of a first sweep varied from 30 ms to 1.2 s on fresh boxes against
    # 0.11 s of solves, profiles/README.md)
    # (the library's cache of freed device blocks is full of the lap sweeps' blocks at this point and would refuse the sweep's own)
    from fastest_lap_b200.nlp import load_library
    load_library().fl_release_cached_memory()
    gg.gg_diagram(params, speeds[lo:hi], n_lat=args.gg_lat, device=local)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    # the sweep is 0.1 s: three of them are timed and the mean is reported (one sweep varied by +-10 % from run to run)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        r = gg.gg_diagram(params, speeds[lo:hi], n_lat=args.gg_lat, device=local)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    t = torch.tensor([wall], dtype=torch.float64, device="cuda" if dist is not None else "cpu")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    local_res = np.concatenate([r["ax_max"], r["ax_min"], r["ax_max_solved"].astype(float), r["ax_min_solved"].astype(float), r["ay_max"][:, None]], axis=1)
    res = gather_results(local_res, args.gg_speeds)
    L = args.gg_lat
    n_nlp = args.gg_speeds * (2 * L + 2)
    return {"workload": "roberto-lot-kart-2016 G-G diagram: %d speeds (50..120 km/h) x %d lateral accelerations x {max, min} longitudinal acceleration "
                        "+ 0 g and maximum-lateral NLPs per speed; 8 variables, 6 equations, 6 tyre-slip rows each; speeds sharded over %d GPU(s)" % (args.gg_speeds, L, world),
            "nlps": n_nlp, "value": n_nlp / wall, "unit": "NLPs/s", "seconds": wall, "sweeps_timed": reps, "solved_fraction_max": float(res[:, 2 * L:3 * L].mean()),
            "solved_fraction_min": float(res[:, 3 * L:4 * L].mean()), "ay_max_range": [float(res[:, -1].min()), float(res[:, -1].max())],
            "reference_cpu": "the reference solves these NLPs one after the other with Ipopt (steady_state.hpp:594-880), ~130 solves per speed"}



def kart_leg(args, torch, local):
    """Callback rounds of BASELINE configs[3]: roberto-lot-kart-2016 (6DOF) on Vendrell, 2000 nodes, derivative form, at the
    golden solution interpolated to the mesh; same ring-of-instances hygiene as the headline leg (rank 0 only)."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.nlp import CollocationNLP, EVAL_ALL, EVAL_JAC, EVAL_HESS, EVAL_NO_VALUES, time_device_ring
    p = workloads.kart_vendrell(2000)
    g = np.load(os.path.join(ROOT, "tests", "golden", "kart_vendrell.npz"))
    x = workloads.resample_point(p, g["s"], g["x"], p.nv)
    lam = np.random.default_rng(1).standard_normal(p.m)
    nlp = CollocationNLP(p, device=local)
    per_instance = 8 * (2 * nlp.n + 2 * nlp.m + nlp.nnz_jac + nlp.nnz_hess + (nlp.n_values + nlp.n_checkpoints) * p.n_points)
    R = max(2, int(np.ceil(1.5 * torch.cuda.get_device_properties(local).L2_cache_size / per_instance)) + 1)
    ring = [nlp] + [CollocationNLP(p, device=local) for _ in range(R - 1)]
    for h in ring:
        h.eval_all(x, lam, 1.0, want=("f",))
    K = max(args.steps, R)
    timed = lambda what, count: time_device_ring(ring, what, count)
    timed(EVAL_ALL, R)
    torch.cuda.synchronize()
    dev_ms = timed(EVAL_ALL, K) / K
    ker_ms = timed(EVAL_JAC | EVAL_HESS | EVAL_NO_VALUES, K) / K          # the derivative kernel alone
    res = {"workload": "roberto-lot-kart-2016 6DOF, Vendrell, closed, derivative form, 2000 trapezoidal nodes, one full callback round per step; ring of %d instances" % R,
           "variant": nlp.variant, "value": p.n_points / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms, "steps": K,
           "kernels_ms": {"node_derivs": ker_ms}, "fp64_ops_per_node_eval": nlp.ops_per_node["all"], "n": nlp.n, "m": nlp.m,
           "nnz_jac": nlp.nnz_jac, "nnz_hess": nlp.nnz_hess}
    for h in ring:
        h.close()
    return res


def single_lap_leg(args, torch, local):
    """One lap at a time (replicas only): BASELINE configs[0] (F1, Catalunya, 500 nodes: the reference's README example,
    "approximately 1 minute") and configs[2] (8000 nodes), solved from the steady-state start on this rank's GPU with the
    parallel-in-the-stages factorisation (the 8000-node mesh by mesh sequencing from the lap on 500 of its nodes, ipm.solve_lap);
    wall time through the public API."""
    from fastest_lap_b200 import workloads
    from fastest_lap_b200.ipm import solve_lap, lap_time
    out = {}
    for name, n in (("f1_catalunya_500", 500), ("f1_catalunya_8000", args.nodes)):
        p = workloads.f1_catalunya(n)
        node = workloads.f1_start_point(p, 50.0, device=local)[:p.nv]
        if n == 500:
            solve_lap(p, node, device=local)            # warm-up of the 500-node case (library initialisation)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r, levels = solve_lap(p, node, device=local)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"nodes": n, "seconds": dt, "status": int(r["status"][0]), "iterations": int(sum(l["iterations"] for l in levels)),
                     "lap_time_s": lap_time(p, r["x"][0], r["obj"][0]), "levels": levels}
    out["f1_catalunya_500"]["reference_lap_time_s"] = 77.791438
    out["reference_cpu"] = "README.md:26 of the reference: approximately 1 minute for the 500-point lap"
    return out


def pin_to_gpu_numa_node(torch, local):
    """Multi-rank runs: this process's threads -- and with them the pinned host buffers it allocates from here on (first touch) -- go
    to the CPUs next to the rank's GPU (/sys/bus/pci/devices/<bus id>/local_cpulist), so that N ranks moving 26 MB per step each do
    not cross the socket interconnect.  Returns what was done, for the JSON line; never fails the run."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu": bdf, "pinned": False, "why": "no local CPU in this process's affinity mask"}
        os.sched_setaffinity(0, cpus)
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        return {"gpu": bdf, "numa_node": int(node), "cpus": len(cpus), "pinned": True}
    except Exception as e:          # containers without sysfs, older torch
        return {"pinned": False, "why": "%s: %s" % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nodes", type=int, default=N_NODES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--laps", type=int, default=N_LAPS, help="laps of the batched callback leg (0 = skip)")
    ap.add_argument("--solve-laps", type=int, default=N_SOLVE, help="laps solved to convergence in the laps/s leg (0 = skip)")
    ap.add_argument("--solve-max-iter", type=int, default=250)
    ap.add_argument("--single-laps", type=int, default=1, help="1: also time single-lap solves (500 and --nodes nodes) on rank 0")
    ap.add_argument("--gg-speeds", type=int, default=32, help="speeds of the G-G sweep leg (0 = skip)")
    ap.add_argument("--gg-lat", type=int, default=64)
    ap.add_argument("--cpu-solver-arms", type=int, default=1, help="1: time the CPU arms of laps/s and NLPs/s (N = 1, rank 0)")
    ap.add_argument("--cpu-arm-procs", type=int, default=32, help="processes of the CPU solver arms (at most the host cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    numa = pin_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from fastest_lap_b200.nlp import CollocationNLP, EVAL_ALL, EVAL_JAC, EVAL_HESS, EVAL_FG, EVAL_NO_VALUES, pinned_empty, time_device_ring
    p, x, lam = workload(args.nodes)
    # Ring of independent instances of the same lap: consecutive timed steps touch different buffers, and one trip round
    # the ring moves more bytes than the L2 holds, so every step finds its data in HBM (no flush kernel in between).
    nlp = CollocationNLP(p, device=local)
    l2_bytes = torch.cuda.get_device_properties(local).L2_cache_size
    per_instance = 8 * (2 * nlp.n + 2 * nlp.m + nlp.nnz_jac + nlp.nnz_hess + (nlp.n_values + nlp.n_checkpoints) * p.n_points)
    R = max(2, int(np.ceil(1.5 * l2_bytes / per_instance)) + 1)
    ring = [nlp] + [CollocationNLP(p, device=local) for _ in range(R - 1)]
    B = nlp.batch
    W = max(args.warmup, 3)
    K = args.steps

    # pinned host buffers for the end-to-end leg
    hx, _ = pinned_empty(B * nlp.n); hl, _ = pinned_empty(B * nlp.m)
    hx[:] = x; hl[:] = lam
    out = {}
    for k, cnt in (("f", B), ("g", B * nlp.m), ("grad", B * nlp.n), ("jac", B * nlp.nnz_jac), ("hess", B * nlp.nnz_hess)):
        out[k], _ = pinned_empty(cnt)
    for h in ring:
        h.eval_all(hx, hl, 1.0, out=out)             # also leaves x, lambda resident on the device

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        for h in ring:
            h.sync()

    def timed(what, count):
        # `count` steps issued back to back on one stream, step k on instance k mod R, one pair of CUDA events around the
        # sequence (fl_time_device_ring): device time of exactly `count` steps, free of the host's launch latency
        return time_device_ring(ring, what, count)

    # ---- device-resident leg: blocks of exactly K steps, repeated until at least 50 ms of steps have been timed (a single
    # block of 20 steps is 0.7 ms: +-5 % run to run); the reported time per step is the mean over all blocks
    timed(EVAL_ALL, max(W, R))
    sampler = ClockSampler(local); sampler.start()
    barrier()
    launches0 = sum(h.kernel_launches for h in ring)
    dev_ms = timed(EVAL_ALL, K)
    barrier()
    launches = sum(h.kernel_launches for h in ring) - launches0
    repeats = 1
    blocks = [dev_ms]
    while sum(blocks) < 50.0 and repeats < 2000:
        blocks.append(timed(EVAL_ALL, K))
        repeats += 1
    dev_ms = sum(blocks) / repeats
    barrier()
    M = max(1, int(np.ceil(50.0 / max(dev_ms, 1e-3))))

    def timed_mean(what):
        timed(what, R)
        return sum(timed(what, K) for _ in range(M)) / M
    # dominant kernel (grad f + Jacobian + Hessian of every node; it needs the checkpoints of the values pass): (a) values +
    # derivatives minus values alone, as the step runs them (programmatic dependent launch lets the two overlap); (b) the
    # derivative kernel launched ALONE on the checkpoints left by the last values pass, its own duration between two events
    ker_ms = timed_mean(EVAL_JAC | EVAL_HESS)
    val_ms = timed_mean(0)
    ker_ms -= val_ms
    ker_alone_ms = timed_mean(EVAL_JAC | EVAL_HESS | EVAL_NO_VALUES)
    fg_ms = timed_mean(EVAL_FG)
    fp64_peak = nlp.measure_fp64_peak()            # DFMA/s, micro-kernel (no FP64 figure in MEASURED_PEAKS.json)
    # ---- end-to-end leg: host buffers in, host buffers out
    for k in range(3):
        ring[k % R].eval_all(hx, hl, 1.0, out=out)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        ring[k % R].eval_all(hx, hl, 1.0, out=out)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, e2e_s, ker_ms, fg_ms, ker_alone_ms], dtype=torch.float64, device="cuda" if dist is not None else "cpu")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fs = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(fs, torch.tensor([float(out["f"][0])], dtype=torch.float64, device="cuda"))     # the final gather of results
    dev_ms, e2e_s, ker_ms, fg_ms, ker_alone_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        nodes = p.n_points * B * world
        ms_per_step = dev_ms / K
        v = nlp.variant
        # algorithmic bytes of the dominant kernel per node-eval: read x, lambda, per-node constants; write grad, Jacobian, Hessian slots
        nv, ncpe = p.nv, p.ncpe
        per_node_read = (nv + ncpe + 5) * 8
        per_node_write = nv * 8 + (nlp.nnz_jac + nlp.nnz_hess) * 8 / p.n_points
        ker_bytes = (per_node_read + per_node_write) * p.n_points * B
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # the roofline figure uses the kernel's OWN duration (launched alone between two events), not the subtraction
        achieved = ker_bytes / (ker_alone_ms / K * 1e-3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_node_derivs<f1_direct, all> @ 8000 nodes"]
            if p.n_points == 8000 and B == 1:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]          # per launch, from the committed ncu --set full capture
        except Exception:
            pass
        line = {"metric": METRIC, "value": nodes / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": headline_config(p),
                "details": {"laps_per_gpu": B, "nnz_jac": nlp.nnz_jac, "nnz_hess": nlp.nnz_hess, "variant": v, "timed_blocks": repeats, "timed_ms": sum(blocks),
                            "l2": "ring of %d independent instances of the lap (%.0f MB touched per trip > %.0f MB L2), one per timed step in turn" % (R, R * per_instance / 1e6, l2_bytes / 1e6), "parallelism": "1 lap per GPU, replicas"},
                "e2e": {"value": nodes / (e2e_s / K), "unit": UNIT, "h2d_bytes_per_step": int(B * (nlp.n + nlp.m) * 8),
                        "d2h_bytes_per_step": int(B * (1 + nlp.m + nlp.n + nlp.nnz_jac + nlp.nnz_hess) * 8), "ms_per_step": e2e_s / K * 1e3,
                        "pcie_gbs": (B * (nlp.n + nlp.m) + B * (1 + nlp.m + nlp.n + nlp.nnz_jac + nlp.nnz_hess)) * 8 / (e2e_s / K) / 1e9,
                        "host_numa": numa if numa is not None else "single rank: not pinned"},
                "gpu_launches": int(launches),
                "kernels_ms": {"node_values": val_ms / K, "values+assemble": fg_ms / K, "node_derivs": ker_ms / K, "node_derivs_alone": ker_alone_ms / K},
                "roofline": {"bound": "hbm", "kernel": "k_node_derivs<%s, all>" % v, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "algorithmic_bytes_per_launch": ker_bytes, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch)",
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                             "bytes_per_node_eval": per_node_read + per_node_write, "fp64_ops_per_node_eval": nlp.ops_per_node["all"],
                             "fp64": {"peak_measured_tflops": 2.0 * fp64_peak / 1e12, "peak_unit": "TFLOP/s (2 x DFMA/s, micro-kernel on this GPU)",
                                      "achieved_gops": nlp.ops_per_node["all"] * p.n_points * B / (ker_alone_ms / K * 1e-3) / 1e9,
                                      "frac_of_instruction_peak": nlp.ops_per_node["all"] * p.n_points * B / (ker_alone_ms / K * 1e-3) / fp64_peak,
                                      "note": "operations of the generated code (add/mul/fma each one instruction) per second over the DFMA issue rate"}},
                "clocks": clocks}
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline(p, x, lam)
            line["cpu_baseline"] = cb
    for h in ring:
        h.close()
    ring = None
    if args.laps > 0:
        peak_ = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        b = batched_leg(args, torch, dist, world, rank, local, peak_)
        if rank == 0:
            line["batched"] = b
            line["gpu_launches"] += b["gpu_launches"]
    if args.solve_laps > 0:
        sl = solve_leg(args, torch, dist, world, rank, local)
        if world > 1:
            # the same sweep per GPU (weak scaling: world x laps in all), next to the fixed-size sweep above (strong scaling)
            wk = solve_leg(args, torch, dist, world, rank, local, weak=True)
            sl["weak"] = {k: wk[k] for k in ("laps", "laps_per_gpu", "value", "unit", "seconds", "converged", "phases_s")}
        if rank == 0:
            line["batched_solve"] = sl
            line["gpu_launches"] += int(sl["rank0_launches"]["solver_launches"]) + 3 * int(sl["rank0_launches"]["full_rounds"]) + 2 * int(sl["rank0_launches"]["fg_rounds"])
    if args.single_laps and rank == 0:
        line["kart_callbacks"] = kart_leg(args, torch, local)
        line["single_lap_solve"] = single_lap_leg(args, torch, local)
    if args.gg_speeds > 0:
        gl = gg_leg(args, torch, dist, world, rank, local)
        if rank == 0:
            line["gg_sweep"] = gl
    if rank == 0:
        # timed CPU arms of the two solver metrics (rank 0, N = 1 only): P processes x 1 core, bounded samples
        if not args.no_cpu_baseline and world == 1 and args.cpu_solver_arms:
            from oracle import cpu_arms
            P = max(1, min(host_threads(), args.cpu_arm_procs))
            if args.solve_laps > 0:
                line["batched_solve"]["cpu_baseline"] = cpu_arms.laps_per_s(ROOT, P)
            if args.gg_speeds > 0:
                line["gg_sweep"]["cpu_baseline"] = cpu_arms.nlps_per_s(ROOT, P)
        # compact summary as the LAST key, so that the tail of the line carries it at every N
        sc = {"n_gpus": world, "node_evals_per_s": line["value"]}
        if "batched" in line:
            sc["batch_node_evals_per_s"] = line["batched"]["value"]; sc["batch_roofline_frac"] = round(line["batched"]["roofline"]["frac"], 4)
        if "batched_solve" in line:
            bs = line["batched_solve"]
            sc.update(laps_per_s=bs["value"], laps=bs["laps"], laps_converged=bs["converged"], laps_seconds=bs["seconds"], laps_scaling="strong",
                      laps_phase_s_max_over_ranks=bs.get("phases_s"))
            if "weak" in bs:
                sc.update(laps_per_s_weak=bs["weak"]["value"], laps_weak=bs["weak"]["laps"])
            if "cpu_baseline" in bs:
                sc.update(cpu_laps_per_s=bs["cpu_baseline"]["value"], cpu_cores=bs["cpu_baseline"]["cores"])
        if "gg_sweep" in line:
            gs = line["gg_sweep"]
            sc.update(nlps_per_s=gs["value"], nlps=gs["nlps"], nlps_solved_fraction=min(gs["solved_fraction_max"], gs["solved_fraction_min"]))
            if "cpu_baseline" in gs:
                sc.update(cpu_nlps_per_s=gs["cpu_baseline"]["value"])
        sc["roofline_frac"] = round(line["roofline"]["frac"], 4)
        line["scale"] = sc
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
