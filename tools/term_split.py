"""Experiment: Hessian of the node Lagrangian split into additive terms (one per tyre + rest); registers of each term's code."""
import sys, os, subprocess, time
sys.path.insert(0, '/root/repo')
from fastest_lap_b200.codegen import models, nlpnode, expr
from fastest_lap_b200.codegen.expr import E, CONST, SYM
from fastest_lap_b200.codegen.emit import Emitter, ParamTable
from fastest_lap_b200.codegen import generate as G

rec = []
orig = models.f1_tire
T_syms = []
def patched(P, pre, kh, lam, N):
    kappa, Fx, Fy = orig(P, pre, kh, lam, N)
    g = kh.g
    i = len(rec)
    tx, ty = E(g, g.sym("T%dx" % i, "T")), E(g, g.sym("T%dy" % i, "T"))
    rec.append(dict(args=(kh, lam, N), real=(Fx, Fy), sym=(tx, ty)))
    return kappa, tx, ty
models.f1_tire = patched
t0 = time.time()
nlpnode.NodeProgram._differentiate = lambda self, *a, **k: None       # skip the flat differentiation
p = nlpnode.NodeProgram(nlpnode.Variant("f1", "direct"))
models.f1_tire = orig
g = p.g
print("built phi with tyre-force symbols in %.1f s; graph nodes %d" % (time.time() - t0, len(g.kind)))

def substitute(root, mapping):
    memo = {}
    for i in g.topo([root]):
        k = g.kind[i]
        if i in mapping: memo[i] = mapping[i]
        elif k in (CONST, SYM): memo[i] = i
        elif k == expr.ADD: memo[i] = g.add(memo[g.a[i]], memo[g.b[i]])
        elif k == expr.SUB: memo[i] = g.sub(memo[g.a[i]], memo[g.b[i]])
        elif k == expr.MUL: memo[i] = g.mul(memo[g.a[i]], memo[g.b[i]])
        else: memo[i] = g.unary(k, memo[g.a[i]])
    return memo[root]

phi = p.phi.i
zero = g.const(0.0)
Tall = [s.i for r in rec for s in r["sym"]]
Gt = {}
for t in Tall:
    d = g.forward([phi], t)[0]
    # affine in T: the coefficient must not depend on any T symbol
    dep = g.depends([d], set(Tall))
    assert d not in dep, "phi is not affine in the tyre forces"
    Gt[t] = d
phi_rest = substitute(phi, {t: zero for t in Tall})
terms = []
for i, r in enumerate(rec):
    tx, ty = r["sym"][0].i, r["sym"][1].i
    terms.append(("tyre%d" % i, g.add(g.mul(Gt[tx], r["real"][0].i), g.mul(Gt[ty], r["real"][1].i))))
terms.append(("rest", phi_rest))
xs = [x.i for x in p.x]
pt = ParamTable(g, p.param_names)
for name, root in terms:
    t0 = time.time()
    grad = g.reverse(root, xs)
    hess = {}
    for c in range(p.nv):
        col = g.forward(grad, xs[c])
        for r_ in range(c, p.nv):
            if col[r_] != zero: hess[(r_, c)] = col[r_]
    outs = list(hess.values())
    heavy = G.checkpoints(g, pt, outs)
    cone = G._cone(g, outs, stop=set(heavy))
    vars_used = sorted({c for (r_, c) in hess} | {r_ for (r_, c) in hess})
    print("%-6s hessian entries %3d over variables %s; ops (checkpoints loaded) %5d, checkpoints %3d  [%.1f s]" % (name, len(hess), vars_used, len(cone), len(heavy), time.time() - t0))
    r = dict(name=name, hess=hess, heavy=heavy)
    terms[[t[0] for t in terms].index(name)] = (name, root, r)

# ---- emit one kernel per term and compile
inputs = {}
pre = []
for k in range(p.nv):
    inputs["x%d" % k] = "x%d" % k; pre.append("const double x%d = io.x(%d);" % (k, k))
tn = p.track_names + p.geom_names
for k, n in enumerate(tn):
    inputs["t_" + n] = "t_" + n; pre.append("const double t_%s = io.t(%d);" % (n, k))
pre.append("const double obj = io.obj();"); inputs["obj"] = "obj"
for r_ in range(p.ncpe):
    inputs["lin%d" % r_] = "lin%d" % r_; pre.append("const double lin%d = io.lin(%d);" % (r_, r_))
    if p.lam_out[r_] is not None:
        inputs["lout%d" % r_] = "lout%d" % r_; pre.append("const double lout%d = io.lout(%d);" % (r_, r_))
for c in range(len(p.uprev)):
    for nm in ("uprev", "unext"):
        inputs["%s%d" % (nm, c)] = "%s%d" % (nm, c); pre.append("const double %s%d = io.%s(%d);" % (nm, c, nm, c))
src = ['#include <cuda_runtime.h>', '#include "/root/repo/fastest_lap_b200/csrc/fl_math.cuh"',
       'struct IO { const double* X; const double* T; const double* D; const double* C; const double* L; double* H; int n;',
       '  __device__ double x(int k) const { return X[k * n]; } __device__ double t(int k) const { return T[k * n]; } __device__ double d(int k) const { return __ldg(D + k); }',
       '  __device__ double c(int k) const { return C[k * n]; } __device__ double lin(int k) const { return L[k * n]; } __device__ double lout(int k) const { return L[(32 + k) * n]; }',
       '  __device__ double obj() const { return L[64 * n]; } __device__ double uprev(int k) const { return X[(20 + k) * n]; } __device__ double unext(int k) const { return X[(24 + k) * n]; }',
       '  __device__ void h(int k, double v) const { H[k * n] = v; } };']
for name, root, r in terms:
    cut = {nd: "io.c(%d)" % k for k, nd in enumerate(r["heavy"])}
    em = Emitter(g, pt, inputs, dfmt="io.d(%d)", cut=cut, order=os.environ.get("ORDER", "dfs"))
    # only the inputs the term uses
    outs = [(e, "io.h(%d, %%s);" % k) for k, e in enumerate(r["hess"].values())]
    body = em.body(outs, indent="    ")
    used = [l for l in pre if (" " + l.split()[2] + " ") in (" " + body.replace("(", " ").replace(")", " ").replace(",", " ").replace(";", " ").replace("-", " ") + " ")]
    src.append("__global__ void __launch_bounds__(128, %s) k_%s(IO io_)\n{\n    IO io = io_; const int i = blockIdx.x * blockDim.x + threadIdx.x; io.X += i; io.T += i; io.C += i; io.L += i; io.H += i;\n%s\n%s\n}" % (
        os.environ.get("MINB", "4"), name, "\n".join("    " + l for l in used), body))
open("/tmp/scratch/term_split.cu", "w").write("\n".join(src))
r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v", "-c", "/tmp/scratch/term_split.cu", "-o", "/tmp/scratch/term_split.o"],
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
import re
for m in re.finditer(r"Compiling entry function '_Z\d+k_(\w+?)2IO'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", r.stdout):
    print("kernel %-6s: %s registers, stack %s B, spill stores %s B" % (m.group(1), m.group(4), m.group(2), m.group(3)))
if "error" in r.stdout: print(r.stdout[-3000:])
