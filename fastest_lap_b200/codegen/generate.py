"""Writes csrc/generated/gen_<variant>.cuh: the straight-line fp64 device code of one collocation node.

    python -m fastest_lap_b200.codegen.generate [variant ...]

Each header defines `struct gen_<variant>` with the sizes and sparsity tables of the variant and four static device
function templates over an accessor object `io`:

    values(io)   node values handed to the element assembly (eval_f, eval_g) + the node's checkpoints (see checkpoints())
    jac_part(w, io)   gradient of f and the two Jacobian blocks of the node (eval_grad_f, eval_jac_g), w-th code partition
    hess_part(w, io)  lower triangle of the node's Lagrangian-Hessian block + control-rate couplings (eval_h)
    all_part(w, io)   both in one pass
The derivative outputs of a node are split into NP_* partitions of balanced operation count (shared sub-expressions are
recomputed per partition); partition w is executed by warp w of a block, the 32 lanes of a warp being 32 nodes.

plus the host function `derive(D)` that fills the parameter-only sub-expressions after the raw vehicle parameters.
The code is what the CppAD tape sweeps of the reference compute (lion::ipopt_cppad_solve, call site
src/core/applications/optimal_laptime.hpp:546-551), differentiated at build time instead of interpreted at run time.
"""
import os
import re
import sys

from .nlpnode import Variant, NodeProgram
from .emit import Emitter, ParamTable, emit_host_derive

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "csrc", "generated")

# warps (= straight-line code partitions) per group of 32 nodes in each derivative kernel
NPARTS = {"jac": int(os.environ.get("FL_NP_JAC", 4)), "hess": int(os.environ.get("FL_NP_HESS", 8)), "all": int(os.environ.get("FL_NP_ALL", 8)),
          "values": int(os.environ.get("FL_NP_VALUES", 1)), "tape": int(os.environ.get("FL_NP_TAPE", 1))}

# arithmetic statements per fenced segment of a derivative partition (0 = no fences, all loads where the compiler puts them)
SEGMENT = int(os.environ.get("FL_SEGMENT", 0))

# statement order of the values pass (emit.py): "level" interleaves the independent chains of a node
VALUES_ORDER = os.environ.get("FL_VALUES_ORDER", "level")
DERIV_ORDER = os.environ.get("FL_DERIV_ORDER", "dfs")

# "G,KG,A": the per-node loads of the derivative partitions go through a shared-memory ring filled by cp.async (emit.py _ring):
# groups of G loads, KG groups in flight, values read A statements ahead of their first use; empty = plain loads
RING = tuple(int(v) for v in os.environ.get("FL_RING", "").split(",")) if os.environ.get("FL_RING") else None

# > 0: tape nodes that this many (or more) derivative partitions share are checkpointed by the values pass (see shared_tape)
SHARE = int(os.environ.get("FL_SHARE", 0))

VARIANTS = {
    "f1_direct": Variant("f1", "direct"),
    "f1_derivative": Variant("f1", "derivative"),
    "f1_direct_3d": Variant("f1", "direct", elevation=True),
    "f1_derivative_3d": Variant("f1", "derivative", elevation=True),
    "kart_direct": Variant("kart", "direct"),
    "kart_derivative": Variant("kart", "derivative"),
    "kart_direct_torque": Variant("kart", "direct", direct_torque=True),
    "kart_derivative_torque": Variant("kart", "derivative", direct_torque=True),
    "kart_steady": Variant("kart", "steady"),
    "f1_steady": Variant("f1", "steady"),
    "f1_direct_aug": Variant("f1", "direct", aug=True),
}


def _cone(g, outs, stop=()):
    """Operation nodes needed for `outs`, not descending below the nodes in `stop` (checkpoints are loaded, not recomputed)."""
    from .expr import CONST, SYM, NEG
    seen, res, stack = set(), set(), list(outs)
    while stack:
        n = stack.pop()
        if n in seen:
            continue
        seen.add(n)
        k = g.kind[n]
        if k in (CONST, SYM) or n in stop:
            continue
        if k != NEG:
            res.add(n)
        stack.extend(g.operands(n))
    return res


def checkpoints(g, pt, outs):
    """The expensive primal intermediates below the derivative outputs: every sin/cos/atan/tanh/exp/reciprocal/sqrt whose
    argument depends on the node variables.  Each expands to 10-80 SASS instructions; the values pass computes them once per
    node and stores them, the derivative partitions load them (one coalesced 8-byte load each) and are left with plain
    multiply-add arithmetic."""
    from .expr import SIN, COS, ATAN, TANH, EXP, RCP, SQRT
    heavy = (SIN, COS, ATAN, TANH, EXP, RCP, SQRT)
    return [n for n in g.topo(outs) if g.kind[n] in heavy and not pt.param_only(n)]


def _groups(p, out_grad, out_ja, out_jb, out_h, out_c, ja, jb, hs):
    """Output groups that share most of their expression cone: per variable k the k-th Jacobian column (+ df/dx_k) and the
    k-th column of the Hessian's lower triangle."""
    grp_j, grp_h = [], []
    for k in range(p.nv):
        grp_j.append([out_grad[k]] + [o for o, key in zip(out_ja, ja) if key[1] == k] + [o for o, key in zip(out_jb, jb) if key[1] == k])
        grp_h.append([o for o, key in zip(out_h, hs) if key[1] == k])
    grp_h.append(out_c)
    return grp_j, grp_h


def shared_tape(g, pt, groups, n_parts, heavy, share):
    """The part of the primal / adjoint tape that `share` or more derivative partitions would each recompute and keep in
    registers from its definition to its last use (the live set that spills, profiles/README.md): its frontier -- the
    shared nodes with a user outside the shared set -- becomes checkpoints too, computed once by the values pass."""
    from .expr import NEG
    grp_j, grp_h = groups
    outs, _ = partition(g, grp_j + grp_h, n_parts, stop=heavy)
    count = {}
    cones = [_cone(g, [e for e, _ in o], heavy) for o in outs]
    for c in cones:
        for n in c:
            count[n] = count.get(n, 0) + 1
    S = {n for n, c in count.items() if c >= share}
    users = {}
    for c in cones:
        for n in c:
            for o in g.operands(n):
                while g.kind[o] == NEG:
                    o = g.a[o]
                users.setdefault(o, set()).add(n)
    roots = {e for o in outs for e, _ in o}
    F = [n for n in sorted(S) if n in roots or any(u not in S for u in users.get(n, ()))]
    return F


def partition(g, groups, n_parts, stop=()):
    """Splits output groups (lists of (expr, statement) pairs) into `n_parts` lists so that the largest part's operation
    count -- the union of the expression cones of its outputs, shared sub-expressions counted once -- is small: groups are
    taken by decreasing cone size and each goes to the part whose cone grows to the smallest size.  One part = one warp
    of the derivative kernel: every warp runs different straight-line code over the same 32 nodes."""
    cones = sorted(((_cone(g, [e for e, _ in grp], stop), grp) for grp in groups if grp), key=lambda c: -len(c[0]))
    bins = [set() for _ in range(n_parts)]
    outs = [[] for _ in range(n_parts)]
    for c, grp in cones:
        b = min(range(n_parts), key=lambda i: (len(bins[i] | c), len(bins[i])))
        bins[b] |= c
        outs[b] += grp
    return outs, [len(b) for b in bins]


def _int_table(name, vals):
    body = ", ".join(str(v) for v in vals) if vals else "0"
    return "    static constexpr int %s[%d] = {%s};" % (name, max(len(vals), 1), body)


def generate(name, variant):
    p = NodeProgram(variant)
    g = p.g
    raw_names = p.param_names
    pt = ParamTable(g, raw_names)

    # symbol name -> local variable holding it
    inputs = {}
    pre = []
    for k in range(p.nv):
        inputs["x%d" % k] = "x%d" % k
        pre.append("const double x%d = io.x(%d);" % (k, k))
    tnames = p.track_names + p.geom_names
    for k, n in enumerate(tnames):
        inputs["t_" + n] = "t_" + n
        pre.append("const double t_%s = io.t(%d);" % (n, k))
    pre_w = ["const double obj = io.obj();"]
    inputs["obj"] = "obj"
    for r in range(p.ncpe):
        inputs["lin%d" % r] = "lin%d" % r
        pre_w.append("const double lin%d = io.lin(%d);" % (r, r))
        if p.lam_out[r] is not None:
            inputs["lout%d" % r] = "lout%d" % r
            pre_w.append("const double lout%d = io.lout(%d);" % (r, r))
    pre_n = []
    for c in range(len(p.uprev)):
        inputs["uprev%d" % c] = "uprev%d" % c
        inputs["unext%d" % c] = "unext%d" % c
        pre_n.append("const double uprev%d = io.uprev(%d);" % (c, c))
        pre_n.append("const double unext%d = io.unext(%d);" % (c, c))

    ja = sorted(p.jacA.keys())
    jb = sorted(p.jacB.keys())
    hs = sorted(p.hess.keys())

    out_values = [(e.i, "io.val(%d, %%s);" % k) for k, e in enumerate(p.values)]
    out_grad = [(p.gradf[k], "io.grad(%d, %%s);" % k) for k in range(p.nv)]
    out_ja = [(p.jacA[key], "io.ja(%d, %%s);" % k) for k, key in enumerate(ja)]
    out_jb = [(p.jacB[key], "io.jb(%d, %%s);" % k) for k, key in enumerate(jb)]
    out_h = [(p.hess[key], "io.h(%d, %%s);" % k) for k, key in enumerate(hs)]
    out_c = [(e, "io.cpl(%d, %%s);" % k) for k, e in enumerate(p.coupling)]

    # checkpoints: slot k of the per-node checkpoint buffer holds node ck[k]
    ck = checkpoints(g, pt, [e for e, _ in out_grad + out_ja + out_jb + out_h + out_c])
    n_heavy = len(ck)
    if SHARE > 0:
        # a full round (Jacobian AND Hessian) loads the shared tape too: it depends on the multipliers, so it is written by
        # the `tape` pass, which replaces the values pass in those rounds; Jacobian-only / Hessian-only rounds load slots
        # [0, NCKV) only
        ck = ck + shared_tape(g, pt, _groups(p, out_grad, out_ja, out_jb, out_h, out_c, ja, jb, hs), NPARTS["all"], set(ck), SHARE)
    cut_all = {n: "io.c(%d)" % k for k, n in enumerate(ck)}
    cut_heavy = {n: "io.c(%d)" % k for k, n in enumerate(ck[:n_heavy])}
    out_ck = [(n, "io.cs(%d, %%s);" % k) for k, n in enumerate(ck[:n_heavy])]
    out_ck_all = [(n, "io.cs(%d, %%s);" % k) for k, n in enumerate(ck)]

    def fn(fname, outputs, pres, use_cut=True, cut=None, order="dfs"):
        cut = cut_heavy if cut is None else cut
        lazy = {}
        if use_cut and (SEGMENT > 0 or RING):
            # derivative partitions: every input is loaded at first use and software-prefetched one segment ahead
            for l in sum(pres, []):
                m = re.match(r"const double (\w+) = (io\.\w+\(\d*\));", l)
                lazy[m.group(1)] = (m.group(1), m.group(2))
            pres = []
        em = Emitter(g, pt, inputs, dfmt="io.d(%d)", order=order, cut=cut if use_cut else None, lazy_inputs=lazy, segment=SEGMENT if use_cut else 0, ring=RING if use_cut else None)
        body = em.body(outputs, indent="        ")
        head = "\n".join("        " + l for l in sum(pres, []))
        return ("    template <class IO> static __device__ __forceinline__ void %s(IO& io)\n    {\n%s\n%s\n    }\n" % (fname, head, body))

    # output groups that share most of their expression cone: per variable k the k-th Jacobian column (+ df/dx_k) and the
    # k-th column of the Hessian's lower triangle
    grp_j, grp_h = _groups(p, out_grad, out_ja, out_jb, out_h, out_c, ja, jb, hs)
    parts = []
    part_sizes = {}
    # values pass: the node values and checkpoints, split the same way (every output is its own group; the pass is one
    # long dependent chain per node, so shorter partitions on more warps cut its latency)
    outs, sizes = partition(g, [[o] for o in out_values + out_ck], NPARTS["values"])
    part_sizes["values"] = sizes
    for w, o in enumerate(outs):
        parts.append(fn("values_p%d" % w, o, [pre], use_cut=False, order=VALUES_ORDER))
    disp = ["    template <class IO> static __device__ __forceinline__ void values_part(int w, IO& io)", "    {", "        switch (w)", "        {"]
    for w in range(len(outs)):
        disp.append("        case %d: values_p%d(io); break;" % (w, w))
    disp += ["        default: break;", "        }", "    }", ""]
    parts.append("\n".join(disp))
    # tape pass: the values pass of a full round = node values + every checkpoint, the shared tape included
    outs, sizes = partition(g, [[o] for o in out_values + out_ck_all], NPARTS["tape"] if SHARE > 0 else NPARTS["values"])
    part_sizes["tape"] = sizes
    n_tape = len(outs)
    if SHARE > 0:
        for w, o in enumerate(outs):
            parts.append(fn("tape_p%d" % w, o, [pre, pre_w, pre_n], use_cut=False, order=VALUES_ORDER))
    disp = ["    template <class IO> static __device__ __forceinline__ void tape_part(int w, IO& io)", "    {", "        switch (w)", "        {"]
    for w in range(len(outs)):
        disp.append("        case %d: %s_p%d(io); break;" % (w, "tape" if SHARE > 0 else "values", w))
    disp += ["        default: break;", "        }", "    }", ""]
    parts.append("\n".join(disp))
    for mode, groups, pres in (("jac", grp_j, [pre, pre_n]), ("hess", grp_h, [pre, pre_w]), ("all", grp_j + grp_h, [pre, pre_w, pre_n])):
        mcut = cut_all if mode == "all" else cut_heavy
        outs, sizes = partition(g, groups, NPARTS[mode], stop=set(mcut))
        part_sizes[mode] = sizes
        for w, o in enumerate(outs):
            parts.append(fn("%s_p%d" % (mode, w), o, pres, cut=mcut, order=DERIV_ORDER))
        disp = ["    template <class IO> static __device__ __forceinline__ void %s_part(int w, IO& io)" % mode, "    {", "        switch (w)", "        {"]
        for w in range(len(outs)):
            disp.append("        case %d: %s_p%d(io); break;" % (w, mode, w))
        disp += ["        default: break;", "        }", "    }", ""]
        parts.append("\n".join(disp))
    # output channels of the model (get_outputs_map of the reference): their own small program, values only
    out_channels = list(getattr(p, "outputs", []))
    if out_channels:
        parts.append(fn("outputs", [(e.i, "io.out(%d, %%s);" % k) for k, (_, e) in enumerate(out_channels)], [pre], use_cut=False, order=VALUES_ORDER))
    else:
        parts.append("    template <class IO> static __device__ __forceinline__ void outputs(IO&) {}\n")
    host = emit_host_derive(g, pt, "derive")
    host = "    static inline " + host.replace("\n", "\n    ")

    counts_all = g.count_ops([e for e, _ in out_values + out_grad + out_ja + out_jb + out_h + out_c])
    counts_val = g.count_ops([e for e, _ in out_values])
    counts_jac = g.count_ops([e for e, _ in out_grad + out_ja + out_jb])
    counts_hes = g.count_ops([e for e, _ in out_h + out_c])

    L = []
    L.append("// GENERATED by fastest_lap_b200/codegen/generate.py (variant %s) -- do not edit, regenerate." % name)
    L.append("// One collocation node of the %s model, %s form%s%s." % (variant.model, variant.form, ", track with elevation" if variant.elevation else "",
                                                                     ", direct axle torque" if variant.direct_torque else ""))
    L.append("// fp64 operation counts (after common sub-expression elimination):")
    for nm, c in (("values", counts_val), ("jac", counts_jac), ("hess", counts_hes), ("all", counts_all)):
        L.append("//   %-6s total %5d : %s" % (nm, sum(c.values()), ", ".join("%s %d" % kv for kv in sorted(c.items()))))
    L.append("#pragma once")
    L.append("struct gen_%s" % name)
    L.append("{")
    L.append("    static constexpr const char* name = \"%s\";" % name)
    L.append("    static constexpr int NV = %d, NCPE = %d, NTD = %d, NALG = %d, NEXT = %d, NCR = %d, NFM = %d;" %
             (p.nv, p.ncpe, p.n_td, p.n_alg, p.n_ext, p.n_ctrl_rows, p.n_fm))
    L.append("    static constexpr int NVAL = %d, NTC = %d, NTRACK = %d, NRAW = %d, ND = %d, NCK = %d;" % (len(p.values), len(tnames), len(p.track_names), len(raw_names), pt.size, len(ck)))
    L.append("    static constexpr int NJA = %d, NJB = %d, NH = %d, NCPL = %d;" % (len(ja), len(jb), len(hs), len(p.coupling)))
    L.append("    static constexpr int IS_DERIVATIVE = %d, IS_3D = %d;" % (int(variant.form == "derivative"), int(variant.elevation)))
    L.append("    static constexpr int IS_STEADY = %d;   // one-node problems (steady-state NLPs): the kernels map a thread to a PROBLEM of the batch" % int(variant.form == "steady"))
    n_aug = getattr(p, "n_aug", 0)
    L.append("    static constexpr int NAUG = %d, AUG_POS = %d, P_AUG = %d;   // augmented states (the last NAUG trapezoid rows); node variable a (q follows); first aug_* parameter" %
             (n_aug, p.aug_pos[0] if n_aug else -1, raw_names.index("aug_bb") if n_aug else -1))
    L.append("    static constexpr int NP_VALUES = %d, NP_JAC = %d, NP_HESS = %d, NP_ALL = %d;   // code partitions of each pass" % (NPARTS["values"], NPARTS["jac"], NPARTS["hess"], NPARTS["all"]))
    L.append("    static constexpr int RING_SLOTS = %d;   // shared-memory ring slots per thread of the derivative partitions (0: plain loads)" % ((RING[1] + 1) * RING[0] if RING else 0))
    L.append("    static constexpr int NCKV = %d, HAS_TAPE = %d, NP_TAPE = %d;   // checkpoints of the values pass; the tape pass writes all NCK" % (n_heavy, int(SHARE > 0), n_tape))
    for mode in ("values", "tape", "jac", "hess", "all"):
        L.append("    // %s partition operation counts: %s (sum %d)" % (mode, part_sizes[mode], sum(part_sizes[mode])))
    L.append("    static constexpr int OPS_VALUES = %d, OPS_JAC = %d, OPS_HESS = %d, OPS_ALL = %d;" %
             (sum(counts_val.values()), sum(counts_jac.values()), sum(counts_hes.values()), sum(counts_all.values())))
    L.append("    static constexpr int NOUT = %d;   // named output channels (outputs())" % len(out_channels))
    L.append("    static constexpr const char* OUT_NAMES[%d] = { %s };" % (max(1, len(out_channels)), ", ".join('"%s"' % nm for nm, _ in out_channels) or '""'))
    L.append(_int_table("CTRL_POS", p.ctrl_pos))
    L.append(_int_table("DCTRL_POS", p.dctrl_pos))
    L.append("    static __host__ __device__ constexpr int ctrl_pos(int c) { return c == 0 ? %d : %d; }" % tuple(p.ctrl_pos))
    L.append("    static __host__ __device__ constexpr int dctrl_pos(int c) { return c == 0 ? %d : %d; }" % (tuple(p.dctrl_pos) if p.dctrl_pos else (0, 0)))
    L.append(_int_table("JA_ROW", [r for r, _ in ja]))
    L.append(_int_table("JA_COL", [c for _, c in ja]))
    L.append(_int_table("JB_ROW", [r for r, _ in jb]))
    L.append(_int_table("JB_COL", [c for _, c in jb]))
    L.append(_int_table("H_ROW", [r for r, _ in hs]))
    L.append(_int_table("H_COL", [c for _, c in hs]))
    L.append("    // slots of the raw parameters that the host fills besides the vehicle's: sigma, diss0, diss1, wheel scales")
    for nm in ("sigma", "diss0", "diss1", "f_wheel_scale", "r_wheel_scale"):
        if nm in raw_names:
            L.append("    static constexpr int P_%s = %d;" % (nm.upper(), raw_names.index(nm)))
    L.append("")
    L.extend(parts)
    L.append(host)
    L.append("};")
    text = "\n".join(L) + "\n"
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, "gen_%s.cuh" % name)
    old = open(path).read() if os.path.exists(path) else None
    if old != text:
        with open(path, "w") as fh:
            fh.write(text)
    return dict(name=name, nv=p.nv, ncpe=p.ncpe, nja=len(ja), njb=len(jb), nh=len(hs), ops=sum(counts_all.values()), nd=pt.size)


def main(argv):
    names = argv or list(VARIANTS)
    for n in names:
        info = generate(n, VARIANTS[n])
        print(info)


if __name__ == "__main__":
    main(sys.argv[1:])
