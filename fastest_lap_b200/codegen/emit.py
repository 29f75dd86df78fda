"""Straight-line fp64 CUDA / C emission of expression DAGs (see expr.py)."""
from .expr import CONST, SYM, ADD, SUB, MUL, NEG, RCP, SQRT, SIN, COS, ATAN, TANH, EXP

import os

_FN = {SQRT: "sqrt", SIN: "sin", COS: "cos", ATAN: "atan", TANH: "tanh", EXP: "exp"}
# device code: the branch-free versions of csrc/fl_math.cuh (FL_LIBM=1 keeps the CUDA math library, for comparison)
LIBM = int(os.environ.get("FL_LIBM", 0))
_FN_DEV = _FN if LIBM else {k: "fl_" + v for k, v in _FN.items()}


def lit(v):
    s = repr(float(v))
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s


class ParamTable:
    """Raw parameters followed by host-derived ones (sub-expressions that depend on parameters only)."""

    def __init__(self, graph, raw_names):
        self.g = graph
        self.raw_names = list(raw_names)
        self.index = {}          # node id -> slot in D
        for k, n in enumerate(self.raw_names):
            self.index[graph.sym("p_" + n, "p")] = k
        self.derived = []        # node ids, in slot order after the raw ones
        self._ponly = {}

    def param_only(self, i):
        g = self.g
        r = self._ponly.get(i)
        if r is None:
            k = g.kind[i]
            if k == CONST:
                r = True
            elif k == SYM:
                r = g.sym_class[g.a[i]] == "p"
            else:
                r = all(self.param_only(o) for o in g.operands(i))
            self._ponly[i] = r
        return r

    def slot(self, i):
        s = self.index.get(i)
        if s is None:
            s = len(self.raw_names) + len(self.derived)
            self.derived.append(i)
            self.index[i] = s
        return s

    @property
    def size(self):
        return len(self.raw_names) + len(self.derived)


class Emitter:
    """Emits one function body: `inputs` maps symbol name -> C expression; outputs are (expr id, "statement with one %s") pairs."""

    def __init__(self, graph, ptable, inputs, dfmt="D[%d]", cut=None, lazy_inputs=None, segment=0, order="dfs", ring=None):
        # cut: node id -> C expression that LOADS the node's value (a checkpoint written by an earlier pass) instead of
        # recomputing it; the load is emitted lazily, right before the first use
        # lazy_inputs: symbol name -> (local name, C expression that loads it), emitted at first use instead of up front
        # segment: > 0 splits the body into segments of that many arithmetic statements separated by `io.fence(t)`; the
        #          loads a segment needs are issued at the start of the PREVIOUS segment (software prefetch into registers)
        self.g, self.pt, self.inputs, self.dfmt, self.cut = graph, ptable, inputs, dfmt, (cut or {})
        self.lazy_inputs, self.segment = (lazy_inputs or {}), int(segment)
        # order: "dfs" = each output's expression tree in post-order (short live ranges, consecutive statements depend on each
        # other); "level" = by latency-weighted depth (independent chains, e.g. the four tyres, interleaved statement by
        # statement, so that a lone warp has several instructions in flight)
        self.order = order
        # ring: (loads per group, groups in flight, statements a value is read ahead of its first use) -- the per-node loads
        # (checkpoints, variables, multipliers, track constants) travel through a small per-thread ring in shared memory filled
        # by cp.async several groups ahead of their use, see _ring()
        self.ring = ring

    def body(self, outputs, indent="    "):
        g, pt = self.g, self.pt
        lines = []
        is_load = set()     # indices into `lines` of statements that only load a value
        name = {}
        K, A, B = g.kind, g.a, g.b

        # which sin/cos share an argument (-> sincos)
        needed = g.topo([e for e, _ in outputs])
        sin_of, cos_of = {}, {}
        for i in needed:
            if pt.param_only(i) or i in self.cut:
                continue
            if K[i] == SIN:
                sin_of[A[i]] = i
            elif K[i] == COS:
                cos_of[A[i]] = i
        pair = {a: (sin_of[a], cos_of[a]) for a in sin_of if a in cos_of}

        def ref(i):
            """C expression for an already emitted node (negations and literals are folded into the consumer)."""
            k = K[i]
            if k == CONST:
                v = A[i]
                return lit(v) if v >= 0 else "(" + lit(v) + ")"
            if k == NEG and not pt.param_only(i):
                return "(-" + ref(A[i]) + ")"
            return name[i]

        def emit(i):
            # iterative post-order DFS
            stack = [(i, False)]
            while stack:
                n, done = stack.pop()
                if n in name or K[n] == CONST:
                    continue
                if n in self.cut:
                    name[n] = "t%d" % n
                    is_load.add(len(lines))
                    lines.append("%sconst double t%d = %s;" % (indent, n, self.cut[n]))
                    continue
                if K[n] != SYM and pt.param_only(n):
                    s = pt.slot(n)
                    name[n] = "d%d" % s
                    is_load.add(len(lines))
                    lines.append("%sconst double d%d = %s;" % (indent, s, self.dfmt % s))
                    continue
                if K[n] == SYM:
                    cls = g.sym_class[A[n]]
                    if cls == "p":
                        s = pt.slot(n)
                        name[n] = "d%d" % s
                        is_load.add(len(lines))
                        lines.append("%sconst double d%d = %s;" % (indent, s, self.dfmt % s))
                    elif A[n] in self.lazy_inputs:
                        local, load = self.lazy_inputs[A[n]]
                        name[n] = local
                        is_load.add(len(lines))
                        lines.append("%sconst double %s = %s;" % (indent, local, load))
                    else:
                        name[n] = self.inputs[A[n]]
                    continue
                if not done:
                    stack.append((n, True))
                    for o in reversed(g.operands(n)):
                        if o not in name and K[o] != CONST:
                            stack.append((o, False))
                    continue
                k = K[n]
                if k == NEG:
                    name[n] = None     # folded by ref()
                    continue
                t = "t%d" % n
                if k == ADD:
                    rhs = "%s + %s" % (ref(A[n]), ref(B[n]))
                elif k == SUB:
                    rhs = "%s - %s" % (ref(A[n]), ref(B[n]))
                elif k == MUL:
                    rhs = "%s * %s" % (ref(A[n]), ref(B[n]))
                elif k == RCP:
                    rhs = ("1.0 / %s" if LIBM else "fl_rcp(%s)") % ref(A[n])
                elif k in (SIN, COS) and A[n] in pair:
                    s_id, c_id = pair[A[n]]
                    if s_id not in name and c_id not in name:
                        lines.append("%sdouble t%d, t%d; %s(%s, &t%d, &t%d);" % (indent, s_id, c_id, "sincos" if LIBM else "fl_sincos", ref(A[n]), s_id, c_id))
                        name[s_id], name[c_id] = "t%d" % s_id, "t%d" % c_id
                    continue
                else:
                    rhs = "%s(%s)" % (_FN_DEV[k], ref(A[n]))
                lines.append("%sconst double %s = %s;" % (indent, t, rhs))
                name[n] = t

        if self.order == "level":
            weight = {RCP: 8, SQRT: 10, SIN: 25, COS: 25, ATAN: 30, TANH: 35, EXP: 20}
            level = {}
            for i in needed:                                   # topological order
                if K[i] in (CONST, SYM) or i in self.cut or pt.param_only(i):
                    level[i] = 0
                else:
                    level[i] = max(level[o] for o in g.operands(i)) + (0 if K[i] == NEG else weight.get(K[i], 1))
            stores = {}
            for e, fmt in outputs:
                stores.setdefault(e, []).append(fmt)
            # the nodes to compute: reachable from the outputs without going below a loaded value
            cone, stack = set(), [e for e, _ in outputs]
            while stack:
                i = stack.pop()
                if i in cone:
                    continue
                cone.add(i)
                if not (K[i] in (CONST, SYM) or i in self.cut or pt.param_only(i)):
                    stack.extend(g.operands(i))
            for i in sorted(cone, key=lambda n: (level[n], n)):
                emit(i)
                for fmt in stores.get(i, ()):
                    lines.append(indent + (fmt % ref(i)))
        else:
            for e, fmt in outputs:
                emit(e)
                lines.append(indent + (fmt % ref(e)))
        if self.segment > 0:
            lines = self._pipeline(lines, is_load, indent)
        if self.ring:
            lines = self._ring(lines, indent)
        return "\n".join(lines)

    def _ring(self, lines, indent):
        """Asynchronous loads through shared memory.  Every per-node load `const double v = io.c(k);` (also x, t, lin, lout, uprev,
        unext) becomes a pair: `io.rput(slot, io.pc(k))` -- an 8-byte cp.async from global to the thread's ring slot, issued KG groups
        of G loads before the value is needed -- and `const double v = io.rget(slot);`, a shared-memory read placed `ahead` statements
        before the first use.  The global-memory latency (an L2 round trip per checkpoint, profiles/README.md) is then covered by
        the arithmetic in between without holding a register per load in flight."""
        import re
        G, KG, ahead = self.ring
        pat = re.compile(r"\s*const double (\w+) = io\.(c|x|t|lin|lout|uprev|unext)\((\d+)\);$")
        loads = []          # (line index, variable, pointer expression)
        for idx, ln in enumerate(lines):
            m = pat.match(ln)
            if m:
                loads.append((idx, m.group(1), "io.p%s(%s)" % (m.group(2), m.group(3))))
        if not loads:
            return lines
        L = len(loads)
        NG = (L + G - 1) // G
        nslot = (KG + 1) * G
        slot = lambda i: i % nslot

        def issue(j):
            out = ["%sio.rput(%d, %s);" % (indent, slot(i), loads[i][2]) for i in range(j * G, min(L, (j + 1) * G))]
            return out + [indent + "io.rcommit();"]

        # where each read goes: `ahead` statements before the original load, never before an earlier read
        pos, last = [], 0
        for idx, _, _ in loads:
            last = max(last, idx - ahead)
            pos.append(last)
        at = {}
        for i, q in enumerate(pos):
            at.setdefault(q, []).append(i)
        skip = {idx for idx, _, _ in loads}
        out = []
        for j in range(min(KG, NG)):
            out += issue(j)
        for idx, ln in enumerate(lines):
            for i in at.get(idx, ()):
                j = i // G
                if i % G == 0:
                    out.append("%sFL_RING_WAIT(%d);" % (indent, min(KG - 1, NG - 1 - j)))
                out.append("%sconst double %s = io.rget(%d);" % (indent, loads[i][1], slot(i)))
                if (i % G == G - 1 or i == L - 1) and j + KG < NG:
                    out += issue(j + KG)
            if idx not in skip:
                out.append(ln)
        return out

    def _pipeline(self, lines, is_load, indent):
        """Segments of `self.segment` arithmetic statements; segment k's loads move to the start of segment k-1, and every
        boundary stores the value computed last through `io.fence` (a store the compiler must assume may alias the inputs,
        so it cannot hoist the following loads above it -- it otherwise moves every load to the top and spills)."""
        import re
        segs, cur, nops, last_tmp = [[]], [], 0, None
        loads = [[]]
        for idx, ln in enumerate(lines):
            if idx in is_load:
                loads[-1].append(ln)
                continue
            segs[-1].append(ln)
            m = re.match(r"\s*const double (t\d+) = ", ln)
            if m:
                nops += 1
                last_tmp = m.group(1)
                if nops % self.segment == 0:
                    segs[-1].append("%sio.fence(%s);" % (indent, last_tmp))
                    segs.append([])
                    loads.append([])
        out = []
        for k in range(len(segs)):
            if k == 0:
                out += loads[0]
            if k + 1 < len(segs):
                out += loads[k + 1]
            out += segs[k]
        return out


def emit_host_derive(graph, ptable, fname):
    """C function computing the derived-parameter slots from the raw ones (run once per vehicle on the host)."""
    g = graph
    K, A, B = g.kind, g.a, g.b
    lines = ["void %s(double* D)" % fname, "{"]
    name = {}
    raw = {i: s for i, s in ptable.index.items() if s < len(ptable.raw_names)}

    def ref(i):
        if K[i] == CONST:
            v = A[i]
            return lit(v) if v >= 0 else "(" + lit(v) + ")"
        return name[i]

    for i in g.topo(list(ptable.derived)):
        k = K[i]
        if k == CONST:
            continue
        if k == SYM:
            name[i] = "D[%d]" % raw[i]
            continue
        t = "h%d" % i
        if k == ADD:
            rhs = "%s + %s" % (ref(A[i]), ref(B[i]))
        elif k == SUB:
            rhs = "%s - %s" % (ref(A[i]), ref(B[i]))
        elif k == MUL:
            rhs = "%s * %s" % (ref(A[i]), ref(B[i]))
        elif k == NEG:
            rhs = "-%s" % ref(A[i])
        elif k == RCP:
            rhs = "1.0 / %s" % ref(A[i])
        else:
            rhs = "%s(%s)" % (_FN[k], ref(A[i]))
        lines.append("    const double %s = %s;" % (t, rhs))
        name[i] = t
    for i in ptable.derived:
        lines.append("    D[%d] = %s;" % (ptable.index[i], ref(i)))
    lines.append("}")
    return "\n".join(lines)
