"""Hash-consed expression DAG with symbolic forward / reverse differentiation.

This is the build-time replacement for the CppAD tape of the reference (lion::ipopt_cppad_solve, call site
src/core/applications/optimal_laptime.hpp:546-551): the vehicle model is traced ONCE here, differentiated symbolically
(reverse sweep for the Lagrangian gradient, forward sweeps over it for the Hessian columns; forward sweeps for the
Jacobian rows), simplified (constant folding, zero/one propagation, common sub-expressions by hash-consing) and emitted
as straight-line fp64 CUDA by emit.py.  Nothing here runs at solve time.
"""
import math

# node kinds
CONST, SYM, ADD, SUB, MUL, NEG, RCP, SQRT, SIN, COS, ATAN, TANH, EXP = range(13)
KIND_NAMES = ["const", "sym", "add", "sub", "mul", "neg", "rcp", "sqrt", "sin", "cos", "atan", "tanh", "exp"]
UNARY = (NEG, RCP, SQRT, SIN, COS, ATAN, TANH, EXP)


class Graph:
    def __init__(self):
        self.kind = []      # node kind
        self.a = []         # first operand id / const value / symbol name
        self.b = []         # second operand id
        self._memo = {}
        self.sym_class = {}  # symbol name -> class ("x", "p", "t", "w", ...)

    # ---- construction -------------------------------------------------------------------------------------------
    def _new(self, kind, a, b=None):
        key = (kind, a, b)
        i = self._memo.get(key)
        if i is None:
            i = len(self.kind)
            self.kind.append(kind)
            self.a.append(a)
            self.b.append(b)
            self._memo[key] = i
        return i

    def const(self, v):
        v = float(v)
        if v == 0.0:
            v = 0.0  # merge -0.0
        return self._new(CONST, v)

    def sym(self, name, cls):
        self.sym_class[name] = cls
        return self._new(SYM, name)

    def is_const(self, i):
        return self.kind[i] == CONST

    def cval(self, i):
        return self.a[i]

    def add(self, x, y):
        K = self.kind
        if K[x] == CONST and K[y] == CONST:
            return self.const(self.a[x] + self.a[y])
        if K[x] == CONST and self.a[x] == 0.0:
            return y
        if K[y] == CONST and self.a[y] == 0.0:
            return x
        if K[y] == NEG:
            return self.sub(x, self.a[y])
        if K[x] == NEG:
            return self.sub(y, self.a[x])
        if x > y:
            x, y = y, x
        return self._new(ADD, x, y)

    def sub(self, x, y):
        K = self.kind
        if x == y:
            return self.const(0.0)
        if K[x] == CONST and K[y] == CONST:
            return self.const(self.a[x] - self.a[y])
        if K[y] == CONST and self.a[y] == 0.0:
            return x
        if K[x] == CONST and self.a[x] == 0.0:
            return self.neg(y)
        if K[y] == NEG:
            return self.add(x, self.a[y])
        return self._new(SUB, x, y)

    def neg(self, x):
        K = self.kind
        if K[x] == CONST:
            return self.const(-self.a[x])
        if K[x] == NEG:
            return self.a[x]
        if K[x] == SUB:
            return self.sub(self.b[x], self.a[x])
        return self._new(NEG, x)

    def mul(self, x, y):
        K = self.kind
        if K[x] == CONST and K[y] == CONST:
            return self.const(self.a[x] * self.a[y])
        if K[y] == CONST:
            x, y = y, x
        if K[x] == CONST:
            c = self.a[x]
            if c == 0.0:
                return x
            if c == 1.0:
                return y
            if c == -1.0:
                return self.neg(y)
        # pull negations out so that CSE sees through signs
        if K[x] == NEG and K[y] == NEG:
            return self.mul(self.a[x], self.a[y])
        if K[x] == NEG:
            return self.neg(self.mul(self.a[x], y))
        if K[y] == NEG:
            return self.neg(self.mul(x, self.a[y]))
        if x > y:
            x, y = y, x
        return self._new(MUL, x, y)

    def rcp(self, x):
        K = self.kind
        if K[x] == CONST:
            return self.const(1.0 / self.a[x])
        if K[x] == NEG:
            return self.neg(self.rcp(self.a[x]))
        if K[x] == RCP:
            return self.a[x]
        return self._new(RCP, x)

    def div(self, x, y):
        if self.kind[y] == CONST:
            return self.mul(x, self.const(1.0 / self.a[y]))
        return self.mul(x, self.rcp(y))

    def unary(self, kind, x):
        if kind == NEG:
            return self.neg(x)
        if kind == RCP:
            return self.rcp(x)
        K = self.kind
        if K[x] == CONST:
            f = {SQRT: math.sqrt, SIN: math.sin, COS: math.cos, ATAN: math.atan, TANH: math.tanh, EXP: math.exp}[kind]
            return self.const(f(self.a[x]))
        if K[x] == NEG:
            if kind in (SIN, ATAN, TANH):
                return self.neg(self._new(kind, self.a[x]))
            if kind == COS:
                return self._new(kind, self.a[x])
        return self._new(kind, x)

    # ---- traversal ----------------------------------------------------------------------------------------------
    def operands(self, i):
        k = self.kind[i]
        if k in (CONST, SYM):
            return ()
        if k in (ADD, SUB, MUL):
            return (self.a[i], self.b[i])
        return (self.a[i],)

    def topo(self, roots):
        """Nodes reachable from roots, operands first (ids are already topologically ordered by construction)."""
        seen = set()
        stack = list(roots)
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            stack.extend(self.operands(i))
        return sorted(seen)

    # ---- differentiation ----------------------------------------------------------------------------------------
    def local_partials(self, i):
        """[(operand, d node / d operand)] as expression ids."""
        k, a, b = self.kind[i], self.a[i], self.b[i]
        one = self.const(1.0)
        if k == ADD:
            return [(a, one), (b, one)]
        if k == SUB:
            return [(a, one), (b, self.const(-1.0))]
        if k == MUL:
            if a == b:
                return [(a, self.mul(self.const(2.0), a))]
            return [(a, b), (b, a)]
        if k == NEG:
            return [(a, self.const(-1.0))]
        if k == RCP:
            return [(a, self.neg(self.mul(i, i)))]
        if k == SQRT:
            return [(a, self.mul(self.const(0.5), self.rcp(i)))]
        if k == SIN:
            return [(a, self.unary(COS, a))]
        if k == COS:
            return [(a, self.neg(self.unary(SIN, a)))]
        if k == ATAN:
            return [(a, self.rcp(self.add(one, self.mul(a, a))))]
        if k == TANH:
            return [(a, self.sub(one, self.mul(i, i)))]
        if k == EXP:
            return [(a, i)]
        return []

    def forward(self, roots, wrt, memo=None):
        """d root / d wrt for every root (wrt is one symbol node id).  Returns list of expression ids."""
        if memo is None:
            memo = {}
        zero = self.const(0.0)
        order = self.topo(roots)
        for i in order:
            if i in memo:
                continue
            k = self.kind[i]
            if k == CONST:
                memo[i] = zero
            elif k == SYM:
                memo[i] = self.const(1.0) if i == wrt else zero
            else:
                acc = zero
                for (op, part) in self.local_partials(i):
                    d = memo[op]
                    if d != zero:
                        acc = self.add(acc, self.mul(part, d))
                memo[i] = acc
        return [memo[r] for r in roots]

    def depends(self, roots, symset):
        """Set of nodes (reachable from roots) that depend on any symbol in symset."""
        dep = set()
        for i in self.topo(roots):
            k = self.kind[i]
            if k == SYM:
                if i in symset:
                    dep.add(i)
            elif k != CONST:
                if any(o in dep for o in self.operands(i)):
                    dep.add(i)
        return dep

    def reverse(self, root, wrt_list):
        """Gradient of one scalar root w.r.t. the symbols in wrt_list (reverse sweep)."""
        zero = self.const(0.0)
        order = self.topo([root])
        active = self.depends([root], set(wrt_list))
        adj = {root: self.const(1.0)}
        for i in reversed(order):
            if i not in adj or i not in active:
                continue
            bar = adj[i]
            if bar == zero:
                continue
            for (op, part) in self.local_partials(i):
                if op not in active:
                    continue
                contrib = self.mul(bar, part)
                adj[op] = self.add(adj[op], contrib) if op in adj else contrib
        return [adj.get(w, zero) for w in wrt_list]

    # ---- evaluation (used by the generator's own self-checks) ------------------------------------------------------
    def evaluate(self, roots, env):
        val = {}
        fn = {SQRT: math.sqrt, SIN: math.sin, COS: math.cos, ATAN: math.atan, TANH: math.tanh, EXP: math.exp}
        for i in self.topo(roots):
            k, a, b = self.kind[i], self.a[i], self.b[i]
            if k == CONST:
                val[i] = a
            elif k == SYM:
                val[i] = env[a]
            elif k == ADD:
                val[i] = val[a] + val[b]
            elif k == SUB:
                val[i] = val[a] - val[b]
            elif k == MUL:
                val[i] = val[a] * val[b]
            elif k == NEG:
                val[i] = -val[a]
            elif k == RCP:
                val[i] = 1.0 / val[a]
            else:
                val[i] = fn[k](val[a])
        return [val[r] for r in roots]

    def count_ops(self, roots, leaf_classes=("p",)):
        counts = {}
        for i in self.topo(roots):
            k = self.kind[i]
            if k in (CONST, SYM):
                continue
            counts[KIND_NAMES[k]] = counts.get(KIND_NAMES[k], 0) + 1
        return counts


class E:
    """Operator-overloading handle on a Graph node, so that the models read like the reference's templated C++."""
    __slots__ = ("g", "i")

    def __init__(self, g, i):
        self.g = g
        self.i = i

    def _c(self, o):
        return o if isinstance(o, E) else E(self.g, self.g.const(o))

    def __add__(self, o):
        return E(self.g, self.g.add(self.i, self._c(o).i))
    __radd__ = __add__

    def __sub__(self, o):
        return E(self.g, self.g.sub(self.i, self._c(o).i))

    def __rsub__(self, o):
        return E(self.g, self.g.sub(self._c(o).i, self.i))

    def __mul__(self, o):
        return E(self.g, self.g.mul(self.i, self._c(o).i))
    __rmul__ = __mul__

    def __truediv__(self, o):
        return E(self.g, self.g.div(self.i, self._c(o).i))

    def __rtruediv__(self, o):
        return E(self.g, self.g.div(self._c(o).i, self.i))

    def __neg__(self):
        return E(self.g, self.g.neg(self.i))


def _un(kind):
    def f(x):
        return E(x.g, x.g.unary(kind, x.i))
    return f


sqrt, sin, cos, atan, tanh, exp = _un(SQRT), _un(SIN), _un(COS), _un(ATAN), _un(TANH), _un(EXP)
