"""TEST INFRASTRUCTURE ONLY -- numpy statement of the block-cyclic-tridiagonal KKT factorisation the device solver runs
(fastest_lap_b200/csrc/kkt_kernels.cuh).

The reference leaves the linear algebra of every interior-point iteration to Ipopt + MUMPS (third-party, absent from the
reference tree; call site /root/reference/src/core/applications/optimal_laptime.hpp:546-551).  The augmented system of the
collocation NLP,

        [ W + Sigma_x + dw I      J^T ] [dx]   [r1]
        [ J                       -D  ] [dy] = [r2],        D = dc (equality rows),  1/(Sigma_s + dw) + dc (inequality rows),

has a chain structure: W is block diagonal per node plus control-control couplings between neighbouring
nodes, and the rows of element e (node e -> node e+1) touch only those two nodes.  With the inequality rows (which touch
the element's right node only) condensed into that node's diagonal block, and the unknowns grouped per stage
z_j = (dx_j, dy of the equality rows of the element that ENDS at node j), the matrix is symmetric block tridiagonal --
cyclic for closed laps -- with blocks of NV + n_eq (28 for the F1 direct form).  Stages are eliminated from the last one
backwards (the order of a Riccati recursion: the Schur complement handed to stage j-1 is the cost-to-go Hessian of node
j-1), stage 0 last; the closing element of a closed lap makes every stage couple with stage 0 through a fill block F_j.

Every pivot block is inverted by Gauss-Jordan sweeps with Bunch-Kaufman pivoting, which also counts negative pivots.
Inertia: by Haynsworth additivity the inertia of the whole matrix is the sum of the inertias of the pivot blocks, so the
factorisation reports exactly what Ipopt asks of MUMPS (n positive, m negative eigenvalues, none zero).
"""
import numpy as np
import scipy.sparse as sp


class StageLayout:
    """Index bookkeeping of one lap: which (variable / equality-row / inequality-row) belongs to which stage."""

    def __init__(self, n_points, n_elements, nv, ncpe, eq_rows, closed):
        self.n_points, self.n_elements, self.nv, self.ncpe, self.closed = n_points, n_elements, nv, ncpe, bool(closed)
        self.eq_rows = np.asarray(eq_rows)
        self.ineq_rows = np.array([r for r in range(ncpe) if r not in set(self.eq_rows.tolist())])
        self.neq = len(self.eq_rows)
        self.nb = nv + self.neq
        self.n_stages = n_elements                      # closed: one per node; open: one per free node (node 0 is fixed)
        self.n, self.m = n_elements * nv, n_elements * ncpe
        # element e ends at free node index: closed -> (e + 1) % N ; open -> e (free nodes are 1..N-1, numbered 0..N-2)
        st_of_el = (np.arange(n_elements) + 1) % n_points if closed else np.arange(n_elements)
        self.stage_of_element = st_of_el
        xs = np.arange(self.n).reshape(n_elements, nv)
        rows = np.arange(self.m).reshape(n_elements, ncpe)
        self.x_index = xs                                # [stage][nv]  (variable node == stage)
        el_of_stage = np.empty(n_elements, int)
        el_of_stage[st_of_el] = np.arange(n_elements)
        self.element_of_stage = el_of_stage
        self.eq_index = rows[el_of_stage][:, self.eq_rows]        # [stage][neq]  global constraint rows
        self.ineq_index = rows[el_of_stage][:, self.ineq_rows]    # [stage][nineq]
        self.all_eq = self.eq_index.ravel()
        self.all_ineq = self.ineq_index.ravel()
        # permutation of the reduced unknown vector (dx, dy_eq) into stage order
        self.perm = np.concatenate([np.concatenate([xs[j], self.n + np.arange(j * self.neq, (j + 1) * self.neq)]) for j in range(n_elements)])


def assemble_blocks(lay, H, J, sig_x, D):
    """Diagonal blocks Dg[stage] and sub-diagonal blocks Lo[stage] = K[stage, stage-1] (Lo[0] = corner K[0, last] of closed
    laps, zero otherwise) of the condensed KKT matrix.  H: n x n symmetric sparse; J: m x n sparse; sig_x: [n] diagonal
    added to H; D: [m] (see module docstring).  Also returns the operator that condenses a right-hand side."""
    n, neq, nb, S = lay.n, lay.neq, lay.nb, lay.n_stages
    Jc = sp.csr_matrix(J)
    Ji = Jc[lay.all_ineq]
    Je = Jc[lay.all_eq]
    Dinv_i = 1.0 / D[lay.all_ineq]
    W = sp.csr_matrix(H) + sp.diags(sig_x) + Ji.T @ sp.diags(Dinv_i) @ Ji
    K = sp.bmat([[W, Je.T], [Je, -sp.diags(D[lay.all_eq])]], format="csr")
    Kp = K[lay.perm][:, lay.perm].tobsr(blocksize=(nb, nb))
    Kp.sort_indices()
    Dg = np.zeros((S, nb, nb))
    Lo = np.zeros((S, nb, nb))
    indptr, indices, data = Kp.indptr, Kp.indices, Kp.data
    for j in range(S):
        for p in range(indptr[j], indptr[j + 1]):
            c = indices[p]
            if c == j:
                Dg[j] = data[p]
            elif c == j - 1:
                Lo[j] = data[p]
            elif j == 0 and c == S - 1 and S > 2:
                Lo[0] = data[p]
            elif c == j + 1 or (j == S - 1 and c == 0):
                pass                                        # upper part, by symmetry
            else:
                raise AssertionError("entry outside the block-cyclic-tridiagonal pattern: stage %d, %d" % (j, c))
    if S == 2 and lay.closed:
        raise NotImplementedError("closed laps need at least three nodes")
    return Dg, Lo


def condense_rhs(lay, J, D, r1, r2):
    """Right-hand side of the condensed system in stage order."""
    Jc = sp.csr_matrix(J)
    Ji = Jc[lay.all_ineq]
    b1 = r1 + Ji.T @ (r2[lay.all_ineq] / D[lay.all_ineq])
    return np.concatenate([b1, r2[lay.all_eq]])[lay.perm].reshape(lay.n_stages, lay.nb)


def expand_solution(lay, J, D, r2, z):
    """(dx, dy) of the full augmented system from the stage-ordered solution z."""
    full = np.empty(lay.n + lay.n_stages * lay.neq)
    full[lay.perm] = z.ravel()
    dx = full[:lay.n]
    dy = np.empty(lay.m)
    dy[lay.all_eq] = full[lay.n:]
    Ji = sp.csr_matrix(J)[lay.all_ineq]
    dy[lay.all_ineq] = (Ji @ dx - r2[lay.all_ineq]) / D[lay.all_ineq]
    return dx, dy


ALPHA = (1.0 + 17.0 ** 0.5) / 8.0


def pivoted_gauss_jordan(A):
    """In-place style inverse of a symmetric matrix by Gauss-Jordan sweeps with Bunch-Kaufman pivoting (candidate: the
    largest diagonal entry of the unswept part; 2x2 pivots when the diagonal is weak).  Returns (inverse, number of
    negative pivots, number of vanishing pivots): the statement of kkt_invert in csrc/kkt_kernels.cuh."""
    M = np.array(A, dtype=float)
    n = len(M)
    swept = np.zeros(n, bool)
    neg = zero = 0
    while not swept.all():
        un = np.where(~swept)[0]
        dv = np.abs(M[un, un])
        k = un[np.argmax(dv)]
        dmax = dv.max()
        others = un[un != k]
        lam = np.max(np.abs(M[others, k])) if len(others) else 0.0
        two = False
        if not (dmax >= ALPHA * lam and dmax > 0.0) and lam > 0.0:
            # Bunch-Kaufman: r = row of the largest entry of column k, sigma = largest off-diagonal entry of column r
            r = others[np.argmax(np.abs(M[others, k]))]
            rest = un[un != r]
            sigma = np.max(np.abs(M[rest, r]))
            if dmax * sigma >= ALPHA * lam * lam:
                pass
            elif abs(M[r, r]) >= ALPHA * sigma:
                k = r
            else:
                two, p, q = True, k, r
        if not two:
            d = M[k, k]
            if d == 0.0:
                zero += 1
                d = 1.0e-300
            neg += d < 0.0
            pv = 1.0 / d
            rowk = M[k, :] * pv
            col = M[:, k].copy()
            M -= np.outer(col, rowk)
            M[:, k] = -col * pv
            M[k, :] = rowk
            M[k, k] = pv
            swept[k] = True
        else:
            a, b, b2, c = M[p, p], M[p, q], M[q, p], M[q, q]
            det = a * c - b * b2
            neg += 1 if det < 0.0 else (2 if a < 0.0 else 0)
            Bi = np.array([[c, -b], [-b2, a]]) / det
            rp = Bi[0, 0] * M[p, :] + Bi[0, 1] * M[q, :]
            rq = Bi[1, 0] * M[p, :] + Bi[1, 1] * M[q, :]
            fp, fq = M[:, p].copy(), M[:, q].copy()
            M -= np.outer(fp, rp) + np.outer(fq, rq)
            M[:, p] = -(fp * Bi[0, 0] + fq * Bi[1, 0])
            M[:, q] = -(fp * Bi[0, 1] + fq * Bi[1, 1])
            M[p, :] = rp
            M[q, :] = rq
            M[p, p], M[p, q], M[q, p], M[q, q] = Bi[0, 0], Bi[0, 1], Bi[1, 0], Bi[1, 1]
            swept[p] = swept[q] = True
    return M, int(neg), int(zero)


class BlockCyclicLDL:
    """Backward block elimination (see module docstring).  Dg, Lo are consumed."""

    def __init__(self, Dg, Lo, cyclic):
        S, nb, _ = Dg.shape
        self.S, self.nb, self.cyclic = S, nb, bool(cyclic) and S > 1
        self.Lo = Lo
        self.F = np.zeros((S, nb, nb)) if self.cyclic else None      # F[j] = K[0, j] after fill, j >= 1
        self.Dinv = np.zeros((S, nb, nb))                             # explicit inverses of the pivot blocks
        self.XL = np.zeros((S, nb, nb))                               # D_j^{-1} L_j
        self.XF = np.zeros((S, nb, nb)) if self.cyclic else None      # D_j^{-1} F_j^T
        pos = neg = zero = 0
        self.ok = True
        if self.cyclic:
            self.F[S - 1] = Lo[0]
            self.F[1] = self.F[1] + Lo[1].T
        for j in range(S - 1, -1, -1):
            Dj = 0.5 * (Dg[j] + Dg[j].T)
            Di, q, z = pivoted_gauss_jordan(Dj)
            pos, neg, zero = pos + nb - q - z, neg + q, zero + z
            self.Dinv[j] = Di
            if not np.all(np.isfinite(Di)):
                self.ok = False
            if j == 0:
                break
            if self.cyclic:
                XF = Di @ self.F[j].T
                self.XF[j] = XF
                Dg[0] -= self.F[j] @ XF
            if j >= 2 or not self.cyclic:
                XL = Di @ Lo[j]
                self.XL[j] = XL
                Dg[j - 1] -= Lo[j].T @ XL
                if self.cyclic:
                    self.F[j - 1] -= self.F[j] @ XL
        self.inertia = (pos, neg, zero)

    def solve(self, b):
        """b: [S][nb] -> z with K z = b."""
        S, cyc = self.S, self.cyclic
        b = np.array(b, dtype=float)
        y = np.zeros_like(b)
        for j in range(S - 1, 0, -1):
            y[j] = self.Dinv[j] @ b[j]
            if j >= 2 or not cyc:
                b[j - 1] -= self.Lo[j].T @ y[j]
            if cyc:
                b[0] -= self.F[j] @ y[j]
        z = np.zeros_like(b)
        z[0] = self.Dinv[0] @ b[0]
        for j in range(1, S):
            z[j] = y[j]
            if j >= 2 or not cyc:
                z[j] -= self.XL[j] @ z[j - 1]
            if cyc:
                z[j] -= self.XF[j] @ z[0]
        return z


class PartitionedLDL:
    """The same elimination with the chain cut at P separator stages (parallel in the stages: what the device runs when few
    laps are active).  The interior of every segment (a, b) between consecutive separators is eliminated backwards,
    independently of the other segments, carrying the fill G_j = K[b, j] towards its right separator exactly as the
    sequential algorithm carries F_j = K[0, j] towards stage 0; what is left is a block-(cyclic-)tridiagonal system over
    the separators with couplings of the original shape (only the x columns of the left separator are non-zero), which
    the sequential algorithm factorises.  The inertia is the sum over all pivot blocks, interior and separator ones."""

    def __init__(self, Dg, Lo, cyclic, n_sep):
        S, nb, _ = Dg.shape
        self.S, self.nb, self.cyclic = S, nb, bool(cyclic)
        P = max(3 if cyclic else 1, min(n_sep, S // 2))
        self.sep = np.array(sorted(set((np.arange(P) * S) // P)))
        P = self.P = len(self.sep)
        self.Lo, self.Dinv = Lo, np.zeros((S, nb, nb))
        self.G = np.zeros((S, nb, nb))           # G[j] = K[b, j] for interior stage j of segment (a, b)
        self.right = np.full(S, -1)              # right separator of every interior stage
        Dr = np.array([Dg[s].copy() for s in self.sep])
        Lr = np.zeros((P, nb, nb))
        pos = neg = zero = 0
        self.segments = []
        for k in range(P):
            a = self.sep[k]
            b = self.sep[(k + 1) % P] if (k + 1 < P or self.cyclic) else None      # right separator (None: open end)
            hi = (self.sep[k + 1] if k + 1 < P else S) - 1
            inter = list(range(a + 1, hi + 1))
            self.segments.append((a, b, inter))
            kb = (k + 1) % P
            if not inter:
                if b is not None:
                    Lr[kb] = Lo[b]               # adjacent separators keep their original coupling
                continue
            Gcur = Lo[b].copy() if b is not None else None         # K[b, b-1] (for b = 0: the corner block K[0, S-1])
            for j in reversed(inter):
                Dj = 0.5 * (Dg[j] + Dg[j].T)
                Di, q, z = pivoted_gauss_jordan(Dj)
                pos, neg, zero = pos + nb - q - z, neg + q, zero + z
                self.Dinv[j] = Di
                XL = Di @ Lo[j]
                tgt = Dg[j - 1] if j - 1 != a else Dr[k]
                tgt -= Lo[j].T @ XL
                if Gcur is not None:
                    self.G[j] = Gcur
                    self.right[j] = kb
                    Dr[kb] -= Gcur @ (Di @ Gcur.T)
                    Gcur = -Gcur @ XL
            if b is not None:
                Lr[kb] = Gcur                    # reduced coupling K~[b, a]
        self.red = BlockCyclicLDL(Dr, Lr, self.cyclic and P > 2)
        rp, rn, rz = self.red.inertia
        self.inertia = (pos + rp, neg + rn, zero + rz)
        self.ok = self.red.ok

    def solve(self, rhs):
        S, nb, P = self.S, self.nb, self.P
        b = np.array(rhs, dtype=float)
        y = np.zeros_like(b)
        br = np.array([b[s] for s in self.sep])
        for k, (a, bb, inter) in enumerate(self.segments):
            for j in reversed(inter):
                y[j] = self.Dinv[j] @ b[j]
                if j - 1 != a:
                    b[j - 1] -= self.Lo[j].T @ y[j]
                else:
                    br[k] -= self.Lo[j].T @ y[j]
                if self.right[j] >= 0:
                    br[self.right[j]] -= self.G[j] @ y[j]
        zr = self.red.solve(br)
        z = np.zeros_like(b)
        for k, (a, bb, inter) in enumerate(self.segments):
            z[a] = zr[k]
            prev = zr[k]
            for j in inter:
                w = self.Lo[j] @ prev
                if self.right[j] >= 0:
                    w = w + self.G[j].T @ zr[self.right[j]]
                z[j] = y[j] - self.Dinv[j] @ w
                prev = z[j]
        return z
