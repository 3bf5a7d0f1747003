"""CPU oracle for the arrowhead hot path of PiecewiseOrthogonalPolynomials.jl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker.

Parity pin status: the reference (Julia) cannot run in this container (no
julia binary, dependencies un-vendored and un-pinned).  The
oracle is therefore pinned against (i) the only known-answer value the
reference's own tests hold for this path, 1.1952730862177243
(/root/reference/test/test_arrowhead.jl:134), (ii) the structural goldens
of test_arrowhead.jl (bandwidths (13,9), block-bandwidths (3,2), (7,7)/(1,1)),
(iii) "same operation on the dense Matrix(A)" equivalences which is what every
testset of the reference checks (test_arrowhead.jl:49-53,64-68,90,99,143-145),
and (iv) Gauss quadrature of the basis functions for the assembly closed forms.
See tests/test_oracle_golden.py.

Three levels:
  L0  dense NumPy linear algebra (ground truth for small n),
  L1  literal restatement of the reference's block algorithms
      (src/arrowhead.jl), loops kept in the reference's order, 0-based,
  L2  packed / element-fastest vectorised form (the layout the C ABI uses),
      usable at medium sizes; the C restatement oracle/pop_oracle.c is the
      same arithmetic for full sizes.

All file:line citations are into /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# ---------------------------------------------------------------------------
# L1 container: the reference's own field structure (src/arrowhead.jl:31-38)
# ---------------------------------------------------------------------------


@dataclass
class BlockArrowhead:
    """A (nh x nh), B[b] (nh x m), C[b] (m x nh), D[k] (q x q), all dense ndarrays.

    Mirrors `struct BBBArrowheadMatrix{A,B,C,D}` (src/arrowhead.jl:31-38).  The
    global ordering is blocks (nh, m, m, ...) (src/arrowhead.jl:57-62): block 1
    are the hats, block K>=2 entry k is bubble K-1 of element k.
    """

    A: np.ndarray
    B: List[np.ndarray]
    C: List[np.ndarray]
    D: List[np.ndarray]

    @property
    def nh(self):
        return self.A.shape[0]

    @property
    def m(self):
        return len(self.D)

    @property
    def q(self):
        return self.D[0].shape[0]

    @property
    def n(self):
        return self.nh + self.m * self.q

    def copy(self):
        return BlockArrowhead(self.A.copy(), [b.copy() for b in self.B], [c.copy() for c in self.C], [d.copy() for d in self.D])

    def getindex(self, K, k, J, j):
        """0-based block indices. src/arrowhead.jl:70-82."""
        if J == 0 and K == 0:
            return self.A[k, j]
        if K == 0:
            return self.B[J - 1][k, j] if J - 1 < len(self.B) else 0.0
        if J == 0:
            return self.C[K - 1][k, j] if K - 1 < len(self.C) else 0.0
        if k != j:
            return 0.0
        return self.D[k][K - 1, J - 1]

    def to_dense(self):
        nh, m, q = self.nh, self.m, self.q
        n = self.n
        S = np.zeros((n, n))
        S[:nh, :nh] = self.A
        for b, Bb in enumerate(self.B):
            if b < q:
                S[:nh, nh + b * m: nh + (b + 1) * m] = Bb
        for b, Cb in enumerate(self.C):
            if b < q:
                S[nh + b * m: nh + (b + 1) * m, :nh] = Cb
        for k in range(m):
            S[nh + k::m, nh + k::m] = self.D[k]
        return S

    def transpose(self):
        """src/arrowhead.jl:66-68."""
        return BlockArrowhead(self.A.T.copy(), [c.T.copy() for c in self.C], [b.T.copy() for b in self.B], [d.T.copy() for d in self.D])


def symmetric_from_upper(P: BlockArrowhead) -> BlockArrowhead:
    """`Symmetric(P)` field views: A->symmetric(A), C = adjoint.(B), D->symmetric.(D)
    (src/arrowhead.jl:154-167)."""
    symm = lambda X: np.triu(X) + np.triu(X, 1).T
    return BlockArrowhead(symm(P.A), [b.copy() for b in P.B], [b.T.copy() for b in P.B], [symm(d) for d in P.D])


def tupleop(op, A: Sequence[np.ndarray], B: Sequence[np.ndarray]):
    """Ragged block-tuple combination, src/arrowhead.jl:392-398."""
    out = []
    for i in range(max(len(A), len(B))):
        if i < len(A) and i < len(B):
            out.append(op(A[i], B[i]))
        elif i < len(A):
            out.append(A[i].copy())
        else:
            out.append(op(np.zeros_like(B[i]), B[i]))
    return out


def block_add(X: BlockArrowhead, Y: BlockArrowhead, sign=+1.0) -> BlockArrowhead:
    """`+`/`-` of arrowheads, src/arrowhead.jl:406-409."""
    op = (lambda a, b: a + b) if sign > 0 else (lambda a, b: a - b)
    return BlockArrowhead(op(X.A, Y.A), tupleop(op, X.B, Y.B), tupleop(op, X.C, Y.C), [op(a, b) for a, b in zip(X.D, Y.D)])


def block_scale(c: float, X: BlockArrowhead) -> BlockArrowhead:
    """`c*A`, src/arrowhead.jl:411-417."""
    return BlockArrowhead(c * X.A, [c * b for b in X.B], [c * b for b in X.C], [c * d for d in X.D])


# ---------------------------------------------------------------------------
# L0 dense ground truth
# ---------------------------------------------------------------------------


def dense_reversecholesky(S: np.ndarray) -> np.ndarray:
    """Upper U with S = U U^T (what MatrixFactorizations.reversecholesky(S).U returns):
    flip both axes, ordinary Cholesky, flip back."""
    J = S[::-1, ::-1]
    L = np.linalg.cholesky(J)
    return L[::-1, ::-1]


# ---------------------------------------------------------------------------
# L1 literal restatement of src/arrowhead.jl
# ---------------------------------------------------------------------------


def ref_mul_vec(alpha, A: BlockArrowhead, x, beta, y):
    """y <- alpha*A*x + beta*y ; src/arrowhead.jl:191-212."""
    nh, m = A.nh, A.m
    y *= beta                                                   # :197
    y[:nh] += alpha * (A.A @ x[:nh])                            # :199
    for k in range(len(A.B)):                                   # :200-202
        y[:nh] += alpha * (A.B[k] @ x[nh + k * m: nh + (k + 1) * m])
    for k in range(len(A.C)):                                   # :203-205
        y[nh + k * m: nh + (k + 1) * m] += alpha * (A.C[k] @ x[:nh])
    for k in range(m):                                          # :207-210
        y[nh + k::m] += alpha * (A.D[k] @ x[nh + k::m])
    return y


def ref_mul_mat_left(alpha, A: BlockArrowhead, X, beta, Y):
    """Y <- alpha*A*X + beta*Y ; src/arrowhead.jl:214-235."""
    nh, m, q = A.nh, A.m, A.q
    Y *= beta
    Y[:nh] += alpha * (A.A @ X[:nh])
    for k in range(min(len(A.B), q)):                           # :223
        Y[:nh] += alpha * (A.B[k] @ X[nh + k * m: nh + (k + 1) * m])
    for k in range(min(len(A.C), q)):                           # :226
        Y[nh + k * m: nh + (k + 1) * m] += alpha * (A.C[k] @ X[:nh])
    for k in range(m):                                          # :231-233
        Y[nh + k::m] += alpha * (A.D[k] @ X[nh + k::m])
    return Y


def ref_mul_mat_right(alpha, X, A: BlockArrowhead, beta, Y):
    """Y <- alpha*X*A + beta*Y ; src/arrowhead.jl:237-258."""
    nh, m = A.nh, A.m
    Y *= beta
    Y[:, :nh] += alpha * (X[:, :nh] @ A.A)                      # :245
    for k in range(len(A.C)):                                   # :246-248
        Y[:, :nh] += alpha * (X[:, nh + k * m: nh + (k + 1) * m] @ A.C[k])
    for k in range(len(A.B)):                                   # :249-251
        Y[:, nh + k * m: nh + (k + 1) * m] += alpha * (X[:, :nh] @ A.B[k])
    for k in range(m):                                          # :253-256
        Y[:, nh + k::m] += alpha * (X[:, nh + k::m] @ A.D[k])
    return Y


def _reversecholesky_upper_inplace(S: np.ndarray, bw: Optional[int] = None):
    """In-place reverse Cholesky of a symmetric matrix given by its upper triangle
    (what MatrixFactorizations' `reversecholesky!(Symmetric(B))` does for the banded
    D blocks and the bidiagonal hat block, call sites src/arrowhead.jl:278,289).
    Third-party arithmetic restated from its definition S = U U^T, U upper:
      for j = n-1..0:  U[j,j] = sqrt(S[j,j] - sum_{c>j} U[j,c]^2)
                       U[i,j] = (S[i,j] - sum_{c>j} U[i,c] U[j,c]) / U[j,j],  i<j
    Raises on a non-positive pivot (the reference throws PosDefException)."""
    n = S.shape[0]
    u = n - 1 if bw is None else bw
    for j in range(n - 1, -1, -1):
        hi = min(n, j + u + 1)
        piv = S[j, j] - np.dot(S[j, j + 1:hi], S[j, j + 1:hi])
        if not piv > 0:
            raise np.linalg.LinAlgError(f"reverse Cholesky: non-positive pivot at {j}")
        S[j, j] = math.sqrt(piv)
        for i in range(max(0, j - u), j):
            hi_i = min(n, i + u + 1)
            S[i, j] = (S[i, j] - np.dot(S[i, j + 1:hi_i], S[j, j + 1:hi_i])) / S[j, j]
    return S


def _bandwidths(X: np.ndarray) -> Tuple[int, int]:
    """(l,u) of a dense matrix (BandedMatrices.bandwidths of the stored structure is
    what the reference branches on; for dense test data we use the numerical one)."""
    r, c = np.nonzero(X)
    if r.size == 0:
        return (-1, -1) if False else (0, 0)
    return int(max(0, (r - c).max())), int(max(0, (c - r).max()))


def ref_reversecholesky_layout(A: BlockArrowhead, lower_B: Optional[bool] = None):
    """In-place S = U U^T on the upper data of A. src/arrowhead.jl:276-292.
    `lower_B` selects `_reverse_chol_lower_B!` (C1 / Neumann, B bandwidths (1,0),
    nh == m+1) or `_reverse_chol_upper_B!` (Dirichlet, (0,1), nh == m-1); the
    reference decides from bandwidths(A.B[1]) (:282); we decide from nh."""
    for k in range(A.m):                                        # :277-279
        _reversecholesky_upper_inplace(A.D[k])
    if len(A.B) > 0:                                            # :281
        if lower_B is None:
            lower_B = A.nh == A.m + 1
        if lower_B:
            _ref_reverse_chol_lower_B(A)
        else:
            _ref_reverse_chol_upper_B(A)
    _reversecholesky_upper_inplace(A.A, 1)                      # :289
    return A


def _ref_reverse_chol_lower_B(A: BlockArrowhead):
    """src/arrowhead.jl:297-344 (1-based loops rewritten 0-based)."""
    nh = A.nh
    m = A.m
    nB = len(A.B)
    assert nh == m + 1                                          # :305
    for b in range(nB - 1, -1, -1):                             # :308
        for k in range(m - 1, -1, -1):
            for bt in range(b + 1, nB):                         # :309
                Akj = A.D[k][b, bt]
                if Akj != 0:
                    for i in (k, k + 1):                        # :313-315
                        A.B[b][i, k] -= A.B[bt][i, k] * Akj
            AkkInv = 1.0 / A.D[k][b, b]                         # :319
            for i in (k, k + 1):
                A.B[b][i, k] *= AkkInv
    n = nh
    for k in range(n - 1, -1, -1):                              # :326 (k is 0-based hat index)
        for bt in range(nB):
            if k != 0:                                          # :328-334
                j = k - 1
                Akj = A.B[bt][k, j]
                for i in (j, j + 1):
                    A.A[i, k] -= A.B[bt][i, j] * Akj
            if k != n - 1:                                      # :336-341
                j = k
                Akj = A.B[bt][k, j]
                A.A[j, k] -= A.B[bt][j, j] * Akj


def _ref_reverse_chol_upper_B(A: BlockArrowhead):
    """src/arrowhead.jl:346-389 (0-based)."""
    nh = A.nh
    m = A.m
    q = A.q
    assert nh == m - 1                                          # :356
    nb = min(len(A.B), q)                                       # :359 (N-1 == q)
    n = nh
    for b in range(nb - 1, -1, -1):
        for k in range(m - 1, -1, -1):
            rows = range(max(0, k - 1), min(k, n - 1) + 1)      # 1-based max(1,k-1):min(k,n)
            for bt in range(b + 1, nb):
                Akj = A.D[k][b, bt]
                if Akj != 0:
                    for i in rows:
                        A.B[b][i, k] -= A.B[bt][i, k] * Akj
            AkkInv = 1.0 / A.D[k][b, b]
            for i in rows:
                A.B[b][i, k] *= AkkInv
    for bt in range(len(A.B)):                                  # :378
        for k in range(n - 1, -1, -1):                          # :379
            Akj = A.B[bt][k, k + 1]                             # :380
            A.A[k, k] -= Akj ** 2
            Akj = A.B[bt][k, k]                                 # :383
            for i in range(max(0, k - 1), k + 1):
                A.A[i, k] -= Akj * A.B[bt][i, k]


def _tri_solve_dense(T: np.ndarray, b: np.ndarray, upper: bool, unit: bool):
    """Banded triangular solves are third-party (BandedMatrices/ArrayLayouts, call
    sites src/arrowhead.jl:440,453,472,481); mathematically plain substitution."""
    n = T.shape[0]
    rng = range(n - 1, -1, -1) if upper else range(n)
    for i in rng:
        if upper:
            s = b[i] - T[i, i + 1:] @ b[i + 1:]
        else:
            s = b[i] - T[i, :i] @ b[:i]
        b[i] = s if unit else s / T[i, i]
    return b


def ref_ldiv_upper(P: BlockArrowhead, dest: np.ndarray, unit=False):
    """dest <- UpperTriangular(P) \\ dest ; src/arrowhead.jl:425-457."""
    nh, m, q = P.nh, P.m, P.q
    for k in range(m):                                          # :439-441
        v = dest[nh + k::m].copy()
        _tri_solve_dense(np.triu(P.D[k]), v, True, unit)
        dest[nh + k::m] = v
    for K in range(min(q, len(P.B))):                           # :449-451
        dest[:nh] -= P.B[K] @ dest[nh + K * m: nh + (K + 1) * m]
    h = dest[:nh].copy()                                        # :453
    _tri_solve_dense(np.triu(P.A), h, True, unit)
    dest[:nh] = h
    return dest


def ref_ldiv_lower(P: BlockArrowhead, dest: np.ndarray, unit=False):
    """dest <- LowerTriangular(P) \\ dest ; src/arrowhead.jl:458-486."""
    nh, m, q = P.nh, P.m, P.q
    h = dest[:nh].copy()                                        # :472
    _tri_solve_dense(np.tril(P.A), h, False, unit)
    dest[:nh] = h
    for K in range(min(q, len(P.C))):                           # :475-477
        dest[nh + K * m: nh + (K + 1) * m] -= P.C[K] @ dest[:nh]
    for k in range(m):                                          # :480-482
        v = dest[nh + k::m].copy()
        _tri_solve_dense(np.tril(P.D[k]), v, False, unit)
        dest[nh + k::m] = v
    return dest


def ref_revchol_solve(U: BlockArrowhead, x: np.ndarray):
    """`ReverseCholesky \\ x` = U' \\ (U \\ x) (MatrixFactorizations; call sites
    test/test_arrowhead.jl:130, examples/adi_poisson.jl:54-55).  U holds the factor in
    its upper data; the adjoint is a lower arrowhead with C = B' (src/arrowhead.jl:66-68)."""
    y = np.array(x, dtype=float, copy=True)
    cols = y.reshape(y.shape[0], -1)
    Ut = U.transpose()
    for c in range(cols.shape[1]):
        v = cols[:, c].copy()
        ref_ldiv_upper(U, v)
        ref_ldiv_lower(Ut, v)
        cols[:, c] = v
    return y


def block_bandwidths(A: BlockArrowhead) -> Tuple[int, int]:
    """src/arrowhead.jl:52-55 with D bandwidths taken numerically."""
    l, u = _bandwidths(A.D[0])
    return max(l, len(A.C)), max(u, len(A.B))


def banded_bandwidths(A: BlockArrowhead) -> Tuple[int, int]:
    """src/arrowhead.jl:526-528: size(A.A,k) + sizes of all but the last C (B) block +
    bandwidth of the last."""
    def bbb(dim, first, blocks, which):
        if len(blocks) == 0:
            return _bandwidths(first)[which]
        tot = first.shape[dim]
        for blk in blocks[:-1]:
            tot += blk.shape[dim]
        return tot + _bandwidths(blocks[-1])[which]
    return bbb(0, A.A, A.C, 0), bbb(1, A.A, A.B, 1)


# ---------------------------------------------------------------------------
# Packed, element-fastest form (the C-ABI layout, include/pop_arrowhead.h)
# ---------------------------------------------------------------------------


@dataclass
class Packed:
    """Packed arrowhead.  Ad[nh], Au[nh-1], Al[nh-1]; B[nB,3,m], C[nC,3,m] with slot
    t<->hat k+t-1 (entries whose hat is outside [0,nh) are zero);
    D[lD+uD+1, q, m] with D[lD+d, j, k] = D_k[j, j+d] (zero when j+d outside [0,q))."""
    m: int
    nh: int
    q: int
    lD: int
    uD: int
    Ad: np.ndarray
    Au: np.ndarray
    Al: np.ndarray
    B: np.ndarray
    C: np.ndarray
    D: np.ndarray

    @property
    def n(self):
        return self.nh + self.m * self.q

    @property
    def nB(self):
        return self.B.shape[0]

    @property
    def nC(self):
        return self.C.shape[0]

    def copy(self):
        return Packed(self.m, self.nh, self.q, self.lD, self.uD, self.Ad.copy(), self.Au.copy(), self.Al.copy(),
                      self.B.copy(), self.C.copy(), self.D.copy())


def pack(A: BlockArrowhead, lD: Optional[int] = None, uD: Optional[int] = None) -> Packed:
    nh, m, q = A.nh, A.m, A.q
    if lD is None or uD is None:
        bw = [_bandwidths(d) for d in A.D]
        lD = max(b[0] for b in bw) if lD is None else lD
        uD = max(b[1] for b in bw) if uD is None else uD
    Ad = np.diag(A.A).copy()
    Au = np.diag(A.A, 1).copy()
    Al = np.diag(A.A, -1).copy()

    def slots(blocks, transpose):
        out = np.zeros((len(blocks), 3, m))
        for b, blk in enumerate(blocks):
            Bm = blk.T if transpose else blk     # -> (nh, m)
            for k in range(m):
                for t in range(3):
                    i = k + t - 1
                    if 0 <= i < nh:
                        out[b, t, k] = Bm[i, k]
            chk = np.zeros_like(Bm)
            for k in range(m):
                for t in range(3):
                    i = k + t - 1
                    if 0 <= i < nh:
                        chk[i, k] = Bm[i, k]
            assert np.array_equal(chk, Bm), "B/C block outside bandwidths (1,1) (src/arrowhead.jl:100-111)"
        return out

    B = slots(A.B, False)
    C = slots(A.C, True)
    D = np.zeros((lD + uD + 1, q, m))
    for k in range(m):
        for d in range(-lD, uD + 1):
            diag = np.diag(A.D[k], d)
            if d >= 0:
                D[lD + d, :q - d, k] = diag
            else:
                D[lD + d, -d:, k] = diag
    return Packed(m, nh, q, lD, uD, Ad, Au, Al, B, C, D)


def unpack(P: Packed) -> BlockArrowhead:
    nh, m, q = P.nh, P.m, P.q
    A = np.diag(P.Ad)
    if nh > 1:
        A += np.diag(P.Au, 1) + np.diag(P.Al, -1)

    def blocks(S, transpose):
        out = []
        for b in range(S.shape[0]):
            Bm = np.zeros((nh, m))
            for k in range(m):
                for t in range(3):
                    i = k + t - 1
                    if 0 <= i < nh:
                        Bm[i, k] = S[b, t, k]
            out.append(Bm.T.copy() if transpose else Bm)
        return out

    Ds = []
    for k in range(m):
        Dk = np.zeros((q, q))
        for d in range(-P.lD, P.uD + 1):
            if abs(d) < q:
                if d >= 0:
                    Dk += np.diag(P.D[P.lD + d, :q - d, k], d)
                else:
                    Dk += np.diag(P.D[P.lD + d, -d:, k], d)
        Ds.append(Dk)
    return BlockArrowhead(A, blocks(P.B, False), blocks(P.C, True), Ds)


def packed_symmetrize(P: Packed) -> Packed:
    """Full arrowhead equal to Symmetric(P) (upper data) (src/arrowhead.jl:154-167)."""
    q, m, u = P.q, P.m, P.uD
    D = np.zeros((2 * u + 1, q, m))
    for d in range(0, u + 1):
        D[u + d] = P.D[P.lD + d]
        if d > 0:
            D[u - d, d:, :] = P.D[P.lD + d, :q - d, :]
    return Packed(P.m, P.nh, P.q, u, u, P.Ad.copy(), P.Au.copy(), P.Au.copy(), P.B.copy(), P.B.copy(), D)


def packed_mul_left(alpha, P: Packed, X, beta, Y):
    """Y <- alpha*P*X + beta*Y, X (n x r). Index form of SURVEY appendix A / src/arrowhead.jl:214-235."""
    nh, m, q = P.nh, P.m, P.q
    X2 = X.reshape(X.shape[0], -1)
    R = np.zeros_like(X2)
    H = X2[:nh]
    R[:nh] = P.Ad[:, None] * H
    if nh > 1:
        R[:nh - 1] += P.Au[:, None] * H[1:]
        R[1:nh] += P.Al[:, None] * H[:-1]
    ks = np.arange(m)
    for b in range(min(P.nB, q)):
        Zb = X2[nh + b * m: nh + (b + 1) * m]
        for t in range(3):
            i = ks + t - 1
            ok = (i >= 0) & (i < nh)
            np.add.at(R, (i[ok],), P.B[b, t, ok][:, None] * Zb[ok])
    for b in range(min(P.nC, q)):
        Rb = R[nh + b * m: nh + (b + 1) * m]
        for t in range(3):
            i = ks + t - 1
            ok = (i >= 0) & (i < nh)
            Rb[ok] += P.C[b, t, ok][:, None] * H[i[ok]]
    for j in range(q):
        Rj = R[nh + j * m: nh + (j + 1) * m]
        for d in range(-P.lD, P.uD + 1):
            if 0 <= j + d < q:
                Rj += P.D[P.lD + d, j][:, None] * X2[nh + (j + d) * m: nh + (j + d + 1) * m]
    Y2 = Y.reshape(Y.shape[0], -1)
    Y2[:] = alpha * R + beta * Y2
    return Y


def packed_revchol(P: Packed) -> Packed:
    """Reverse Cholesky on the upper data of P (lower data ignored) -> packed upper factor
    U (lD=0, nC=0).  Index form of src/arrowhead.jl:276-389."""
    nh, m, q, u = P.nh, P.m, P.q, P.uD
    Dd = P.D[P.lD:].copy()                      # Dd[d][j][k] = D_k[j, j+d]
    for j in range(q - 1, -1, -1):
        s = Dd[0, j].copy()
        for d in range(1, u + 1):
            if j + d < q:
                s -= Dd[d, j] ** 2
        if not np.all(s > 0):
            raise np.linalg.LinAlgError("non-positive pivot in D block")
        Dd[0, j] = np.sqrt(s)
        for d in range(1, u + 1):
            i = j - d
            if i < 0:
                continue
            acc = Dd[d, i].copy()
            for c in range(j + 1, min(i + u, q - 1) + 1):
                acc -= Dd[c - i, i] * Dd[c - j, j]
            Dd[d, i] = acc / Dd[0, j]
    B = P.B.copy()
    nb = min(P.nB, q)
    for b in range(nb - 1, -1, -1):
        for bt in range(b + 1, nb):
            if bt - b <= u:
                B[b] -= B[bt] * Dd[bt - b, b][None, :]
        B[b] *= (1.0 / Dd[0, b])[None, :]
    Ad, Au = P.Ad.copy(), P.Au.copy()
    ks = np.arange(m)
    # only blocks inside the truncation take part (the reference's loop over all of A.B at :378 is
    # the same thing whenever length(A.B) <= N-1, which its own callers guarantee)
    for b in range(nb):
        for t in range(3):
            i = ks + t - 1
            ok = (i >= 0) & (i < nh)
            np.subtract.at(Ad, i[ok], B[b, t, ok] ** 2)
        for t in range(2):      # pairs (t, t+1) couple hats i, i+1
            i = ks + t - 1
            ok = (i >= 0) & (i + 1 < nh)
            np.subtract.at(Au, i[ok], B[b, t, ok] * B[b, t + 1, ok])
    for i in range(nh - 1, -1, -1):
        s = Ad[i] - (Au[i] ** 2 if i < nh - 1 else 0.0)
        if not s > 0:
            raise np.linalg.LinAlgError("non-positive pivot in hat block")
        Ad[i] = np.sqrt(s)              # np.sqrt keeps the dtype (extended-precision runs of the oracle)
        if i > 0:
            Au[i - 1] = Au[i - 1] / Ad[i]
    return Packed(m, nh, q, 0, u, Ad, Au, np.zeros_like(Au), B, np.zeros((0, 3, m)), Dd)


def packed_trsm(U: Packed, X, upper=True, unit=False):
    """In place X <- T^{-1} X.  upper=True: T = UpperTriangular(U) (data Ad,Au,B,D[d>=0]);
    upper=False: T = LowerTriangular(U) (data Ad,Al,C,D[d<=0]).  src/arrowhead.jl:425-486."""
    nh, m, q = U.nh, U.m, U.q
    X2 = X.reshape(X.shape[0], -1)
    H = X2[:nh]
    Z = lambda j: X2[nh + j * m: nh + (j + 1) * m]
    ks = np.arange(m)
    dg = U.D[U.lD]
    if upper:
        for j in range(q - 1, -1, -1):
            v = Z(j).copy()
            for d in range(1, U.uD + 1):
                if j + d < q:
                    v -= U.D[U.lD + d, j][:, None] * Z(j + d)
            Z(j)[:] = v if unit else v / dg[j][:, None]
        for b in range(min(U.nB, q)):
            for t in range(3):
                i = ks + t - 1
                ok = (i >= 0) & (i < nh)
                np.subtract.at(H, (i[ok],), U.B[b, t, ok][:, None] * Z(b)[ok])
        for i in range(nh - 1, -1, -1):
            v = H[i].copy()
            if i + 1 < nh:
                v -= U.Au[i] * H[i + 1]
            H[i] = v if unit else v / U.Ad[i]
    else:
        for i in range(nh):
            v = H[i].copy()
            if i > 0:
                v -= U.Al[i - 1] * H[i - 1]
            H[i] = v if unit else v / U.Ad[i]
        for b in range(min(U.nC, q)):
            for t in range(3):
                i = ks + t - 1
                ok = (i >= 0) & (i < nh)
                Z(b)[ok] -= U.C[b, t, ok][:, None] * H[i[ok]]
        for j in range(q):
            v = Z(j).copy()
            for d in range(1, U.lD + 1):
                if j - d >= 0:
                    v -= U.D[U.lD - d, j][:, None] * Z(j - d)
            Z(j)[:] = v if unit else v / dg[j][:, None]
    return X


def packed_astype(P: Packed, dtype) -> Packed:
    """Same arrowhead in another floating type (np.longdouble gives the noise-floor reference)."""
    f = lambda a: np.asarray(a).astype(dtype)
    return Packed(P.m, P.nh, P.q, P.lD, P.uD, f(P.Ad), f(P.Au), f(P.Al), f(P.B), f(P.C), f(P.D))


def packed_transpose(P: Packed) -> Packed:
    q, m = P.q, P.m
    D = np.zeros((P.lD + P.uD + 1, q, m), dtype=P.D.dtype)
    # new lD = old uD ; D'[j, j+d] = D[j+d, j] -> old band -d at row j+d
    for d in range(-P.uD, P.lD + 1):
        src = P.D[P.lD - d]          # old band -d : D[r, r-d]
        if d >= 0:
            D[P.uD + d, :q - d, :] = src[d:, :] if d > 0 else src
        else:
            D[P.uD + d, -d:, :] = src[:q + d, :]
    return Packed(P.m, P.nh, P.q, P.uD, P.lD, P.Ad.copy(), P.Al.copy(), P.Au.copy(), P.C.copy(), P.B.copy(), D)


def packed_solve(U: Packed, X):
    """X <- U^{-T} U^{-1} X (the ReverseCholesky solve)."""
    packed_trsm(U, X, upper=True)
    packed_trsm(packed_transpose(U), X, upper=False)
    return X


def packed_axpby(alpha, X: Packed, beta, Y: Packed) -> Packed:
    """alpha*X + beta*Y with ragged B/C and different D bandwidths (src/arrowhead.jl:392-417)."""
    assert (X.m, X.nh, X.q) == (Y.m, Y.nh, Y.q)
    lD, uD = max(X.lD, Y.lD), max(X.uD, Y.uD)
    D = np.zeros((lD + uD + 1, X.q, X.m))
    D[lD - X.lD: lD + X.uD + 1] += alpha * X.D
    D[lD - Y.lD: lD + Y.uD + 1] += beta * Y.D

    def rag(a, b):
        nb = max(a.shape[0], b.shape[0])
        out = np.zeros((nb, 3, X.m))
        out[:a.shape[0]] += alpha * a
        out[:b.shape[0]] += beta * b
        return out
    return Packed(X.m, X.nh, X.q, lD, uD, alpha * X.Ad + beta * Y.Ad, alpha * X.Au + beta * Y.Au, alpha * X.Al + beta * Y.Al,
                  rag(X.B, Y.B), rag(X.C, Y.C), D)


# ---------------------------------------------------------------------------
# Assembly closed forms (uniform mesh), upper data only as in the reference
# ---------------------------------------------------------------------------

C1, DIRICHLET = 0, 1


def _assemble_mass_uniform(basis: int, m: int, N: int, h: float) -> Packed:
    """The reference's own expressions on a range (src/continuouspolynomial.jl:274-290, src/dirichlet.jl:180-196)."""
    q = N - 1
    nh = m + 1 if basis == C1 else m - 1
    if basis == C1:
        Ad = np.full(nh, 2 * h / 3)
        Ad[0] = Ad[-1] = h / 3
    else:
        Ad = np.full(nh, 2 * h / 3)
    Au = np.full(max(nh - 1, 0), h / 6)
    B = np.zeros((2, 3, m))
    ks = np.arange(m)
    # left hat of element k: C1 -> hat k (slot 1), Dirichlet -> hat k-1 (slot 0)
    tl, tr = (1, 2) if basis == C1 else (0, 1)
    for t, v1, v2 in ((tl, h / 6, -h / 30), (tr, h / 6, h / 30)):
        i = ks + t - 1
        ok = (i >= 0) & (i < nh)
        B[0, t, ok] = v1
        B[1, t, ok] = v2
    D = np.zeros((3, q, m))
    i1 = np.arange(1, q + 1, dtype=float)
    D[0] = ((h / 2) * 4.0 / ((2 * i1 - 1) * (2 * i1 + 1) * (2 * i1 + 3)))[:, None]
    if q > 2:
        i2 = i1[:q - 2]
        D[2, :q - 2] = ((h / 2) * (-2.0) / ((2 * i2 + 1) * (2 * i2 + 3) * (2 * i2 + 5)))[:, None]
    return Packed(m, nh, q, 0, 2, Ad, Au, np.zeros_like(Au), B, np.zeros((0, 3, m)), D)


def _assemble_weaklaplacian_uniform(basis: int, m: int, N: int, h: float) -> Packed:
    """src/continuouspolynomial.jl:348-357 (C1), src/dirichlet.jl:169-178 (Dirichlet)."""
    q = N - 1
    nh = m + 1 if basis == C1 else m - 1
    si = 1.0 / h
    Ad = np.full(nh, -2 * si)
    if basis == C1:
        Ad[0] = Ad[-1] = -si
    Au = np.full(max(nh - 1, 0), si)
    D = np.zeros((1, q, m))
    i1 = np.arange(1, q + 1, dtype=float)
    D[0] = (-4.0 / (h * (2 * i1 + 1)))[:, None]
    return Packed(m, nh, q, 0, 0, Ad, Au, np.zeros_like(Au), np.zeros((0, 3, m)), np.zeros((0, 3, m)), D)


def _element_steps(h, m: int) -> np.ndarray:
    """Element lengths: a scalar step (uniform mesh, `step(r)`) or the m values `diff(points)`."""
    hk = np.asarray(h, dtype=np.float64)
    if hk.ndim == 0:
        hk = np.full(m, float(hk))
    if hk.shape != (m,) or np.any(hk <= 0):
        raise ValueError("need m positive element lengths")
    return hk


def assemble_mass(basis: int, m: int, N: int, h) -> Packed:
    """grammatrix(C)[KR,KR] upper data, KR = Block.(oneto(N)).
    C1: src/continuouspolynomial.jl:274-290 ; Dirichlet: src/dirichlet.jl:180-196.
    `h` scalar: the closed forms of a uniform mesh, as in the reference.  `h` array (element lengths of a
    non-uniform mesh): the same element integrals with h_k per element, i.e. what the reference's generic
    `L' * grammatrix(P) * L` (src/continuouspolynomial.jl:259-265,293-297) evaluates to."""
    q = N - 1
    nh = m + 1 if basis == C1 else m - 1
    if np.ndim(h) == 0:
        return _assemble_mass_uniform(basis, m, N, float(h))
    hk = _element_steps(h, m)
    # hat i of the C1 basis sits on node i (elements i-1 and i); the Dirichlet basis drops nodes 0 and m
    node = np.zeros(m + 1)
    node[:-1] += hk / 3
    node[1:] += hk / 3
    if basis == C1:
        Ad = node.copy()
        Au = hk / 6                       # hats k, k+1 share element k
    else:
        Ad = node[1:-1].copy()
        Au = hk[1:-1] / 6                 # hats i, i+1 (nodes i+1, i+2) share element i+1
    B = np.zeros((2, 3, m))
    ks = np.arange(m)
    # left hat of element k: C1 -> hat k (slot 1), Dirichlet -> hat k-1 (slot 0)
    tl, tr = (1, 2) if basis == C1 else (0, 1)
    for t, v1, v2 in ((tl, hk / 6, -hk / 30), (tr, hk / 6, hk / 30)):
        i = ks + t - 1
        ok = (i >= 0) & (i < nh)
        B[0, t, ok] = v1[ok]
        B[1, t, ok] = v2[ok]
    D = np.zeros((3, q, m))
    i1 = np.arange(1, q + 1, dtype=float)
    D[0] = (hk / 2)[None, :] * 4.0 / ((2 * i1 - 1) * (2 * i1 + 1) * (2 * i1 + 3))[:, None]
    if q > 2:
        i2 = i1[:q - 2]
        D[2, :q - 2] = (hk / 2)[None, :] * (-2.0) / ((2 * i2 + 1) * (2 * i2 + 3) * (2 * i2 + 5))[:, None]
    return Packed(m, nh, q, 0, 2, Ad, Au, np.zeros_like(Au), B, np.zeros((0, 3, m)), D)


def assemble_weaklaplacian(basis: int, m: int, N: int, h, a=None) -> Packed:
    """weaklaplacian(C)[KR,KR] upper data (negative semi-definite).
    C1: src/continuouspolynomial.jl:348-357 ; Dirichlet: src/dirichlet.jl:169-178 (uniform meshes; the reference
    defines it for ranges only).  With an array of element lengths: -(diff C)' G (diff C) with the per-element
    scalings of `diff` on a non-uniform mesh (src/continuouspolynomial.jl:318-325)."""
    q = N - 1
    nh = m + 1 if basis == C1 else m - 1
    if np.ndim(h) == 0 and a is None:
        return _assemble_weaklaplacian_uniform(basis, m, N, float(h))
    hk = _element_steps(h, m)
    # a: piecewise-constant diffusion coefficient, one value per element: -(D*C)' * (a .* (D*C)) of examples/heat.jl:77-79
    si = (1.0 if a is None else np.asarray(a, dtype=np.float64)) / hk
    node = np.zeros(m + 1)
    node[:-1] -= si
    node[1:] -= si
    if basis == C1:
        Ad = node.copy()
        Au = si.copy()
    else:
        Ad = node[1:-1].copy()
        Au = si[1:-1].copy()
    D = np.zeros((1, q, m))
    i1 = np.arange(1, q + 1, dtype=float)
    D[0] = -4.0 * si[None, :] / (2 * i1 + 1)[:, None]
    return Packed(m, nh, q, 0, 0, Ad, Au, np.zeros_like(Au), np.zeros((0, 3, m)), np.zeros((0, 3, m)), D)


# -- independent check of the closed forms: Gauss quadrature of the basis --------


def _legendre_table(x, nmax):
    P = np.zeros((nmax + 1, x.size))
    P[0] = 1
    if nmax >= 1:
        P[1] = x
    for n in range(1, nmax):
        P[n + 1] = ((2 * n + 1) * x * P[n] - n * P[n - 1]) / (n + 1)
    return P


def bubble_values(xi, nmax):
    """W_n(xi) = C_n^{(-1/2)}(xi) = (P_{n-2} - P_n)/(2n-1), n = 2..nmax: the bubble of
    block K = n (`bubblebasis`, src/continuouspolynomial.jl:37,47) and its xi-derivative -P_{n-1}."""
    P = _legendre_table(np.atleast_1d(np.asarray(xi, dtype=float)), nmax)
    W = np.zeros((nmax + 1, P.shape[1]))
    dW = np.zeros_like(W)
    for n in range(2, nmax + 1):
        W[n] = (P[n - 2] - P[n]) / (2 * n - 1)
        dW[n] = -P[n - 1]
    return W, dW


def quadrature_matrices(basis: int, m: int, N: int, h, a=None):
    """Dense mass and weak-Laplacian (= -stiffness) of the C1 / Dirichlet basis on a mesh of m elements
    (scalar h: uniform; array: element lengths) with blocks 1..N, computed by Gauss-Legendre quadrature."""
    q = N - 1
    hk = _element_steps(h, m)
    ng = N + 4
    xg, wg = np.polynomial.legendre.leggauss(ng)
    W, dW = bubble_values(xg, N)
    nhC = m + 1
    nfull = nhC + m * q
    Mm = np.zeros((nfull, nfull))
    Km = np.zeros((nfull, nfull))
    hat = np.stack([(1 - xg) / 2, (1 + xg) / 2])
    dhat = np.stack([np.full_like(xg, -0.5), np.full_like(xg, 0.5)])
    for k in range(m):
        idx = [k, k + 1] + [nhC + j * m + k for j in range(q)]
        F = np.vstack([hat, W[2:N + 1]])
        dF = np.vstack([dhat, dW[2:N + 1]])
        Ml = (F * wg) @ F.T * (hk[k] / 2)
        Kl = (dF * wg) @ dF.T * (2 / hk[k]) * (1.0 if a is None else a[k])   # a: diffusion coefficient of the element
        Mm[np.ix_(idx, idx)] += Ml
        Km[np.ix_(idx, idx)] += Kl
    if basis == DIRICHLET:
        keep = [i for i in range(nfull) if i not in (0, m)]
        Mm = Mm[np.ix_(keep, keep)]
        Km = Km[np.ix_(keep, keep)]
    return Mm, -Km


# ---------------------------------------------------------------------------
# Coefficients / evaluation needed only for the reference's golden value
# ---------------------------------------------------------------------------


def c1_expand(f, df, points: np.ndarray, N: int) -> np.ndarray:
    """Coefficients of f in ContinuousPolynomial{1}(points) blocks 1..N (what
    `transform(C, f)[KR]` converges to): hats interpolate at the nodes, bubble K=n takes
    -c_{n-1} where c_l are the Legendre coefficients of d f/d xi on the element."""
    m = len(points) - 1
    q = N - 1
    nh = m + 1
    c = np.zeros(nh + m * q)
    c[:nh] = [f(p) for p in points]
    xg, wg = np.polynomial.legendre.leggauss(N + 30)
    P = _legendre_table(xg, N)
    for k in range(m):
        a, b = points[k], points[k + 1]
        x = (a + b) / 2 + (b - a) / 2 * xg
        g = np.array([df(t) for t in x]) * (b - a) / 2
        for n in range(2, N + 1):
            l = n - 1
            cl = (2 * l + 1) / 2 * np.sum(wg * g * P[l])
            c[nh + (n - 2) * m + k] = -cl
    return c


def c1_evaluate(c: np.ndarray, points: np.ndarray, N: int, x: float) -> float:
    """(C[:,KR]*c)[x], src/continuouspolynomial.jl:39-53."""
    m = len(points) - 1
    nh = m + 1
    k = min(max(int(np.searchsorted(points, x, side="right")) - 1, 0), m - 1)
    a, b = points[k], points[k + 1]
    xi = (2 * x - a - b) / (b - a)
    val = c[k] * (1 - xi) / 2 + c[k + 1] * (1 + xi) / 2
    W, _ = bubble_values(np.array([xi]), N)
    for n in range(2, N + 1):
        val += c[nh + (n - 2) * m + k] * W[n, 0]
    return float(val)


# ---------------------------------------------------------------------------
# Right-hand-side pipeline of examples/adi_poisson.jl:60-76,105: per-cell Legendre
# analysis, the conversion P \ C (src/continuouspolynomial.jl:226-235) and the weak form (Q'P) fp (Q'P)'
# ---------------------------------------------------------------------------


def legendre_grid_transform(p: int, kind: str = "gauss"):
    """Nodes z (reference cell [-1,1]) and the p x p analysis matrix T with  Legendre coefficients = T @ values:
    what `plan_grid_transform(legendre(..), (p, p))` (examples/adi_poisson.jl:64) provides.  T = V^-1 with the
    Legendre Vandermonde V[s,a] = P_a(z_s); the node family is the caller's (`gauss`: Gauss-Legendre,
    `chebyshev1`: first-kind Chebyshev points), the coefficients of a resolved function do not depend on it."""
    if kind == "gauss":
        z, w = np.polynomial.legendre.leggauss(p)
        V = np.polynomial.legendre.legvander(z, p - 1)
        T = ((2 * np.arange(p) + 1) / 2)[:, None] * V.T * w[None, :]
    elif kind == "chebyshev1":
        z = np.cos(np.pi * (2 * np.arange(p)[::-1] + 1) / (2 * p))
        V = np.polynomial.legendre.legvander(z, p - 1)
        T = np.linalg.inv(V)
    else:
        raise ValueError(kind)
    return z, T


def analysis_2d(vals: np.ndarray, n: int, p: int, T: np.ndarray) -> np.ndarray:
    """examples/adi_poisson.jl:60-76: vals[i*p+s, j*p+t] = f at node s of cell i (x) and node t of cell j (y);
    returns F with F[i + n*a, j + n*b] = sum_st T[a,s] vals[...] T[b,t]  (`F[i+1:n:n*p, j+1:n:n*p] = T * f_.(z, z')`):
    coefficients in P (x) P, P = ContinuousPolynomial{0}, interlaced like every block vector here."""
    V4 = np.asarray(vals, dtype=np.float64).reshape(n, p, n, p)
    C = np.einsum("as,isjt,bt->aibj", T, V4, T, optimize=True)
    return C.reshape(n * p, n * p)


def lowering_dense(basis: int, m: int, N: int) -> np.ndarray:
    """(P \\ C)[KR,KR] (src/continuouspolynomial.jl:226-235) resp. (P \\ Q)[KR,KR] = (P\\C)(C\\Q) (src/dirichlet.jl:35-53),
    KR = Block.(1:N), dense: rows = piecewise Legendre degree a (0..N-1) of element k at a*m+k, columns = the
    C1 / Dirichlet ordering.  A = [1/2 1/2] bidiagonal, B_1 = I/3, C_1 = [-1/2 +1/2], D = tridiag(v_j above, -v_j
    below), v_j = 1/(2j+1): hat_k = (P_0 - P_1)/2, hat_{k+1} = (P_0 + P_1)/2, W_{j+1} = (P_{j-1} - P_{j+1})/(2j+1)."""
    q = N - 1
    nhC = m + 1
    R = np.zeros((m * N, nhC + m * q))
    for k in range(m):
        R[k, k] = R[k, k + 1] = 0.5                       # degree 0
        if N >= 2:
            R[m + k, k], R[m + k, k + 1] = -0.5, 0.5      # degree 1
        for j in range(1, q + 1):                         # bubble j <-> W_{j+1}
            v = 1.0 / (2 * j + 1)
            col = nhC + (j - 1) * m + k
            R[(j - 1) * m + k, col] += v
            if j + 1 <= N - 1:
                R[(j + 1) * m + k, col] -= v
    if basis == DIRICHLET:
        keep = [c for c in range(R.shape[1]) if c not in (0, m)]
        R = R[:, keep]
    return R


def lowering_by_quadrature(basis: int, m: int, N: int) -> np.ndarray:
    """The same matrix from Gauss quadrature of the basis functions against P_a (independent of the closed form)."""
    q = N - 1
    xg, wg = np.polynomial.legendre.leggauss(N + 4)
    Pt = _legendre_table(xg, N)
    W, _ = bubble_values(xg, N)
    funcs = np.vstack([(1 - xg) / 2, (1 + xg) / 2, W[2:N + 1]])          # hat_k, hat_{k+1}, W_2..W_N on one element
    coef = ((2 * np.arange(N) + 1) / 2)[:, None] * ((Pt[:N] * wg) @ funcs.T)   # [a, func]
    nhC = m + 1
    R = np.zeros((m * N, nhC + m * q))
    for k in range(m):
        rows = np.arange(N) * m + k
        R[rows, k] += coef[:, 0]
        R[rows, k + 1] += coef[:, 1]
        for j in range(1, q + 1):
            R[rows, nhC + (j - 1) * m + k] += coef[:, 1 + j]
    if basis == DIRICHLET:
        keep = [c for c in range(R.shape[1]) if c not in (0, m)]
        R = R[:, keep]
    return R


def legendre_mass_diag(m: int, N: int, h) -> np.ndarray:
    """grammatrix(ContinuousPolynomial{0})[KR,KR] (src/continuouspolynomial.jl:259-272): diagonal h_k/(2a+1) at a*m+k."""
    hk = _element_steps(h, m)
    return (hk[None, :] / (2 * np.arange(N)[:, None] + 1)).reshape(-1)


def weakform_dense(basis: int, m: int, N: int, h) -> np.ndarray:
    """(Q'P)[KR,KR] = (P\\Q)' (P'P)  (examples/adi_poisson.jl:105, src/continuouspolynomial.jl:299-304)."""
    return lowering_dense(basis, m, N).T * legendre_mass_diag(m, N, h)[None, :]


def weakform_rhs(basis: int, m: int, N: int, h, fp: np.ndarray) -> np.ndarray:
    """F = (Q'P) fp (Q'P)'  (examples/adi_poisson.jl:105)."""
    A = weakform_dense(basis, m, N, h)
    return A @ fp @ A.T


def c1_from_legendre(basis: int, m: int, N: int, Y: np.ndarray) -> np.ndarray:
    """Coefficients in ContinuousPolynomial{1} / DirichletPolynomial blocks 1..N-1 of the continuous piecewise
    polynomial whose Legendre coefficients (N degrees per element, P ordering) are the columns of Y: the
    `Pl.R \\ dat` + interlacing of ContinuousPolynomialTransform (src/continuouspolynomial.jl:84-146,186-213).
    Restated as the least-squares solution with the truncated conversion P \\ C (exact for data in its range), which
    is independent of the back-substitution the CUDA kernel uses."""
    R = lowering_dense(basis, m, N)
    nh = m + 1 if basis == C1 else m - 1
    R = R[:, :nh + m * (N - 2)]                 # W_N needs Legendre degree N: not representable with N degrees
    return np.linalg.lstsq(R, np.asarray(Y, dtype=np.float64), rcond=None)[0]


def legendre_from_c1(basis: int, m: int, N: int, X: np.ndarray) -> np.ndarray:
    """The `Pl.R \\ dat` step of the INVERSE transform `Pl \\ cfs` (src/continuouspolynomial.jl:129-146,171-184;
    src/dirichlet.jl:97-110): coefficients in C1 / Dirichlet blocks 1..N-1 (columns of X) -> Legendre coefficients, N
    degrees per element in P ordering = the truncated conversion (P \\ C)[Block.(1:N), Block.(1:N-1)]."""
    nh = m + 1 if basis == C1 else m - 1
    return lowering_dense(basis, m, N)[:, :nh + m * (N - 2)] @ np.asarray(X, dtype=np.float64)


def synthesis_2d(F: np.ndarray, n: int, p: int, S: np.ndarray) -> np.ndarray:
    """`legendretransform \\ dat` for a 2D field (the inverse of analysis_2d): F[i + n a, j + n b] = tensor Legendre
    coefficient (a, b) of cell (i, j), S[s, a] = P_a(z_s); returns vals[i p + s, j p + t]."""
    C4 = np.asarray(F, dtype=np.float64).reshape(p, n, p, n)              # [a, i, b, j]
    V = np.einsum("sa,aibj,tb->isjt", S, C4, S, optimize=True)
    return V.reshape(n * p, n * p)


def heat_theta_dense(Md, Kd, u0, dt, theta, nsteps):
    """u <- (M + theta dt K)^-1 (M - (1-theta) dt K) u, nsteps times (examples/heat.jl:26-45 for theta = 1/2, :58-63 for 1)."""
    A = Md + theta * dt * Kd
    B = Md - (1 - theta) * dt * Kd
    u = np.array(u0, dtype=np.float64)
    for _ in range(nsteps):
        u = np.linalg.solve(A, B @ u)
    return u


def basis_eval_matrix(basis: int, points: np.ndarray, N: int, x: np.ndarray) -> np.ndarray:
    """Q[x, KR] / C[x, KR]: values of the basis functions (columns, arrowhead order) at the points x."""
    m = len(points) - 1
    q = N - 1
    nhC = m + 1
    E = np.zeros((len(x), nhC + m * q))
    for r, xv in enumerate(x):
        k = min(max(int(np.searchsorted(points, xv, side="right")) - 1, 0), m - 1)
        a, b = points[k], points[k + 1]
        xi = (2 * xv - a - b) / (b - a)
        E[r, k], E[r, k + 1] = (1 - xi) / 2, (1 + xi) / 2
        Wv, _ = bubble_values(np.array([xi]), N)
        for n_ in range(2, N + 1):
            E[r, nhC + (n_ - 2) * m + k] = Wv[n_, 0]
    if basis == DIRICHLET:
        keep = [c for c in range(E.shape[1]) if c not in (0, m)]
        E = E[:, keep]
    return E


# ---------------------------------------------------------------------------
# ADI driver (examples/adi_poisson.jl:11-59) and heat stepping (examples/heat.jl:58-63)
# ---------------------------------------------------------------------------


def mobius(z, a, b, c, d, alpha):
    """examples/adi_poisson.jl:11-18."""
    t1 = a * (-alpha * b + b + alpha * c + c) - 2 * b * c
    t2 = a * (alpha * (b + c) - b + c) - 2 * alpha * b * c
    t3 = 2 * a - (alpha + 1) * b + (alpha - 1) * c
    t4 = -alpha * (-2 * a + b + c) - b + c
    return (t1 * z + t2) / (t3 * z + t4)


def adi_num_iterations(a, b, c, d, tol=1e-15):
    """examples/adi_poisson.jl:47-48."""
    gamma = (c - a) * (d - b) / ((c - b) * (d - a))
    return int(math.ceil(math.log(16 * gamma) * math.log(4 / tol) / math.pi ** 2))


def adi_shifts(J, a, b, c, d):
    """examples/adi_poisson.jl:23-40 (Elliptic.K / Jacobi.dn -> scipy.special)."""
    from scipy.special import ellipj, ellipk
    gamma = (c - a) * (d - b) / ((c - b) * (d - a))
    alpha = -1 + 2 * gamma + 2 * math.sqrt(gamma ** 2 - gamma)
    if alpha < 1e7:
        mpar = 1 - 1 / alpha ** 2
        K = ellipk(mpar)
        u = (2 * np.arange(J) + 1) * K / (2 * J)
        dn = ellipj(u, mpar)[2]
    else:
        K = 2 * math.log(2) + math.log(alpha) + (-1 + 2 * math.log(2) + math.log(alpha)) / alpha ** 2 / 4
        m1 = 1 / alpha ** 2
        u = (np.arange(J) + 0.5) * K / J
        dn = 1 / np.cosh(u) + m1 / 4 * (np.sinh(u) * np.cosh(u) + u) * np.tanh(u) / np.cosh(u)
    p = np.array([mobius(-alpha * i, a, b, c, d, alpha) for i in dn])
    qq = np.array([mobius(alpha * i, a, b, c, d, alpha) for i in dn])
    return p, qq


def adi_dense(A, B, C, F, a, b, c, d, tol=1e-15, J=None):
    """The reference iteration (examples/adi_poisson.jl:42-59) with dense solves (L0)."""
    X = np.zeros_like(F)
    if J is None:
        J = adi_num_iterations(a, b, c, d, tol)
    p, qq = adi_shifts(J, a, b, c, d)
    for j in range(J):
        X = np.linalg.solve((C - B / p[j]).T, (((A / p[j] - C) @ X - F / p[j])).T).T
        X = np.linalg.solve(C - A / qq[j], X @ (B / qq[j] - C) - F / qq[j])
    return X


def adi_packed(Mp: Packed, Kp: Packed, F, a, b, c, d, tol=1e-15, J=None):
    """Same iteration with A = M, B = -M, C = K (examples/adi_poisson.jl:107) through the
    packed structured kernels: 2 shifted factorisations, 2 solves, 2 products per step."""
    n = Mp.n
    X = np.zeros((n, n))
    if J is None:
        J = adi_num_iterations(a, b, c, d, tol)
    p, qq = adi_shifts(J, a, b, c, d)
    for j in range(J):
        # X = ((A/p - C) X - F/p) / revchol(C - B/p)       A=M, B=-M, C=K
        T1 = packed_symmetrize(packed_axpby(1.0 / p[j], Mp, -1.0, Kp))
        R = np.zeros_like(X)
        packed_mul_left(1.0, T1, X, 0.0, R)
        R -= F / p[j]
        U1 = packed_revchol(packed_axpby(1.0, Kp, 1.0 / p[j], Mp))
        Rt = np.ascontiguousarray(R.T)
        packed_solve(U1, Rt)
        X = Rt.T
        # X = revchol(C - A/q) \ (X (B/q - C) - F/q)
        T2 = packed_symmetrize(packed_axpby(-1.0 / qq[j], Mp, -1.0, Kp))
        Rt = np.zeros_like(X)
        packed_mul_left(1.0, T2, np.ascontiguousarray(X.T), 0.0, Rt)   # (X T2)^T = T2 X^T (T2 symmetric)
        R = Rt.T - F / qq[j]
        U2 = packed_revchol(packed_axpby(1.0, Kp, -1.0 / qq[j], Mp))
        R = np.ascontiguousarray(R)
        packed_solve(U2, R)
        X = R
    return X


def heat_steps_packed(Mp: Packed, Kp: Packed, u0, dt: float, nsteps: int):
    """Backward Euler with one reused factorisation, examples/heat.jl:58-63:
    A = factorize(M - h*Delta); u <- A \\ (M*u).  Kp is -Delta (positive)."""
    U = packed_revchol(packed_axpby(1.0, Mp, dt, Kp))
    Ms = packed_symmetrize(Mp)
    u = np.array(u0, dtype=float, copy=True)
    for _ in range(nsteps):
        r = np.zeros_like(u)
        packed_mul_left(1.0, Ms, u, 0.0, r)
        packed_solve(U, r)
        u = r
    return u


def heat2d_steps_packed(Mp: Packed, Kp: Packed, U0, dt: float, nsteps: int):
    """2D extension: factored (LOD) backward Euler
    U <- S^{-1} (M U M) S^{-1}, S = M + dt K factored once."""
    Uf = packed_revchol(packed_axpby(1.0, Mp, dt, Kp))
    Ms = packed_symmetrize(Mp)
    U = np.array(U0, dtype=float, copy=True)
    for _ in range(nsteps):
        for _dim in range(2):
            R = np.zeros_like(U)
            packed_mul_left(1.0, Ms, U, 0.0, R)
            packed_solve(Uf, R)
            U = np.ascontiguousarray(R.T)
    return U
