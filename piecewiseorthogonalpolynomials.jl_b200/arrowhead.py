"""Host-side mirror of the reference's arrowhead API on top of the C ABI.

Names follow /root/reference/src/arrowhead.jl: BBBArrowheadMatrix (:31-41), Symmetric /
UpperTriangular / LowerTriangular / Unit* field views (:154-167, :489-521), `+ - * /` (:392-417),
mul (:191-258), reversecholesky (:266-292), triangular ldiv (:425-486).  Python spelling of the
Julia operators:   A @ X -> A*X,  X @ A -> X*A,  ldiv(F, X) -> F \\ X,  rdiv(X, F) -> X / F.

Matrices/vectors may be NumPy arrays (host buffers: the *_host entry points copy in and out)
or CUDA torch tensors in column-major layout (device pointers, no copies).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import Context, DimensionMismatch, PopError, PosDefException, StructureError, default_context  # noqa: F401


# --------------------------------------------------------------------------------------
# packing of the reference's block form (A, B-tuple, C-tuple, D-vector) into the ABI layout
# --------------------------------------------------------------------------------------
def _bandwidths(X: np.ndarray):
    r, c = np.nonzero(X)
    if r.size == 0:
        return 0, 0
    return int(max(0, (r - c).max())), int(max(0, (c - r).max()))


def _pack_blocks(A, B, Cb, D):
    A = np.asarray(A, dtype=np.float64)
    nh = A.shape[0]
    if A.ndim != 2 or A.shape[1] != nh:
        raise DimensionMismatch(-2, "A must be square")
    lA, uA = _bandwidths(A)
    if lA > 1 or uA > 1:                     # @assert -1 <= lambda,mu <= 1, src/arrowhead.jl:94-97
        raise StructureError(-3, "A must have bandwidths within (1,1)")
    D = [np.asarray(d, dtype=np.float64) for d in D]
    m = len(D)
    if m == 0:
        raise DimensionMismatch(-2, "D must hold at least one block")
    q = D[0].shape[0]
    for d in D:
        if d.shape != (q, q):
            raise DimensionMismatch(-2, "all D blocks must be q x q")
    bws = [_bandwidths(d) for d in D]
    lD, uD = max(b[0] for b in bws), max(b[1] for b in bws)
    Ad = np.ascontiguousarray(np.diag(A))
    Au = np.ascontiguousarray(np.diag(A, 1))
    Al = np.ascontiguousarray(np.diag(A, -1))

    def slots(blocks, transpose, name):
        out = np.zeros((len(blocks), 3, m))
        for b, blk in enumerate(blocks):
            Bm = np.asarray(blk, dtype=np.float64)
            Bm = Bm.T if transpose else Bm
            if Bm.shape != (nh, m):           # @assert size(op) == (xi, m), :101 / :107
                raise DimensionMismatch(-2, f"{name}[{b}] has the wrong size")
            chk = np.zeros_like(Bm)
            for t in range(3):
                k = np.arange(m)
                i = k + t - 1
                ok = (i >= 0) & (i < nh)
                out[b, t, ok] = Bm[i[ok], k[ok]]
                chk[i[ok], k[ok]] = Bm[i[ok], k[ok]]
            if not np.array_equal(chk, Bm):   # :102-104 / :108-110
                raise StructureError(-3, f"{name}[{b}] must have bandwidths within (1,1)")
        return out

    Bp = slots(B, False, "B")
    Cp = slots(Cb, True, "C")
    Dp = np.zeros((lD + uD + 1, q, m))
    for k, d in enumerate(D):
        for dd in range(-lD, uD + 1):
            if abs(dd) >= q:
                continue
            diag = np.diag(d, dd)
            if dd >= 0:
                Dp[lD + dd, :q - dd, k] = diag
            else:
                Dp[lD + dd, -dd:, k] = diag
    return dict(m=m, nh=nh, q=q, lD=lD, uD=uD, Ad=Ad, Au=Au, Al=Al, B=Bp, C=Cp, D=Dp)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class BBBArrowheadMatrix:
    """Device-resident block-banded-block arrowhead matrix (src/arrowhead.jl:31-41).

    BBBArrowheadMatrix(A, B, C, D): A (nh x nh, bandwidths <= (1,1)), B tuple of nh x m blocks,
    C tuple of m x nh blocks, D list of m banded q x q blocks (all the same bandwidths)."""

    __array_ufunc__ = None      # let `ndarray @ A` reach __rmatmul__

    def __init__(self, A=None, B=(), C=(), D=None, *, ctx: Optional[Context] = None, _handle=None):
        self.ctx = ctx or default_context()
        if _handle is not None:
            self._h = _handle
            return
        p = _pack_blocks(A, B, C, D)
        self._h = self._upload(self.ctx, p)

    @staticmethod
    def _upload(ctx, p, uniform=False):
        desc = L.pop_desc(p["m"], p["nh"], p["q"], p["B"].shape[0], p["C"].shape[0], p["lD"], p["uD"], 1 if uniform else 0)
        h = C.c_void_p()
        arrs = [np.ascontiguousarray(p[k], dtype=np.float64) for k in ("Ad", "Au", "Al", "B", "C", "D")]
        L.check(L.lib().pop_arrowhead_upload(ctx.handle, C.byref(desc), *[_ptr(a) for a in arrs], C.byref(h)), "BBBArrowheadMatrix")
        return h

    @classmethod
    def from_packed(cls, m, nh, q, lD, uD, Ad, Au=None, Al=None, B=None, C_=None, D=None, uniform=False, ctx=None):
        """Construct from the packed ABI layout (include/pop_arrowhead.h)."""
        ctx = ctx or default_context()
        z = lambda shape: np.zeros(shape)
        p = dict(m=m, nh=nh, q=q, lD=lD, uD=uD, Ad=np.asarray(Ad, float), Au=z(max(nh - 1, 0)) if Au is None else np.asarray(Au, float),
                 Al=z(max(nh - 1, 0)) if Al is None else np.asarray(Al, float), B=z((0, 3, m)) if B is None else np.asarray(B, float),
                 C=z((0, 3, m)) if C_ is None else np.asarray(C_, float), D=np.asarray(D, float))
        return cls(ctx=ctx, _handle=cls._upload(ctx, p, uniform))

    # -- shape ------------------------------------------------------------------------
    @property
    def desc(self):
        d = L.pop_desc()
        L.check(L.lib().pop_arrowhead_desc(self._h, C.byref(d)))
        return d

    @property
    def n(self):
        d = self.desc
        return d.nh + d.m * d.q

    @property
    def shape(self):
        return (self.n, self.n)

    def blockaxes(self):
        """Block sizes (nh, m, m, ...), src/arrowhead.jl:57-62."""
        d = self.desc
        return (d.nh,) + (d.m,) * d.q

    def packed(self):
        """Download the packed fields (dict of NumPy arrays, D expanded to [bands][q][m])."""
        d = self.desc
        out = dict(m=d.m, nh=d.nh, q=d.q, lD=d.lD, uD=d.uD, Ad=np.zeros(d.nh), Au=np.zeros(max(d.nh - 1, 0)), Al=np.zeros(max(d.nh - 1, 0)),
                   B=np.zeros((d.nB, 3, d.m)), C=np.zeros((d.nC, 3, d.m)), D=np.zeros((d.lD + d.uD + 1, d.q, d.m)))
        L.check(L.lib().pop_arrowhead_download(self.ctx.handle, self._h, 1, *[_ptr(out[k]) for k in ("Ad", "Au", "Al", "B", "C", "D")]))
        return out

    def to_dense(self):
        """Matrix(A) via getindex semantics (src/arrowhead.jl:70-87)."""
        p = self.packed()
        m, nh, q, lD, uD = p["m"], p["nh"], p["q"], p["lD"], p["uD"]
        n = nh + m * q
        S = np.zeros((n, n))
        S[np.arange(nh), np.arange(nh)] = p["Ad"]
        if nh > 1:
            S[np.arange(nh - 1), np.arange(1, nh)] = p["Au"]
            S[np.arange(1, nh), np.arange(nh - 1)] = p["Al"]
        k = np.arange(m)
        for t in range(3):
            i = k + t - 1
            ok = (i >= 0) & (i < nh)
            for b in range(min(p["B"].shape[0], q)):
                S[i[ok], nh + b * m + k[ok]] = p["B"][b, t, ok]
            for b in range(min(p["C"].shape[0], q)):
                S[nh + b * m + k[ok], i[ok]] = p["C"][b, t, ok]
        for dd in range(-lD, uD + 1):
            for j in range(q):
                if 0 <= j + dd < q:
                    S[nh + j * m + k, nh + (j + dd) * m + k] = p["D"][lD + dd, j]
        return S

    # -- algebra (src/arrowhead.jl:392-417) ---------------------------------------------
    def _axpby(self, a, other, b):
        h = C.c_void_p()
        L.check(L.lib().pop_axpby(self.ctx.handle, float(a), self._h, float(b), other._h, C.byref(h)))
        return BBBArrowheadMatrix(ctx=self.ctx, _handle=h)

    def __add__(self, o):
        return self._axpby(1.0, _plain(o), 1.0)

    def __sub__(self, o):
        return self._axpby(1.0, _plain(o), -1.0)

    def __neg__(self):
        return self._axpby(-1.0, self, 0.0)

    def __mul__(self, c):
        return self._axpby(float(c), self, 0.0)

    __rmul__ = __mul__

    def __truediv__(self, c):
        return self._axpby(1.0 / float(c), self, 0.0)

    def copy(self):
        h = C.c_void_p()
        L.check(L.lib().pop_arrowhead_copy(self.ctx.handle, self._h, C.byref(h)))
        return BBBArrowheadMatrix(ctx=self.ctx, _handle=h)

    @property
    def T(self):
        return _Transposed(self)

    # -- products ------------------------------------------------------------------------
    def __matmul__(self, X):
        return mul(self, X)

    def __rmatmul__(self, X):
        return mul(X, self)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().pop_arrowhead_free(self.ctx.handle, self._h)
                self._h = None
        except Exception:
            pass


class _Wrapper:
    sym = L.GENERAL
    trans = 0
    __array_ufunc__ = None

    def __init__(self, data: BBBArrowheadMatrix):
        self.data = data

    @property
    def ctx(self):
        return self.data.ctx

    @property
    def n(self):
        return self.data.n

    @property
    def shape(self):
        return self.data.shape

    def __matmul__(self, X):
        return mul(self, X)

    def __rmatmul__(self, X):
        return mul(X, self)


class _Transposed(_Wrapper):
    trans = 1

    def to_dense(self):
        return self.data.to_dense().T


class Symmetric(_Wrapper):
    """Symmetric(P): reads the upper data of P (src/arrowhead.jl:154-167)."""
    sym = L.SYM_UPPER

    def to_dense(self):
        U = np.triu(self.data.to_dense())
        return U + np.triu(U, 1).T

    def _lin(self, a, o, b):
        if not isinstance(o, Symmetric):
            raise TypeError("Symmetric arrowheads combine with Symmetric arrowheads")
        return Symmetric(self.data._axpby(a, o.data, b))

    def __add__(self, o):
        return self._lin(1.0, o, 1.0)

    def __sub__(self, o):
        return self._lin(1.0, o, -1.0)

    def __neg__(self):
        return Symmetric(-self.data)

    def __mul__(self, c):
        return Symmetric(self.data * c)

    __rmul__ = __mul__

    def __truediv__(self, c):
        return Symmetric(self.data / c)

    @property
    def parent(self):
        return self.data


class _Triangular(_Wrapper):
    uplo = L.UPPER
    unit = 0

    def __init__(self, data, trans=0):
        super().__init__(data)
        self.trans = trans

    @property
    def T(self):
        return type(self)(self.data, 1 - self.trans)

    def to_dense(self):
        S = self.data.to_dense()
        Tm = np.triu(S) if self.uplo == L.UPPER else np.tril(S)
        if self.unit:
            np.fill_diagonal(Tm, 1.0)
        return Tm.T if self.trans else Tm


class UpperTriangular(_Triangular):
    """src/arrowhead.jl:489-504: fields A->upper, B kept, C empty, D->upper."""


class UnitUpperTriangular(_Triangular):
    unit = 1


class LowerTriangular(_Triangular):
    uplo = L.LOWER


class UnitLowerTriangular(_Triangular):
    uplo = L.LOWER
    unit = 1


def _plain(A):
    if isinstance(A, BBBArrowheadMatrix):
        return A
    raise TypeError(f"expected a BBBArrowheadMatrix, got {type(A).__name__}")


# --------------------------------------------------------------------------------------
# panels: NumPy (host) or torch CUDA tensors (device)
# --------------------------------------------------------------------------------------
def _is_torch(X):
    return hasattr(X, "data_ptr") and hasattr(X, "is_cuda")


class _Panel:
    """Column-major view of a user array: pointer, leading dimension, rows, cols."""

    def __init__(self, X, writable):
        self.torch = _is_torch(X)
        self.vector = X.ndim == 1
        if self.torch:
            if not X.is_cuda:
                raise TypeError("torch tensors must live on the GPU; pass NumPy arrays for host data")
            if str(X.dtype) != "torch.float64":
                raise TypeError("Float64 only")
            if self.vector:
                if X.stride(0) != 1:
                    raise ValueError("vector must be contiguous")
                self.rows, self.cols, self.ld = X.shape[0], 1, max(X.shape[0], 1)
            else:
                if X.ndim != 2 or (X.shape[0] > 1 and X.stride(0) != 1):
                    raise ValueError("device matrices must be column-major (stride(0) == 1); use x.t().contiguous().t()")
                self.rows, self.cols = X.shape
                self.ld = X.stride(1) if X.shape[1] > 1 else max(self.rows, 1)
            self.ptr = X.data_ptr()
            self.keep = X
        else:
            X = np.asarray(X)
            if X.dtype != np.float64:
                raise TypeError("Float64 only")
            if self.vector:
                if not X.flags.c_contiguous:
                    raise ValueError("vector must be contiguous")
                self.rows, self.cols, self.ld = X.shape[0], 1, max(X.shape[0], 1)
            else:
                if X.ndim != 2 or not X.flags.f_contiguous:
                    raise ValueError("host matrices must be Fortran (column-major) ordered; use np.asfortranarray")
                self.rows, self.cols, self.ld = X.shape[0], X.shape[1], max(X.shape[0], 1)
            if writable and not X.flags.writeable:
                raise ValueError("output array is read-only")
            self.ptr = X.ctypes.data
            self.keep = X


def _new_like(X, shape):
    if _is_torch(X):
        import torch
        # column-major device buffer with an even leading dimension (16 B aligned columns)
        rows = shape[0]
        cols = shape[1] if len(shape) > 1 else 1
        ld = (rows + 1) & ~1
        buf = torch.empty((cols, ld), dtype=torch.float64, device=X.device)
        out = buf.t()[:rows]
        return out[:, 0] if len(shape) == 1 else out
    return np.empty(shape, dtype=np.float64, order="F")


def _copy_like(X):
    Y = _new_like(X, tuple(X.shape))
    if _is_torch(X):
        Y.copy_(X)
    else:
        Y[...] = X
    return Y


def _op_of(A):
    """(handle owner, sym, trans) of a BBBArrowheadMatrix / Symmetric / transposed wrapper."""
    if isinstance(A, BBBArrowheadMatrix):
        return A, L.GENERAL, 0
    if isinstance(A, (Symmetric, _Transposed)):
        return A.data, A.sym, A.trans
    raise TypeError(f"not an arrowhead operator: {type(A).__name__}")


def _is_op(A):
    return isinstance(A, (BBBArrowheadMatrix, _Wrapper))


def mul(A, X, alpha=1.0, beta=0.0, out=None):
    """mul!(out, A, X, alpha, beta): A*X (arrowhead left) or X*A (arrowhead right);
    src/arrowhead.jl:191-258."""
    if _is_op(A):
        side, op, mat = L.LEFT, A, X
    elif _is_op(X):
        side, op, mat = L.RIGHT, X, A
    else:
        raise TypeError("one operand must be an arrowhead")
    base, sym, trans = _op_of(op)
    n = base.n
    px = _Panel(mat, False)
    if side == L.RIGHT and px.vector:
        # row vector times matrix:  x' A = (A' x)'
        side, trans = L.LEFT, 1 - trans
    vec_len = px.rows if side == L.LEFT else px.cols
    if vec_len != n:
        raise DimensionMismatch(-2, f"DimensionMismatch: arrowhead is {n}x{n}, operand has {vec_len}")
    if out is None:
        if beta != 0.0:
            raise ValueError("beta != 0 needs `out`")
        out = _new_like(mat, tuple(mat.shape))
    py = _Panel(out, True)
    if (py.rows, py.cols) != (px.rows, px.cols) or py.torch != px.torch:
        raise DimensionMismatch(-2, "DimensionMismatch: out has the wrong shape")
    nrhs = px.cols if side == L.LEFT else px.rows
    fn = L.lib().pop_mul if px.torch else L.lib().pop_mul_host
    if px.torch:
        _bind_stream(base.ctx)
    L.check(fn(base.ctx.handle, base._h, side, sym, trans, float(alpha), px.ptr, px.ld, float(beta), py.ptr, py.ld, nrhs), "mul")
    return out


def _bind_stream(ctx):
    import torch
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)


def ldiv(T, X, overwrite=False, algo=L.SOLVE_AUTO, out=None):
    """T \\ X for a triangular arrowhead or a ReverseCholesky factorisation
    (src/arrowhead.jl:425-486; test/test_arrowhead.jl:64-68,130).  out: a host array of the shape and memory order of X that
    receives `F \\ X` while X stays untouched (NumPy panels and ReverseCholesky only: the two directions of the host link
    then work on different buffers)."""
    return _div(T, X, L.LEFT, overwrite, algo, out)


def rdiv(X, T, overwrite=False, algo=L.SOLVE_AUTO, out=None):
    """X / T (rows), e.g. `X / reversecholesky(Symmetric(S))` in examples/adi_poisson.jl:54."""
    return _div(T, X, L.RIGHT, overwrite, algo, out)


def _div(T, X, side, overwrite, algo, out=None):
    if out is not None:
        px, po = _Panel(X, False), _Panel(out, True)
        if px.torch or po.torch or not isinstance(T, ReverseCholesky) or px.vector:
            raise TypeError("out= is for host matrices solved with a ReverseCholesky factorisation")
        if (px.rows, px.cols) != (po.rows, po.cols):
            raise DimensionMismatch(-2, f"DimensionMismatch: X is {px.rows}x{px.cols}, out is {po.rows}x{po.cols}")
        nrhs, veclen = (px.cols, px.rows) if side == L.LEFT else (px.rows, px.cols)
        if veclen != T.factor.n:
            raise DimensionMismatch(-2, f"DimensionMismatch: factor is {T.factor.n}x{T.factor.n}, right-hand side has {veclen}")
        L.check(L.lib().pop_solve_host_to(T.factor.ctx.handle, T.factor._h, side, px.ptr, px.ld, po.ptr, po.ld, nrhs, algo), "ldiv")
        return out
    Y = X if overwrite else _copy_like(X)
    p = _Panel(Y, True)
    if p.vector:
        side_eff, nrhs, veclen = L.LEFT, 1, p.rows
        flip = side == L.RIGHT          # x'/T = (T' \ x)'
    else:
        side_eff, flip = side, False
        nrhs = p.cols if side == L.LEFT else p.rows
        veclen = p.rows if side == L.LEFT else p.cols
    if isinstance(T, ReverseCholesky):
        base = T.factor
        if veclen != base.n:
            raise DimensionMismatch(-2, f"DimensionMismatch: factor is {base.n}x{base.n}, right-hand side has {veclen}")
        if p.torch:
            _bind_stream(base.ctx)
        fn = L.lib().pop_solve if p.torch else L.lib().pop_solve_host
        L.check(fn(base.ctx.handle, base._h, side_eff, p.ptr, p.ld, nrhs, algo), "ldiv")
        return Y
    if not isinstance(T, _Triangular):
        raise TypeError("ldiv/rdiv need a triangular arrowhead or a ReverseCholesky")
    base = T.data
    if veclen != base.n:
        raise DimensionMismatch(-2, f"DimensionMismatch: matrix is {base.n}x{base.n}, right-hand side has {veclen}")
    trans = T.trans ^ (1 if flip else 0)
    if p.torch:
        _bind_stream(base.ctx)
    fn = L.lib().pop_trsm if p.torch else L.lib().pop_trsm_host
    L.check(fn(base.ctx.handle, base._h, side_eff, T.uplo, trans, T.unit, p.ptr, p.ld, nrhs), "ldiv")
    return Y


# --------------------------------------------------------------------------------------
# reverse Cholesky (MatrixFactorizations API; src/arrowhead.jl:266-292)
# --------------------------------------------------------------------------------------
class ReverseCholesky:
    """S = U U'.  `.U` is an UpperTriangular arrowhead, `.L` its adjoint; `ldiv(F, x)`,
    `rdiv(X, F)` solve with it."""

    def __init__(self, factor: BBBArrowheadMatrix):
        self.factor = factor

    @property
    def U(self):
        return UpperTriangular(self.factor)

    @property
    def L(self):
        return UpperTriangular(self.factor, trans=1)

    @property
    def ctx(self):
        return self.factor.ctx

    def solve(self, X, overwrite=False, algo=0):
        return ldiv(self, X, overwrite, algo)


def reversecholesky(S) -> ReverseCholesky:
    """reversecholesky(Symmetric(S)) (README.md:34-37): copies, factors, returns ReverseCholesky."""
    base = S.data if isinstance(S, Symmetric) else _plain(S)
    h = C.c_void_p()
    L.check(L.lib().pop_revchol(base.ctx.handle, base._h, C.byref(h)), "reversecholesky")
    return ReverseCholesky(BBBArrowheadMatrix(ctx=base.ctx, _handle=h))


def as_reversecholesky(U) -> ReverseCholesky:
    """ReverseCholesky from an arrowhead whose upper data already hold the factor U (S = U U'), e.g. one computed by the
    reference on the host and uploaded: no second factorisation (pop_arrowhead_as_factor)."""
    base = U.data if isinstance(U, _Wrapper) else _plain(U)
    L.check(L.lib().pop_arrowhead_as_factor(base.ctx.handle, base._h), "as_reversecholesky")
    return ReverseCholesky(base)


def reversecholesky_inplace(S) -> ReverseCholesky:
    """reversecholesky!(Symmetric(S)): overwrites the upper data of S (test/test_arrowhead.jl:90)."""
    base = S.data if isinstance(S, Symmetric) else _plain(S)
    L.check(L.lib().pop_revchol_inplace(base.ctx.handle, base._h), "reversecholesky!")
    return ReverseCholesky(base)


def reversecholesky_shifted(K, M, alpha: Sequence[float], sigma: Sequence[float]):
    """Factors of alpha[i]*K + sigma[i]*M for all i in one batched launch sequence."""
    Kb = K.data if isinstance(K, Symmetric) else _plain(K)
    Mb = M.data if isinstance(M, Symmetric) else _plain(M)
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    s = np.ascontiguousarray(sigma, dtype=np.float64)
    if a.shape != s.shape or a.ndim != 1:
        raise ValueError("alpha and sigma must be 1-D of equal length")
    ns = a.size
    hs = (C.c_void_p * ns)()
    info = np.zeros(ns, dtype=np.int32)
    rc = L.lib().pop_revchol_shifted_batch(Kb.ctx.handle, Kb._h, Mb._h, _ptr(a), _ptr(s), ns, hs, info.ctypes.data_as(C.c_void_p))
    facs = [ReverseCholesky(BBBArrowheadMatrix(ctx=Kb.ctx, _handle=C.c_void_p(hs[i]))) for i in range(ns) if hs[i]]
    L.check(rc, "reversecholesky (batched)")
    return facs
