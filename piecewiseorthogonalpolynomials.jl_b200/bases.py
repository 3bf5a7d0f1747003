"""Bases and the lazy (infinite) mass / weak-Laplacian arrowheads they assemble.

Mirrors, for the hot path only, /root/reference/src/continuouspolynomial.jl:274-290,348-357 and
src/dirichlet.jl:169-196 (closed forms on a uniform mesh) plus the truncation `M[KR,KR]` with
`KR = Block.(oneto(N))` (src/arrowhead.jl:141-149).  The entries are generated on the device by
pop_assemble; nothing is computed on the host.  Non-uniform meshes (the reference's generic
`L'*G*L`, src/continuouspolynomial.jl:259-265,293-297, and `diff`, :318-325) go through pop_assemble_mesh: one
D block per element, factor / solve / product on the per-element kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from .arrowhead import BBBArrowheadMatrix, Symmetric


@dataclass(frozen=True)
class BlockRange:
    """Block.(oneto(N)) -- the only block range the arrowhead `sublayout` accepts
    (src/arrowhead.jl:141-142)."""
    N: int

    def __len__(self):
        return self.N


class Block:
    @staticmethod
    def oneto(N: int) -> BlockRange:
        if N < 1:
            raise ValueError("Block.oneto(N) needs N >= 1")
        return BlockRange(int(N))


def _mesh_steps(points):
    """(uniform?, step, element lengths): a range (`AbstractRange` in the reference) takes the closed forms of
    src/continuouspolynomial.jl:274-290; any other increasing grid is a non-uniform mesh."""
    p = np.asarray(points, dtype=np.float64)
    if p.ndim != 1 or p.size < 2:
        raise ValueError("points must be a 1-D grid with at least two nodes")
    h = np.diff(p)
    if np.any(h <= 0):
        raise ValueError("points must be increasing")
    uniform = bool(np.max(np.abs(h - h[0])) <= 8 * np.finfo(float).eps * max(1.0, np.abs(p).max()))
    return uniform, float((p[-1] - p[0]) / (p.size - 1)), np.ascontiguousarray(h)


class _Basis:
    basis_id = -1

    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64)
        self.uniform, self.h, self.hk = _mesh_steps(self.points)
        self.m = self.points.size - 1

    def __eq__(self, o):
        return type(self) is type(o) and np.array_equal(self.points, o.points)

    def __hash__(self):
        return hash((type(self).__name__, self.points.tobytes()))


class ContinuousPolynomial(_Basis):
    """ContinuousPolynomial{1}(points): hats + bubbles, Neumann-type (src/continuouspolynomial.jl:7-15)."""
    basis_id = L.BASIS_C1

    def __init__(self, order, points=None):
        if points is None:
            order, points = 1, order
        if order != 1:
            raise NotImplementedError("only ContinuousPolynomial{1} is on the arrowhead path")
        super().__init__(points)

    @property
    def nh(self):
        return self.m + 1


class DirichletPolynomial(_Basis):
    """DirichletPolynomial(points): end hats removed (src/dirichlet.jl)."""
    basis_id = L.BASIS_DIRICHLET

    @property
    def nh(self):
        return self.m - 1


class LazyArrowhead:
    """Linear combination of the infinite mass and weak-Laplacian matrices of one basis;
    `A[KR, KR]` materialises the truncation on the device as Symmetric(BBBArrowheadMatrix)."""

    def __init__(self, basis: _Basis, cm: float, cl: float, ctx=None, coef=None):
        self.basis, self.cm, self.cl, self.ctx = basis, float(cm), float(cl), ctx
        self.coef = coef          # per-element diffusion coefficients of the weak-Laplacian part (None: 1)

    def _comb(self, o, s):
        if not isinstance(o, LazyArrowhead) or o.basis != self.basis:
            raise TypeError("lazy arrowheads combine only within one basis")
        ca, cb = (self.coef if self.cl != 0 else None), (o.coef if o.cl != 0 else None)
        if self.cl != 0 and o.cl != 0 and not ((ca is None and cb is None) or (ca is not None and cb is not None and np.array_equal(ca, cb))):
            raise TypeError("weak Laplacians with different diffusion coefficients do not combine lazily")
        return LazyArrowhead(self.basis, self.cm + s * o.cm, self.cl + s * o.cl, self.ctx or o.ctx, ca if ca is not None else cb)

    def __add__(self, o):
        return self._comb(o, 1.0)

    def __sub__(self, o):
        return self._comb(o, -1.0)

    def __neg__(self):
        return LazyArrowhead(self.basis, -self.cm, -self.cl, self.ctx, self.coef)

    def __mul__(self, c):
        return LazyArrowhead(self.basis, self.cm * c, self.cl * c, self.ctx, self.coef)

    __rmul__ = __mul__

    def __truediv__(self, c):
        return LazyArrowhead(self.basis, self.cm / c, self.cl / c, self.ctx, self.coef)

    def __getitem__(self, idx) -> Symmetric:
        KR, JR = idx
        if not (isinstance(KR, BlockRange) and isinstance(JR, BlockRange)) or KR.N != JR.N:
            raise TypeError("arrowhead truncation needs [Block.oneto(N), Block.oneto(N)]")
        ctx = self.ctx or L.default_context()
        hm, hl = C.c_void_p(), C.c_void_p()
        pm, pl = (C.byref(hm) if self.cm != 0 else None), (C.byref(hl) if self.cl != 0 else None)
        coef = self.coef if self.cl != 0 else None
        if coef is not None:
            hk = self.basis.hk if not self.basis.uniform else np.full(self.basis.m, float(self.basis.h))
            L.check(L.lib().pop_assemble_mesh_coef(ctx.handle, self.basis.basis_id, self.basis.m, KR.N, hk.ctypes.data_as(C.c_void_p),
                                                   coef.ctypes.data_as(C.c_void_p), pm, pl), "assemble")
        elif self.basis.uniform:
            L.check(L.lib().pop_assemble(ctx.handle, self.basis.basis_id, self.basis.m, KR.N, self.basis.h, pm, pl), "assemble")
        else:
            hk = self.basis.hk
            L.check(L.lib().pop_assemble_mesh(ctx.handle, self.basis.basis_id, self.basis.m, KR.N, hk.ctypes.data_as(C.c_void_p), pm, pl), "assemble")
        Mh = BBBArrowheadMatrix(ctx=ctx, _handle=hm) if self.cm != 0 else None
        Lh = BBBArrowheadMatrix(ctx=ctx, _handle=hl) if self.cl != 0 else None
        if Mh is not None and Lh is not None:
            return Symmetric(Mh._axpby(self.cm, Lh, self.cl))
        if Mh is not None:
            return Symmetric(Mh if self.cm == 1.0 else Mh * self.cm)
        if Lh is not None:
            return Symmetric(Lh if self.cl == 1.0 else Lh * self.cl)
        raise ValueError("zero operator")


def grammatrix(basis: _Basis, ctx=None) -> LazyArrowhead:
    """M = C'C (src/continuouspolynomial.jl:274-290, src/dirichlet.jl:180-196)."""
    return LazyArrowhead(basis, 1.0, 0.0, ctx)


def weaklaplacian(basis: _Basis, ctx=None, a=None) -> LazyArrowhead:
    """Delta = weaklaplacian(C), negative semi-definite (src/continuouspolynomial.jl:348-357,
    src/dirichlet.jl:169-178).  a: one positive diffusion coefficient per element, the piecewise-constant case of
    `-(D*C)' * (a .* (D*C))` (examples/heat.jl:77-79)."""
    if a is not None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (basis.m,):
            raise L.DimensionMismatch(-2, f"DimensionMismatch: {basis.m} elements, {a.shape} coefficients")
    return LazyArrowhead(basis, 0.0, 1.0, ctx, a)
