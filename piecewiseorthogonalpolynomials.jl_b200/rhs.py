"""Right-hand-side pipeline of examples/adi_poisson.jl on the device.

  legendre_grid_transform : nodes + analysis matrix of one cell, what `plan_grid_transform(legendre(0..dx), (p,p))`
                            (examples/adi_poisson.jl:64) provides (host arithmetic, O(p^3) once)
  analysis_2d             : `analysis_2D` (examples/adi_poisson.jl:60-76): values on the cell grids -> tensor Legendre
                            coefficients in P (x) P, P = ContinuousPolynomial{0}  (pop_legendre_analysis_2d)
  weakform_apply / weakform_rhs : (Q'P)[KR,KR] X, X (Q'P)[KR,KR]' and F = (Q'P) fp (Q'P)'  (examples/adi_poisson.jl:105)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .arrowhead import _Panel, _bind_stream, _new_like
from .bases import BlockRange, _Basis


def legendre_grid_transform(p: int, kind: str = "gauss"):
    """(z, T): p nodes on the reference cell [-1,1] and the analysis matrix (Legendre coefficients = T @ values),
    T = V^-1 with V[s,a] = P_a(z_s).  `gauss`: Gauss-Legendre nodes (T = diag((2a+1)/2) V' diag(w), no inversion);
    `chebyshev1`: first-kind Chebyshev points."""
    if kind == "gauss":
        z, w = np.polynomial.legendre.leggauss(p)
        V = np.polynomial.legendre.legvander(z, p - 1)
        T = ((2 * np.arange(p) + 1) / 2)[:, None] * V.T * w[None, :]
    elif kind == "chebyshev1":
        z = np.cos(np.pi * (2 * np.arange(p)[::-1] + 1) / (2 * p))
        T = np.linalg.inv(np.polynomial.legendre.legvander(z, p - 1))
    else:
        raise ValueError(f"unknown node family {kind!r}")
    return z, np.asfortranarray(T)


def cell_grid(points, z):
    """All nodes of a uniform mesh, cell by cell: x[i*p + s] = node s of cell i (what `f_.(z, z')` is evaluated on)."""
    pts = np.asarray(points, dtype=np.float64)
    a, b = pts[:-1], pts[1:]
    return ((a + b)[:, None] / 2 + (b - a)[:, None] / 2 * np.asarray(z)[None, :]).reshape(-1)


def analysis_2d(vals, n: int, p: int, T, ctx=None):
    """vals: CUDA torch tensor, (n p) x (n p) column-major, vals[i p + s, j p + t] = f(x_{i,s}, y_{j,t}).
    Returns F with F[i + n a, j + n b] = tensor Legendre coefficient (a, b) of cell (i, j)."""
    pv = _Panel(vals, False)
    if not pv.torch:
        raise TypeError("analysis_2d works on device-resident (CUDA torch) values")
    if pv.rows != n * p or pv.cols != n * p:
        raise L.DimensionMismatch(-2, f"DimensionMismatch: values are {pv.rows}x{pv.cols}, expected {n * p}x{n * p}")
    T = np.asfortranarray(np.asarray(T, dtype=np.float64))
    if T.shape != (p, p):
        raise L.DimensionMismatch(-2, f"DimensionMismatch: analysis matrix is {T.shape}, expected {(p, p)}")
    ctx = ctx or L.default_context()
    F = _new_like(vals, (n * p, n * p))
    pf = _Panel(F, True)
    _bind_stream(ctx)
    L.check(L.lib().pop_legendre_analysis_2d(ctx.handle, int(n), int(p), T.ctypes.data_as(C.c_void_p), pv.ptr, pv.ld, pf.ptr, pf.ld), "analysis_2D")
    return F


def weakform_apply(Q: _Basis, KR: BlockRange, X, side=L.LEFT, ctx=None):
    """(Q'P)[KR,KR] @ X (side LEFT, X has m N rows) or X @ (Q'P)[KR,KR]' (side RIGHT, X has m N columns)."""
    N = KR.N
    px = _Panel(X, False)
    if not px.torch:
        raise TypeError("weakform_apply works on device-resident (CUDA torch) panels")
    n = Q.nh + Q.m * (N - 1)
    nP = Q.m * N
    if (px.rows if side == L.LEFT else px.cols) != nP:
        raise L.DimensionMismatch(-2, f"DimensionMismatch: (Q'P) is {n}x{nP}, panel is {px.rows}x{px.cols}")
    cnt = px.cols if side == L.LEFT else px.rows
    Y = _new_like(X, (n, cnt) if side == L.LEFT else (cnt, n))
    py = _Panel(Y, True)
    ctx = ctx or L.default_context()
    _bind_stream(ctx)
    hk = None if Q.uniform else Q.hk.ctypes.data_as(C.c_void_p)
    L.check(L.lib().pop_weakform_apply(ctx.handle, Q.basis_id, Q.m, int(N), float(Q.h), hk, side, px.ptr, px.ld, py.ptr, py.ld, cnt), "weak form")
    return Y


def weakform_rhs(Q: _Basis, KR: BlockRange, fp, ctx=None):
    """F = (Q'P)[KR,KR] * fp * (Q'P)[KR,KR]'  (examples/adi_poisson.jl:105)."""
    return weakform_apply(Q, KR, weakform_apply(Q, KR, fp, L.LEFT, ctx), L.RIGHT, ctx)


class DiscontinuityError(ValueError):
    """ArgumentError("Discontinuity in data on order of ...") of src/continuouspolynomial.jl:105-107."""


def legendre_to_c1(Q: _Basis, N: int, X, side=L.LEFT, check=True, ctx=None):
    """Legendre coefficients (N degrees per element, P ordering; columns of X for side LEFT, rows for RIGHT) of a continuous
    piecewise polynomial -> coefficients in Q's blocks 1..N-1: the `R \\ dat` + interlacing step of
    ContinuousPolynomialTransform (src/continuouspolynomial.jl:84-146).  `check` reproduces the reference's continuity test."""
    px = _Panel(X, False)
    if not px.torch:
        raise TypeError("legendre_to_c1 works on device-resident (CUDA torch) panels")
    nP, n = Q.m * N, Q.nh + Q.m * (N - 2)
    if (px.rows if side == L.LEFT else px.cols) != nP:
        raise L.DimensionMismatch(-2, f"DimensionMismatch: {N} Legendre degrees on {Q.m} elements are {nP} entries, panel is {px.rows}x{px.cols}")
    cnt = px.cols if side == L.LEFT else px.rows
    Y = _new_like(X, (n, cnt) if side == L.LEFT else (cnt, n))
    py = _Panel(Y, True)
    ctx = ctx or L.default_context()
    _bind_stream(ctx)
    jump = C.c_double(0.0)
    L.check(L.lib().pop_legendre_to_c1(ctx.handle, Q.basis_id, Q.m, int(N), side, px.ptr, px.ld, py.ptr, py.ld, cnt,
                                       C.byref(jump) if check else None), "transform")
    if check:
        scale = max(1.0, float(X.abs().max()))
        if jump.value > 1000 * N * np.finfo(float).eps * scale:
            raise DiscontinuityError(f"Discontinuity in data on order of {jump.value}.")
    return Y


def transform_2d(Q: _Basis, vals, p: int, T, check=True, ctx=None):
    """plan_grid_transform(Q, Block(p-1, p-1)) * vals (src/continuouspolynomial.jl:148-160, test_continuouspolynomial.jl:29-33):
    values on the p x p tensor grid of every cell -> coefficients in Q (x) Q, blocks 1..p-1 per dimension."""
    fp = analysis_2d(vals, Q.m, p, T, ctx)
    return legendre_to_c1(Q, p, legendre_to_c1(Q, p, fp, L.LEFT, check, ctx), L.RIGHT, check, ctx)


def transform(Q: _Basis, f, p: int, kind="gauss", check=True, ctx=None):
    """transform(Q, f)[Block.(1:p-1)] for a function of one variable (examples/heat.jl:41-47): f on the p nodes of every
    cell, Legendre analysis per cell (O(m p^2), host), then the device conversion.  Returns a CUDA vector."""
    import torch
    z, T = legendre_grid_transform(p, kind)
    x = cell_grid(Q.points, z)
    vals = np.asarray([f(t) for t in x], dtype=np.float64).reshape(Q.m, p)           # [cell, node]
    leg = (vals @ T.T).T.reshape(-1)                                                 # degree a of cell k at a*m + k
    ctx = ctx or L.default_context()
    dev = torch.from_numpy(np.ascontiguousarray(leg)).to(f"cuda:{ctx.device}")
    return legendre_to_c1(Q, p, dev.reshape(-1, 1), L.LEFT, check, ctx)[:, 0]


# ---- inverse transforms (`Pl \\ cfs`, src/continuouspolynomial.jl:129-146,171-184; src/dirichlet.jl:97-110,143-163) -----------

def legendre_synthesis_matrix(z, p: int):
    """S[s, a] = P_a(z_s): values on the nodes z of the reference cell = S @ Legendre coefficients (the inverse of the
    analysis matrix of legendre_grid_transform)."""
    return np.asfortranarray(np.polynomial.legendre.legvander(np.asarray(z, dtype=np.float64), p - 1))


def c1_to_legendre(Q: _Basis, N: int, X, side=L.LEFT, ctx=None):
    """Coefficients in Q's blocks 1..N-1 (columns of X for side LEFT, rows for RIGHT) -> Legendre coefficients, N degrees per
    element in P ordering: the `Pl.R \\ dat` step of the inverse transform, i.e. the truncated conversion P \\ Q
    (src/continuouspolynomial.jl:226-235, src/dirichlet.jl:35-53).  Inverse of legendre_to_c1."""
    px = _Panel(X, False)
    if not px.torch:
        raise TypeError("c1_to_legendre works on device-resident (CUDA torch) panels")
    nP, n = Q.m * N, Q.nh + Q.m * (N - 2)
    if (px.rows if side == L.LEFT else px.cols) != n:
        raise L.DimensionMismatch(-2, f"DimensionMismatch: blocks 1..{N - 1} of the basis are {n} entries, panel is {px.rows}x{px.cols}")
    cnt = px.cols if side == L.LEFT else px.rows
    Y = _new_like(X, (nP, cnt) if side == L.LEFT else (cnt, nP))
    py = _Panel(Y, True)
    ctx = ctx or L.default_context()
    _bind_stream(ctx)
    L.check(L.lib().pop_c1_to_legendre(ctx.handle, Q.basis_id, Q.m, int(N), side, px.ptr, px.ld, py.ptr, py.ld, cnt), "inverse transform")
    return Y


def synthesis_2d(F, n: int, p: int, S, ctx=None):
    """Inverse of analysis_2d: F[i + n a, j + n b] (tensor Legendre coefficients, CUDA torch tensor) -> values
    vals[i p + s, j p + t] on the cell grids; S = legendre_synthesis_matrix(z, p)."""
    pf = _Panel(F, False)
    if not pf.torch:
        raise TypeError("synthesis_2d works on device-resident (CUDA torch) coefficients")
    if pf.rows != n * p or pf.cols != n * p:
        raise L.DimensionMismatch(-2, f"DimensionMismatch: coefficients are {pf.rows}x{pf.cols}, expected {n * p}x{n * p}")
    S = np.asfortranarray(np.asarray(S, dtype=np.float64))
    if S.shape != (p, p):
        raise L.DimensionMismatch(-2, f"DimensionMismatch: synthesis matrix is {S.shape}, expected {(p, p)}")
    ctx = ctx or L.default_context()
    V = _new_like(F, (n * p, n * p))
    pv = _Panel(V, True)
    _bind_stream(ctx)
    L.check(L.lib().pop_legendre_synthesis_2d(ctx.handle, int(n), int(p), S.ctypes.data_as(C.c_void_p), pf.ptr, pf.ld, pv.ptr, pv.ld), "synthesis_2D")
    return V


def itransform_2d(Q: _Basis, Cf, p: int, S, ctx=None):
    """plan_grid_transform(Q, Block(p-1, p-1)) \\ Cf (src/continuouspolynomial.jl:171-184, test_continuouspolynomial.jl:33
    `F \\ V ≈ vals`): coefficients in Q (x) Q -> values on the p x p tensor grid of every cell."""
    return synthesis_2d(c1_to_legendre(Q, p, c1_to_legendre(Q, p, Cf, L.LEFT, ctx), L.RIGHT, ctx), Q.m, p, S, ctx)


def itransform(Q: _Basis, c, p: int, kind="gauss", ctx=None):
    """plan_grid_transform(Q, Block(p-1)) \\ c for one vector of coefficients (src/continuouspolynomial.jl:129-146,
    test_continuouspolynomial.jl:19,26): the device conversion to Legendre coefficients, then the O(m p^2) synthesis on
    the p nodes of every cell on the host (as `transform` does the analysis).  Returns (x, values) as NumPy arrays."""
    z, _ = legendre_grid_transform(p, kind)
    leg = c1_to_legendre(Q, p, c.reshape(-1, 1), L.LEFT, ctx)[:, 0].cpu().numpy().reshape(p, Q.m)      # [degree, cell]
    vals = (legendre_synthesis_matrix(z, p) @ leg).T.reshape(-1)                                    # node s of cell k at k p + s
    return cell_grid(Q.points, z), vals
