"""ctypes wrapper of oracle/pop_oracle.c (CPU oracle at full size; TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

from .arrowhead_oracle import Packed

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpop_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_revchol.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n: int):
    """OpenMP thread count of the C restatement (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def transpose(A: np.ndarray) -> np.ndarray:
    """A' of a Fortran-ordered matrix, Fortran-ordered."""
    assert A.dtype == np.float64 and A.ndim == 2 and A.flags.f_contiguous
    B = np.empty((A.shape[1], A.shape[0]), order="F")
    lib().oracle_transpose(_p(A), C.c_int64(A.shape[0]), C.c_int64(A.shape[0]), C.c_int64(A.shape[1]), _p(B), C.c_int64(B.shape[0]))
    return B


def _upper(P: Packed):
    """upper data with per-element D: (Ad, Au, B, D[uD+1][q][m]) as fresh contiguous arrays"""
    D = np.ascontiguousarray(np.broadcast_to(P.D[P.lD:], (P.uD + 1, P.q, P.m)), dtype=np.float64).copy()
    return (np.ascontiguousarray(P.Ad, dtype=np.float64).copy(), np.ascontiguousarray(P.Au, dtype=np.float64).copy(),
            np.ascontiguousarray(P.B, dtype=np.float64).copy(), D)


def revchol(P: Packed) -> Packed:
    Ad, Au, B, D = _upper(P)
    info = lib().oracle_revchol(P.m, P.nh, P.q, P.nB, P.uD, _p(Ad), _p(Au), _p(B), _p(D))
    if info:
        raise np.linalg.LinAlgError(f"non-positive pivot {info}")
    return Packed(P.m, P.nh, P.q, 0, P.uD, Ad, Au, np.zeros_like(Au), B, np.zeros((0, 3, P.m)), D)


def solve(U: Packed, X: np.ndarray):
    """In place on a Fortran-ordered (n x nrhs) array or a vector."""
    assert X.dtype == np.float64 and (X.ndim == 1 or X.flags.f_contiguous)
    Ad, Au, B, D = _upper(U)
    nrhs = 1 if X.ndim == 1 else X.shape[1]
    lib().oracle_solve(U.m, U.nh, U.q, U.nB, U.uD, _p(Ad), _p(Au), _p(B), _p(D), _p(X), C.c_int64(X.shape[0]), C.c_int64(nrhs))
    return X


def solve_refloop(U: Packed, X: np.ndarray):
    """The same solve with the reference's loop structure (column by column, elements threaded in the upper solve, serial in
    the lower one; src/arrowhead.jl:439,480): a timing context for the favourable column-parallel port, same arithmetic."""
    assert X.dtype == np.float64 and (X.ndim == 1 or X.flags.f_contiguous)
    Ad, Au, B, D = _upper(U)
    nrhs = 1 if X.ndim == 1 else X.shape[1]
    lib().oracle_solve_refloop(U.m, U.nh, U.q, U.nB, U.uD, _p(Ad), _p(Au), _p(B), _p(D), _p(X), C.c_int64(X.shape[0]), C.c_int64(nrhs))
    return X


def mul_sym(P: Packed, X: np.ndarray, alpha=1.0, beta=0.0, Y=None):
    assert X.dtype == np.float64 and (X.ndim == 1 or X.flags.f_contiguous)
    Ad, Au, B, D = _upper(P)
    if Y is None:
        Y = np.zeros_like(X, order="F")
    nrhs = 1 if X.ndim == 1 else X.shape[1]
    lib().oracle_mul_sym(P.m, P.nh, P.q, P.nB, P.uD, _p(Ad), _p(Au), _p(B), _p(D), C.c_double(alpha), _p(X), C.c_int64(X.shape[0]),
                         C.c_double(beta), _p(Y), C.c_int64(Y.shape[0]), C.c_int64(nrhs))
    return Y


def adi(Mp: Packed, Kp: Packed, F: np.ndarray, a, b, c, d, J: int, iterations=None, X0=None):
    """The reference ADI iteration (examples/adi_poisson.jl:42-59, A = M, B = -M, C = K) on the C kernels:
    per iteration 2 shifted factorisations, 2 products, 2 solves; X / F as (F \\ X')' (explicit transposes).
    Runs `iterations` (default all J) iterations of the J-shift schedule and returns X."""
    from .arrowhead_oracle import adi_shifts, packed_axpby
    n = Mp.n
    p, qq = adi_shifts(J, a, b, c, d)
    X = np.zeros((n, n), order="F") if X0 is None else np.asfortranarray(X0)
    Ft = transpose(np.asfortranarray(F))
    for j in range(J if iterations is None else iterations):
        # X = ((M/p - K) X - F/p) / revchol(K + M/p)
        R = np.array(F, order="F", copy=True)
        mul_sym(packed_axpby(1.0 / p[j], Mp, -1.0, Kp), X, 1.0, -1.0 / p[j], R)
        Rt = transpose(R)
        solve(revchol(packed_axpby(1.0, Kp, 1.0 / p[j], Mp)), Rt)
        # X = revchol(K - M/q) \\ (X (-M/q - K) - F/q);  (X T)' = T X' for symmetric T
        Wt = np.array(Ft, order="F", copy=True)
        mul_sym(packed_axpby(-1.0 / qq[j], Mp, -1.0, Kp), Rt, 1.0, -1.0 / qq[j], Wt)
        X = transpose(Wt)
        solve(revchol(packed_axpby(1.0, Kp, -1.0 / qq[j], Mp)), X)
    return X
