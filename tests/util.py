"""Shared test helpers: seeded random arrowheads of the shapes used by
/root/reference/test/test_arrowhead.jl, and oracle <-> product conversions."""
import numpy as np

from oracle import arrowhead_oracle as O


def random_general(rng, n, p, nB=2, nC=3, lower_B=True, lD=1, uD2=True, diag1=False, same_D=True):
    """`n` hats; lower_B: n-1 elements, B lower-bidiagonal (test_arrowhead.jl:9-15);
    else Dirichlet-type: n-2 hats, n-1 elements, B upper-bidiagonal (:137-146)."""
    m = n - 1
    nh = n if lower_B else n - 2
    A = np.diag(rng.standard_normal(nh) + 10) + np.diag(rng.standard_normal(nh - 1), 1) + np.diag(rng.standard_normal(nh - 1), -1)

    def Bblk():
        Bm = np.zeros((nh, m))
        if lower_B:
            Bm[np.arange(m), np.arange(m)] = rng.standard_normal(m)
            Bm[np.arange(m) + 1, np.arange(m)] = rng.standard_normal(m)
        else:
            Bm[np.arange(nh), np.arange(nh)] = rng.standard_normal(nh)
            Bm[np.arange(nh), np.arange(nh) + 1] = rng.standard_normal(nh)
        return Bm

    B = [Bblk() for _ in range(nB)]
    Cb = [Bblk().T.copy() for _ in range(nC)]

    def Dblk():
        Dk = np.diag(rng.standard_normal(p) + 10)
        if uD2 and p > 2:
            Dk += np.diag(rng.standard_normal(p - 2), 2)
        if lD >= 1 and p > 1:
            Dk += np.diag(rng.standard_normal(p - 1), -1)
        if diag1 and p > 1:
            Dk += np.diag(rng.standard_normal(p - 1), 1)
        return Dk

    if same_D:
        D0 = Dblk()
        D = [D0.copy() for _ in range(m)]
    else:
        D = [Dblk() for _ in range(m)]
    return O.BlockArrowhead(A, B, Cb, D)


def to_product(P, A: O.BlockArrowhead):
    """oracle block form -> product BBBArrowheadMatrix (through the reference-style constructor)."""
    return P.BBBArrowheadMatrix(A.A, tuple(A.B), tuple(A.C), A.D)


def packed_to_product(P, pk: O.Packed, uniform=False):
    D = pk.D[:, :, :1].copy() if uniform else pk.D
    return P.BBBArrowheadMatrix.from_packed(pk.m, pk.nh, pk.q, pk.lD, pk.uD, pk.Ad, pk.Au, pk.Al, pk.B, pk.C, D, uniform=uniform)


def product_to_packed(A) -> O.Packed:
    p = A.packed()
    return O.Packed(p["m"], p["nh"], p["q"], p["lD"], p["uD"], p["Ad"], p["Au"], p["Al"], p["B"], p["C"], p["D"])


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
