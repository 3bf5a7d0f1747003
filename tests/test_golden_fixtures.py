"""Committed golden fixtures (tests/golden/, generator: tests/golden/make_golden.py).

CPU part: the structured oracle (the restatement of src/arrowhead.jl the GPU is checked against)
reproduces the dense-LAPACK fixtures.  GPU part (-m gpu): the CUDA path through the C ABI reproduces
them too.  Tolerance: relative difference <= 1e-12 (north star), written at each assert."""
import json
import os

import numpy as np
import pytest

from oracle import arrowhead_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "arrowhead_golden.npz"))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))
TOL = 1e-12
SOLVE_CASES = [k for k in META["cases"] if not k.startswith("adi")]


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def _packed(case):
    c = META["cases"][case]
    basis = O.C1 if c["basis"] == "c1" else O.DIRICHLET
    h = 2.0 / c["m"]
    Mp, Lp = O.assemble_mass(basis, c["m"], c["N"], h), O.assemble_weaklaplacian(basis, c["m"], c["N"], h)
    return Mp, Lp, O.packed_axpby(-1.0, Lp, c["sigma"], Mp)


def test_reference_goldens_recorded():
    r = META["reference_goldens"]
    assert r["helmholtz_value_at_0.1"] == 1.1952730862177243          # test/test_arrowhead.jl:134
    assert r["bandwidths"] == [13, 9] and r["blockbandwidths"] == [3, 2]


@pytest.mark.parametrize("case", SOLVE_CASES)
def test_oracle_reproduces_fixtures(case):
    Mp, Lp, Sp = _packed(case)
    Up = O.packed_revchol(Sp)
    assert relerr(O.unpack(Up).to_dense(), G[f"{case}/U"]) < TOL                      # factor
    assert relerr(O.packed_solve(Up, G[f"{case}/B"].copy()), G[f"{case}/X"]) < TOL    # F \ B
    Y = np.zeros_like(G[f"{case}/B"])
    O.packed_mul_left(1.0, O.packed_symmetrize(Mp), G[f"{case}/B"], 0.0, Y)
    assert relerr(Y, G[f"{case}/MB"]) < TOL                                            # M * B


def test_oracle_adi_fixture():
    c = META["cases"]["adi_dirichlet_m4_N8"]
    h = 2.0 / c["m"]
    Mp = O.assemble_mass(O.DIRICHLET, c["m"], c["N"], h)
    Kp = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.DIRICHLET, c["m"], c["N"], h), 0.0, Mp)
    X = O.adi_packed(Mp, Kp, G["adi_dirichlet_m4_N8/F"], c["a"], c["b"], -c["b"], -c["a"], tol=1e-15)
    assert relerr(X, G["adi_dirichlet_m4_N8/X"]) < 1e-11       # converged iterate; cond(M) amplifies the last bits


# --------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the same fixtures
# --------------------------------------------------------------------------------------------------
def _gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pop_b200 as P
    return torch, P


def _dev(torch, x):
    rows, cols = x.shape
    ld = (rows + 1) & ~1
    buf = torch.zeros((cols, ld), dtype=torch.float64, device="cuda")
    buf[:, :rows] = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    return buf.t()[:rows]


@pytest.mark.gpu
@pytest.mark.parametrize("case", SOLVE_CASES)
def test_gpu_reproduces_fixtures(case):
    torch, P = _gpu()
    c = META["cases"][case]
    pts = np.linspace(-1, 1, c["m"] + 1)
    basis = P.ContinuousPolynomial(1, pts) if c["basis"] == "c1" else P.DirichletPolynomial(pts)
    KR = P.Block.oneto(c["N"])
    M = P.grammatrix(basis)[KR, KR]
    S = (-P.weaklaplacian(basis) + c["sigma"] * P.grammatrix(basis))[KR, KR]
    F = P.reversecholesky(S)
    assert relerr(F.U.to_dense(), G[f"{case}/U"]) < TOL
    B = G[f"{case}/B"]
    for algo in (P.SOLVE_AUTO, P.SOLVE_FUSED_SMEM, P.SOLVE_STAGED):
        X = P.ldiv(F, _dev(torch, B), algo=algo).cpu().numpy()
        assert relerr(X, G[f"{case}/X"]) < TOL, algo
    assert relerr((M @ _dev(torch, B)).cpu().numpy(), G[f"{case}/MB"]) < TOL
    assert relerr((_dev(torch, np.ascontiguousarray(B.T)) @ M).cpu().numpy(), G[f"{case}/BtM"]) < TOL     # X * A


@pytest.mark.gpu
def test_gpu_adi_fixture():
    torch, P = _gpu()
    c = META["cases"]["adi_dirichlet_m4_N8"]
    Q = P.DirichletPolynomial(np.linspace(-1, 1, c["m"] + 1))
    KR = P.Block.oneto(c["N"])
    M = P.grammatrix(Q)[KR, KR]
    K = -P.weaklaplacian(Q)[KR, KR]
    X = P.adi_poisson(M, K, _dev(torch, G["adi_dirichlet_m4_N8/F"]), c["a"], c["b"], tol=1e-15)
    assert relerr(X.cpu().numpy(), G["adi_dirichlet_m4_N8/X"]) < 1e-11
