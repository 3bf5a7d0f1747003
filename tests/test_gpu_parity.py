"""GPU parity tests: the CUDA path (through the C ABI, via the host mirror of the reference API)
against the CPU oracle on the same seeded inputs.  Tolerances: the north star asks for a relative
solution difference and residual <= 1e-12 in Float64; each assert states its bound.

Testsets of /root/reference/test/test_arrowhead.jl are restated one by one (line numbers cited)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import pop_b200 as P  # noqa: E402
from oracle import arrowhead_oracle as O  # noqa: E402
from tests.util import packed_to_product, product_to_packed, random_general, relerr, to_product  # noqa: E402

TOL = 1e-12


def dev(x):
    """NumPy (any order) -> column-major CUDA tensor with an even leading dimension."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return torch.from_numpy(x.copy()).cuda()
    rows, cols = x.shape
    ld = (rows + 1) & ~1
    buf = torch.zeros((cols, ld), dtype=torch.float64, device="cuda")
    buf[:, :rows] = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    return buf.t()[:rows]


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _fused_env():
    old = {k: os.environ.get(k) for k in ("POP_FUSED_CS", "POP_DISABLE_FUSED", "POP_F2_CS", "POP_F2_OCC", "POP_F2_SLOTS", "POP_DISABLE_FUSED2")}
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


# --------------------------------------------------------------------------------------
# test_arrowhead.jl testsets
# --------------------------------------------------------------------------------------
def test_constructor_and_dense_roundtrip():
    # :9-15
    rng = np.random.default_rng(0)
    for lower in (True, False):
        Ao = random_general(rng, 5, 5, nB=2, nC=3, lower_B=lower, same_D=False)
        A = to_product(P, Ao)
        assert isinstance(A, P.BBBArrowheadMatrix)
        assert A.shape == (Ao.n, Ao.n)
        assert np.array_equal(A.to_dense(), Ao.to_dense())
        assert A.blockaxes() == (Ao.nh,) + (Ao.m,) * Ao.q


def test_algebra():
    # :17-37  (exact equalities, as in the reference)
    rng = np.random.default_rng(1)
    Ao = random_general(rng, 4, 5)
    A = to_product(P, Ao)
    S = Ao.to_dense()
    assert np.array_equal((2 * A).to_dense(), 2 * S)
    assert np.array_equal((A * 2).to_dense(), 2 * S)
    assert np.array_equal((A + A).to_dense(), 2 * S)
    assert np.all((A - A).to_dense() == 0)
    assert np.array_equal((A / 2).to_dense(), S / 2)
    At = to_product(P, Ao.transpose())
    assert np.array_equal((A + At).to_dense(), S + S.T)
    assert np.array_equal((A - At).to_dense(), S - S.T)
    assert np.array_equal((-A).to_dense(), -S)
    # ragged B tuples (tupleop, :392-398): Laplacian (no B) + mass (two B blocks)
    C = P.ContinuousPolynomial(1, np.linspace(-1, 1, 4))
    KR = P.Block.oneto(6)
    M = P.grammatrix(C)[KR, KR]
    Lp = P.weaklaplacian(C)[KR, KR]
    Hm = (-Lp + M).to_dense()
    np.testing.assert_array_equal(Hm, -Lp.to_dense() + M.to_dense())


def test_mul():
    # :39-54
    rng = np.random.default_rng(2)
    Ao = random_general(rng, 4, 5)
    A = to_product(P, Ao)
    S = Ao.to_dense()
    n = Ao.n
    x = rng.standard_normal(n)
    X = rng.standard_normal((n, n))
    Xt = rng.standard_normal((n, 5))
    for wrap in (dev, np.asfortranarray):          # device pointers and host buffers
        get = host if wrap is dev else np.asarray
        assert relerr(get(A @ wrap(x)), S @ x) < TOL                 # A*x
        assert relerr(get(A.T @ wrap(x)), S.T @ x) < TOL             # A'*x
        assert relerr(get(A @ wrap(X)), S @ X) < TOL                 # A*X
        assert relerr(get(A @ wrap(Xt)), S @ Xt) < TOL               # A*X~
        assert relerr(get(wrap(X) @ A), X @ S) < TOL                 # X*A
    Y0 = rng.standard_normal((n, 5))
    out = dev(Y0)
    P.mul(A, dev(Xt), alpha=0.7, beta=-1.3, out=out)                 # mul!(Y, A, X, alpha, beta)
    assert relerr(host(out), 0.7 * S @ Xt - 1.3 * Y0) < TOL
    Sy = P.Symmetric(A)
    assert relerr(host(Sy @ dev(Xt)), Sy.to_dense() @ Xt) < TOL
    assert relerr(host(dev(X) @ Sy), X @ Sy.to_dense()) < TOL
    with pytest.raises(P.DimensionMismatch):
        A @ dev(rng.standard_normal(n + 1))


@pytest.mark.parametrize("lower_B", [True, False])
def test_triangular(lower_B):
    # :56-81
    rng = np.random.default_rng(3)
    Ao = random_general(rng, 5, 5, nB=2, nC=2, lower_B=lower_B, same_D=False)
    A = to_product(P, Ao)
    n = Ao.n
    c = rng.standard_normal(n)
    Cm = rng.standard_normal((n, 7))
    R = rng.standard_normal((6, n))
    for T in (P.UpperTriangular(A), P.UnitUpperTriangular(A), P.LowerTriangular(A), P.UnitLowerTriangular(A), P.UpperTriangular(A).T):
        Td = T.to_dense()
        assert relerr(host(P.ldiv(T, dev(c))), np.linalg.solve(Td, c)) < 1e-11        # T \ c
        assert relerr(host(P.rdiv(dev(c), T)), np.linalg.solve(Td.T, c)) < 1e-11      # c' / T
        assert relerr(host(P.ldiv(T, dev(Cm))), np.linalg.solve(Td, Cm)) < 1e-11
        assert relerr(host(P.rdiv(dev(R), T)), np.linalg.solve(Td.T, R.T).T) < 1e-11  # X / T
        assert relerr(P.ldiv(T, np.asfortranarray(Cm)), np.linalg.solve(Td, Cm)) < 1e-11


@pytest.mark.parametrize("diag1", [False, True])
def test_reversecholesky_lower_B(diag1):
    # :83-100
    rng = np.random.default_rng(4)
    Ao = random_general(rng, 3, 5, nB=2, nC=0, lD=0, diag1=diag1)
    A = to_product(P, Ao)
    Uw = O.dense_reversecholesky(O.symmetric_from_upper(Ao).to_dense())
    F = P.reversecholesky(P.Symmetric(A))
    assert relerr(F.U.to_dense(), Uw) < TOL
    F2 = P.reversecholesky_inplace(P.Symmetric(A.copy()))
    assert relerr(F2.U.to_dense(), Uw) < TOL
    assert relerr(F.L.to_dense(), Uw.T) < TOL


def test_reversecholesky_dirichlet_upper_B():
    # :137-146
    rng = np.random.default_rng(5)
    Ao = random_general(rng, 5, 5, nB=2, nC=0, lower_B=False, lD=0)
    A = to_product(P, Ao)
    Ssym = O.symmetric_from_upper(Ao).to_dense()
    U = P.reversecholesky(P.Symmetric(A)).U.to_dense()
    assert relerr(U, O.dense_reversecholesky(Ssym)) < TOL
    assert relerr(U @ U.T, Ssym) < TOL


def test_helmholtz_reference_golden():
    # :117-135 -- the reference's known-answer value
    r = np.linspace(-1, 1, 4)
    C = P.ContinuousPolynomial(1, r)
    Delta = P.weaklaplacian(C)
    M = P.grammatrix(C)
    KR = P.Block.oneto(100)
    F = P.reversecholesky((-Delta + M)[KR, KR])
    cexp = O.c1_expand(np.exp, np.exp, r, 100)            # transform(C, exp)[KR] (out-of-scope transform: oracle)
    x = M[KR, KR] @ cexp
    c = P.ldiv(F, x)
    assert abs(O.c1_evaluate(c, r, 100, 0.1) - 1.1952730862177243) < 5e-13
    cd = host(P.ldiv(F, dev(x)))
    assert relerr(cd, c) < TOL


def test_not_positive_definite_and_structure_errors():
    rng = np.random.default_rng(6)
    Ao = random_general(rng, 4, 5, nB=2, nC=0, lD=0)
    Ao.D[1][2, 2] = -5.0
    with pytest.raises(P.PosDefException) as ei:
        P.reversecholesky(P.Symmetric(to_product(P, Ao)))
    assert ei.value.code == Ao.nh + 2 * Ao.m + 1 + 1         # 1-based global index of the bad pivot
    Ao = random_general(rng, 4, 5, nB=2, nC=0, lD=0)
    Ao.A[1, 1] = -100.0
    with pytest.raises(P.PosDefException) as ei:
        P.reversecholesky(P.Symmetric(to_product(P, Ao)))
    assert ei.value.code == 2
    F = P.reversecholesky(P.Symmetric(to_product(P, random_general(rng, 4, 5, nB=2, nC=0, lD=0))))
    with pytest.raises(P.DimensionMismatch):
        P.ldiv(F, dev(np.zeros(7)))


# --------------------------------------------------------------------------------------
# assembly + factor + solve against the oracle (staged and fused kernels, all cluster sizes)
# --------------------------------------------------------------------------------------
def _assemble(basis, m, N):
    h = 2.0 / m
    if basis == "c1":
        B = P.ContinuousPolynomial(1, np.linspace(-1, 1, m + 1))
        bo = O.C1
    else:
        B = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
        bo = O.DIRICHLET
    KR = P.Block.oneto(N)
    M = P.grammatrix(B)[KR, KR]
    Lp = P.weaklaplacian(B)[KR, KR]
    return M, Lp, O.assemble_mass(bo, m, N, h), O.assemble_weaklaplacian(bo, m, N, h)


@pytest.mark.parametrize("basis,m,N", [("c1", 3, 10), ("dirichlet", 4, 12), ("c1", 7, 5), ("dirichlet", 9, 3), ("c1", 5, 2), ("c1", 4, 1)])
def test_assembly_matches_oracle(basis, m, N):
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    np.testing.assert_allclose(M.data.to_dense(), O.unpack(Mo).to_dense(), rtol=2e-16, atol=0)
    np.testing.assert_allclose(Lp.data.to_dense(), O.unpack(Lo).to_dense(), rtol=2e-16, atol=0)


def test_readme_case_goldens():
    """BASELINE config 1: README case, survey goldens."""
    M, Lp, Mo, Lo = _assemble("c1", 3, 10)
    F = P.reversecholesky(-Lp + M)
    Lf = F.L.to_dense()
    for (i, j), v in {(1, 1): 0.9818490617583832, (2, 1): -0.8671084871570119, (5, 5): 1.4452827020171803, (8, 1): -0.020179587355387243,
                      (11, 5): -0.006841062695774555, (31, 25): -0.0002447828624173438, (31, 31): 0.5621263590512574}.items():
        assert abs(Lf[i - 1, j - 1] - v) < 1e-14
    b = np.arange(1, 32) / 31
    for algo in (P.SOLVE_STAGED, P.SOLVE_AUTO):
        x = host(P.ldiv(F, dev(b), algo=algo))
        for i, v in {1: 0.11098908536056581, 4: 0.18024989943421074, 5: 0.06638120377479541, 31: 3.1655746467356325}.items():
            assert abs(x[i - 1] - v) < 1e-13 * max(1, abs(v))
    Uo = O.packed_revchol(O.packed_axpby(-1.0, Lo, 1.0, Mo))
    assert relerr(F.U.to_dense(), O.unpack(Uo).to_dense()) < TOL


SHAPES = [("c1", 16, 12), ("dirichlet", 8, 9), ("dirichlet", 16, 24), ("c1", 13, 7), ("dirichlet", 11, 6), ("c1", 64, 40), ("dirichlet", 32, 64),
          ("c1", 2, 3), ("dirichlet", 2, 4), ("c1", 1, 6),
          ("dirichlet", 12, 100), ("c1", 6, 129), ("c1", 4, 140), ("dirichlet", 70, 17), ("c1", 33, 65)]


@pytest.mark.parametrize("basis,m,N", SHAPES)
@pytest.mark.parametrize("sigma", [1.0, 1e4])
def test_solve_staged_and_fused_vs_oracle(basis, m, N, sigma):
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    S = -Lp + sigma * M
    So = O.packed_axpby(-1.0, Lo, sigma, Mo)
    Uo = O.packed_revchol(So)
    F = P.reversecholesky(S)
    assert relerr(F.U.to_dense(), O.unpack(Uo).to_dense()) < TOL
    n = So.n
    rng = np.random.default_rng(7)
    Bm = rng.standard_normal((n, 37))
    want = O.packed_solve(Uo, Bm.copy())
    os.environ.pop("POP_FUSED_CS", None)
    os.environ.pop("POP_F2_CS", None)
    got = {"staged": host(P.ldiv(F, dev(Bm), algo=P.SOLVE_STAGED))}
    for cs in (1, 2, 4, 8):
        if cs > m:
            continue
        if (cs - 1) * -(-m // cs) >= m:
            continue                  # an empty trailing rank: not a valid partition
        os.environ["POP_FUSED_CS"] = str(cs)
        got[f"fused-smem cs={cs}"] = host(P.ldiv(F, dev(Bm), algo=P.SOLVE_FUSED_SMEM))
        os.environ["POP_F2_CS"] = str(cs)
        got[f"fused-reg cs={cs}"] = host(P.ldiv(F, dev(Bm), algo=P.SOLVE_FUSED))
    os.environ.pop("POP_FUSED_CS", None)
    os.environ.pop("POP_F2_CS", None)
    got["auto"] = host(P.ldiv(F, dev(Bm)))
    got["host"] = P.ldiv(F, np.asfortranarray(Bm))
    got["vector"] = host(P.ldiv(F, dev(Bm[:, 0])))[:, None]
    Sd = O.unpack(O.packed_symmetrize(So)).to_dense()
    for name, x in got.items():
        ref = want if name != "vector" else want[:, :1]
        rhs = Bm if name != "vector" else Bm[:, :1]
        assert relerr(x, ref) < TOL, name
        assert np.linalg.norm(Sd @ x - rhs) / (np.linalg.norm(Sd, 2) * np.linalg.norm(x)) < TOL, name
    # X / F on rows (examples/adi_poisson.jl:54)
    R = rng.standard_normal((9, n))
    assert relerr(host(P.rdiv(dev(R), F)), O.packed_solve(Uo, R.T.copy()).T) < TOL


def test_solve_per_element_blocks_general_bands():
    """Distinct D_k per element and a non-zero first superdiagonal (test_arrowhead.jl:96)."""
    rng = np.random.default_rng(8)
    for lower in (True, False):
        Ao = random_general(rng, 19, 11, nB=2, nC=0, lower_B=lower, lD=0, diag1=True, same_D=False)
        for k in range(Ao.m):      # make S comfortably positive definite
            Ao.D[k] += 5 * np.eye(Ao.q)
        A = to_product(P, Ao)
        F = P.reversecholesky(P.Symmetric(A))
        Ssym = O.symmetric_from_upper(Ao).to_dense()
        assert relerr(F.U.to_dense(), O.dense_reversecholesky(Ssym)) < TOL
        Bm = rng.standard_normal((Ao.n, 21))
        want = np.linalg.solve(Ssym, Bm)
        for algo in (P.SOLVE_STAGED, P.SOLVE_FUSED, P.SOLVE_FUSED_SMEM):
            for cs in ("1", "2"):
                os.environ["POP_FUSED_CS"] = cs
                assert relerr(host(P.ldiv(F, dev(Bm), algo=algo)), want) < 1e-11


def test_shifted_batch_factor_and_solve():
    """BASELINE config 3 in miniature: Dirichlet, many shifts sigma, batched factor + solve."""
    m, N = 16, 16
    M, Lp, Mo, Lo = _assemble("dirichlet", m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    sig = np.logspace(np.log10(np.pi ** 2 / 4), 8, 12)
    facs = P.reversecholesky_shifted(K, M, np.ones_like(sig), sig)
    assert len(facs) == len(sig)
    rng = np.random.default_rng(9)
    Bm = rng.standard_normal((Ko.n, 16))
    for s, F in zip(sig, facs):
        Uo = O.packed_revchol(O.packed_axpby(1.0, Ko, s, Mo))
        assert relerr(F.U.to_dense(), O.unpack(Uo).to_dense()) < TOL
        assert relerr(host(P.ldiv(F, dev(Bm))), O.packed_solve(Uo, Bm.copy())) < TOL
    with pytest.raises(P.PosDefException):
        P.reversecholesky_shifted(K, M, np.ones(3), np.array([1.0, -1e9, 2.0]))


@pytest.mark.parametrize("basis,m,N", [("dirichlet", 8, 12), ("c1", 16, 9), ("dirichlet", 32, 20)])
def test_fused_solve_mul_axpy(basis, m, N):
    """One ADI half step R <- T (S^{-1} R) + gamma F (examples/adi_poisson.jl:54-55), fused and staged."""
    import ctypes as C
    from pop_b200 import _lib as L
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    sigma, gamma = 3.7, -0.4
    S, T = -Lp + sigma * M, (0.5 * M + Lp)
    F = P.reversecholesky(S)
    n = S.n
    rng = np.random.default_rng(10)
    R0, Fm = rng.standard_normal((n, 19)), rng.standard_normal((n, 19))
    Uo = O.packed_revchol(O.packed_axpby(-1.0, Lo, sigma, Mo))
    To = O.packed_symmetrize(O.packed_axpby(0.5, Mo, 1.0, Lo))
    Xo = O.packed_solve(Uo, R0.copy())
    want = np.zeros_like(Xo)
    O.packed_mul_left(1.0, To, Xo, 0.0, want)
    want += gamma * Fm
    for disable2, disable in (("0", "0"), ("1", "0"), ("1", "1")):      # register kernel, shared-memory kernel, staged
        os.environ["POP_DISABLE_FUSED2"] = disable2
        os.environ["POP_DISABLE_FUSED"] = disable
        R, Fd = dev(R0), dev(Fm)
        L.check(L.lib().pop_solve_mul_axpy(F.ctx.handle, F.factor._h, T.data._h, gamma, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1), 19))
        torch.cuda.synchronize()
        assert relerr(host(R), want) < TOL, (disable2, disable)
    # a dense panel with an ODD leading dimension (ld = n, columns only 8-byte aligned): no TMA path applies, the call
    # must fall back to the staged kernels and still return the half step (never an error on a half-updated panel)
    if n % 2 == 1:
        os.environ["POP_DISABLE_FUSED2"] = "0"
        os.environ["POP_DISABLE_FUSED"] = "0"
        R = torch.from_numpy(np.ascontiguousarray(R0.T)).cuda().t()        # stride(1) == n, odd
        Fd = torch.from_numpy(np.ascontiguousarray(Fm.T)).cuda().t()
        assert R.stride(1) == n
        L.check(L.lib().pop_solve_mul_axpy(F.ctx.handle, F.factor._h, T.data._h, gamma, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1), 19))
        torch.cuda.synchronize()
        assert relerr(host(R), want) < TOL


def test_adi_small_vs_oracle():
    """examples/adi_poisson.jl:42-59 at m=4, N=12 (J=42)."""
    m, N = 4, 12
    M, Lp, Mo, Lo = _assemble("dirichlet", m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Mo.n
    Y = np.random.default_rng(1).standard_normal((n, n))
    Fm = Kd @ Y @ Md + Md @ Y @ Kd
    lam = np.linalg.eigvals(np.linalg.solve(Kd, Md)).real
    a, b = lam.min(), lam.max()
    X = host(P.adi_poisson(M, K, dev(Fm), a, b))
    Xo = O.adi_packed(Mo, Ko, Fm, a, b, -b, -a)
    assert np.abs(X - Xo).max() <= 1e-11 * np.abs(Xo).max()
    Yg = np.linalg.solve(Kd, X.T).T
    assert relerr(Yg, Y) < 1e-12
    res = Kd @ Yg @ Md + Md @ Yg @ Kd - Fm
    assert np.linalg.norm(res) / np.linalg.norm(Fm) < 1e-12
    # spectrum bounds by Lanczos on the device kernels enclose the dense ones (examples/adi_poisson.jl:91-94)
    a2, b2 = P.spectrum_bounds(M, K)
    assert a2 <= a and b2 >= b and a2 > 0.99 * a and b2 < 1.01 * b


@pytest.mark.parametrize("basis,m,N", [("dirichlet", 6, 10), ("c1", 5, 8), ("dirichlet", 9, 14)])
def test_element_decomposed_row_half_step_and_adi(basis, m, N):
    """csrc/pop_row.cu on one GPU (world = 1): the row half step R <- (R S^-1) T + gamma F against the oracle's
    dense algebra, and the transpose-free ADI against the oracle's ADI (examples/adi_poisson.jl:42-59)."""
    import ctypes as C
    from pop_b200 import _lib as L
    from pop_b200.arrowhead import _bind_stream
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Mo.n
    rng = np.random.default_rng(5)
    Rm, Fm = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    if basis == "c1":
        Kd = Kd + 0.3 * Md            # the pure Neumann Laplacian is singular: test the row kernels with a definite operator
        K = K + 0.3 * M
    Sd, Td = Kd + 3.7 * Md, 0.5 * Md - Kd
    want = np.linalg.solve(Sd, Rm.T).T @ Td - 0.4 * Fm
    dd = P.ElementDecomposition(M)
    assert np.array_equal(dd.columns, np.arange(n))           # one rank: the local panel is the matrix itself
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    R, Fd = dd.panel(), dd.panel()
    R.copy_(torch.from_numpy(Rm).cuda())
    Fd.copy_(torch.from_numpy(Fm).cuda())
    _bind_stream(M.data.ctx)
    L.check(L.lib().pop_dd_row_half_step(M.data.ctx.handle, dd._h, Fs.factor._h, T.data._h, -0.4, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1)))
    torch.cuda.synchronize()
    assert relerr(host(R), want) < TOL
    # plain X / F on rows (no product)
    R.copy_(torch.from_numpy(Rm).cuda())
    L.check(L.lib().pop_dd_row_half_step(M.data.ctx.handle, dd._h, Fs.factor._h, None, 0.0, None, 0, R.data_ptr(), R.stride(1)))
    torch.cuda.synchronize()
    assert relerr(host(R), np.linalg.solve(Sd, Rm.T).T) < TOL
    if basis == "dirichlet":
        Y = rng.standard_normal((n, n))
        Fa = Kd @ Y @ Md + Md @ Y @ Kd
        lam = np.linalg.eigvals(np.linalg.solve(Kd, Md)).real
        a, b = lam.min(), lam.max()
        Fd.copy_(torch.from_numpy(Fa).cuda())
        X = host(P.adi_poisson_dd(M, K, Fd, a, b, dd=dd))
        Xo = O.adi_packed(Mo, Ko, Fa, a, b, -b, -a)
        assert np.abs(X - Xo).max() <= 1e-11 * np.abs(Xo).max()
        Yg = np.linalg.solve(Kd, X.T).T
        res = Kd @ Yg @ Md + Md @ Yg @ Kd - Fa
        assert np.linalg.norm(res) / np.linalg.norm(Fa) < 1e-12
    dd.close()


@pytest.mark.parametrize("basis,m,N,nrhs", [("c1", 64, 40, 300), ("dirichlet", 65, 33, 257), ("c1", 7, 5, 3)])
def test_host_panels_flat_and_chunked(basis, m, N, nrhs, monkeypatch):
    """pop_solve_host / pop_trsm_host on host buffers: the flat address-aligned pipeline (contiguous panels), its fall-back
    to column chunks, several block sizes, a host base address that is only 8-byte aligned; all equal to the device path."""
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    F = P.reversecholesky(-Lp + 1.25 * M)
    n = Mo.n
    rng = np.random.default_rng(17)
    Bm = rng.standard_normal((n, nrhs))
    want = host(P.ldiv(F, dev(Bm)))
    for flat, mb in (("1", "1"), ("1", "32"), ("0", "1")):
        monkeypatch.setenv("POP_HOST_FLAT", flat)
        monkeypatch.setenv("POP_HOST_CHUNK_MB", mb)
        assert np.array_equal(P.ldiv(F, np.asfortranarray(Bm)), want), (flat, mb)
        # a panel that starts 8 bytes into an aligned allocation
        raw = np.zeros(n * nrhs + 1)
        Xo = raw[1:].reshape((n, nrhs), order="F")
        Xo[...] = Bm
        P.ldiv(F, Xo, overwrite=True)
        assert np.array_equal(Xo, want) and raw[0] == 0.0, (flat, mb)
        U = F.U
        assert relerr(P.ldiv(U, np.asfortranarray(Bm)), host(P.ldiv(U, dev(Bm)))) < 1e-15
        # Y = F \ X into another host panel (pop_solve_host_to): X untouched; also with a destination of another alignment
        Xin = np.asfortranarray(Bm)
        Yout = np.full((n, nrhs), np.nan, order="F")
        assert P.ldiv(F, Xin, out=Yout) is Yout
        assert np.array_equal(Yout, want) and np.array_equal(Xin, Bm), (flat, mb)
        raw2 = np.full(n * nrhs + 3, np.nan)
        Y2 = raw2[3:].reshape((n, nrhs), order="F")
        P.ldiv(F, Xin, out=Y2)
        assert np.array_equal(Y2, want) and np.isnan(raw2[:3]).all(), (flat, mb)
        Rin = np.asfortranarray(Bm.T.copy())                     # rows are the systems: X / F
        Rout = np.empty_like(Rin)
        P.rdiv(Rin, F, out=Rout)
        assert relerr(Rout, want.T) < 1e-14 and np.array_equal(Rin, Bm.T), (flat, mb)
    with pytest.raises(P.DimensionMismatch):
        P.ldiv(F, np.asfortranarray(Bm), out=np.empty((n + 1, nrhs), order="F"))


@pytest.mark.parametrize("basis", ["c1", "dirichlet"])
def test_nonuniform_mesh_assembly_factor_solve(basis):
    """grammatrix / weaklaplacian on a non-uniform mesh (pop_assemble_mesh, one D block per element),
    reverse Cholesky, F \\ X, X / F and the product against the oracle (closed forms verified by quadrature)."""
    pts = np.concatenate([[-1.0], -1.0 + 2.0 * np.sort(np.random.default_rng(3).random(10)), [1.0]])
    hk = np.diff(pts)
    m, N = hk.size, 12
    bo = O.C1 if basis == "c1" else O.DIRICHLET
    Q = P.ContinuousPolynomial(1, pts) if basis == "c1" else P.DirichletPolynomial(pts)
    assert not Q.uniform
    KR = P.Block.oneto(N)
    M, Lp = P.grammatrix(Q)[KR, KR], P.weaklaplacian(Q)[KR, KR]
    Mo, Lo = O.assemble_mass(bo, m, N, hk), O.assemble_weaklaplacian(bo, m, N, hk)
    Md, Ld = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Lo)).to_dense()
    assert relerr(M.to_dense(), Md) < 1e-15
    assert relerr(Lp.to_dense(), Ld) < 1e-15
    S = -Lp + 1.5 * M
    Sd = -Ld + 1.5 * Md
    F = P.reversecholesky(S)
    Uo = O.packed_revchol(O.packed_axpby(-1.0, Lo, 1.5, Mo))
    assert relerr(F.U.to_dense(), O.unpack(Uo).to_dense()) < TOL
    n = Mo.n
    rng = np.random.default_rng(9)
    Bm = rng.standard_normal((n, 33))
    want = O.packed_solve(Uo, Bm.copy())
    for algo in (P.SOLVE_AUTO, P.SOLVE_STAGED, P.SOLVE_FUSED_SMEM):
        X = host(P.ldiv(F, dev(Bm), algo=algo))
        assert relerr(X, want) < TOL, algo
        assert np.linalg.norm(Sd @ X - Bm) / (np.linalg.norm(Sd, 2) * np.linalg.norm(X)) < TOL
    R = rng.standard_normal((7, n))
    assert relerr(host(P.rdiv(dev(R), F)), np.linalg.solve(Sd, R.T).T) < TOL
    assert relerr(host(S @ dev(Bm)), Sd @ Bm) < TOL


@pytest.mark.parametrize("basis,m,N,nrows", [("c1", 3, 10, 1), ("dirichlet", 8, 12, 9), ("c1", 16, 40, 37), ("dirichlet", 32, 64, 50), ("dirichlet", 6, 80, 5)])
def test_rdiv_one_pass_rows(basis, m, N, nrows):
    """X / F (examples/adi_poisson.jl:54) on a row panel: pop_solve(side=RIGHT) runs the one-pass row kernel
    (csrc/pop_row1.cuh) when the chain fits (q <= 64), the staged row kernels otherwise; both against the oracle."""
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    S = -Lp + 2.5 * M
    So = O.packed_axpby(-1.0, Lo, 2.5, Mo)
    Uo = O.packed_revchol(So)
    F = P.reversecholesky(S)
    n = Mo.n
    rng = np.random.default_rng(21)
    R = rng.standard_normal((nrows, n))
    want = O.packed_solve(Uo, R.T.copy()).T
    auto = host(P.rdiv(dev(R), F))
    staged = host(P.rdiv(dev(R), F, algo=P.SOLVE_STAGED))
    assert relerr(auto, want) < TOL
    assert relerr(staged, want) < TOL
    assert relerr(auto, staged) < 1e-13
    if N - 1 <= 64:
        assert relerr(host(P.rdiv(dev(R), F, algo=P.SOLVE_FUSED)), want) < TOL    # must not fall back
    else:
        with pytest.raises(Exception):
            P.rdiv(dev(R), F, algo=P.SOLVE_FUSED)


@pytest.mark.parametrize("lower", [True, False])
def test_rdiv_one_pass_general_bands(lower, monkeypatch):
    """X / F with a shared D block whose FIRST super-diagonal is non-zero (test_arrowhead.jl:96): the one-pass row kernel
    keeps the whole chain on one thread (no even/odd split) and applies both bands."""
    rng = np.random.default_rng(31)
    Ao = random_general(rng, 14, 11, nB=2, nC=0, lower_B=lower, lD=0, diag1=True, same_D=True)
    for k in range(Ao.m):
        Ao.D[k] += 5 * np.eye(Ao.q)
    A = packed_to_product(P, O.pack(Ao), uniform=True)
    F = P.reversecholesky(P.Symmetric(A))
    Ssym = O.symmetric_from_upper(Ao).to_dense()
    R = rng.standard_normal((21, Ao.n))
    want = np.linalg.solve(Ssym, R.T).T
    monkeypatch.setenv("POP_ROW1_STRICT", "1")
    monkeypatch.setenv("POP_ROW2", "0")           # this test is about csrc/pop_row1.cuh
    for cs in ("1", "0"):
        monkeypatch.setenv("POP_ROW1_CS", cs)
        assert relerr(host(P.rdiv(dev(R), F, algo=P.SOLVE_FUSED)), want) < 1e-11
    assert relerr(host(P.rdiv(dev(R), F, algo=P.SOLVE_STAGED)), want) < 1e-11


@pytest.mark.parametrize("basis,m,N,cs,rb", [
    ("dirichlet", 8, 12, 2, 16), ("dirichlet", 8, 12, 4, 16), ("dirichlet", 8, 12, 8, 16), ("c1", 8, 9, 4, 16), ("c1", 8, 9, 8, 8),
    ("dirichlet", 16, 9, 8, 8), ("dirichlet", 12, 20, 4, 16), ("c1", 6, 21, 2, 8), ("dirichlet", 8, 40, 4, 16), ("c1", 4, 64, 2, 16),
    ("dirichlet", 8, 65, 8, 16), ("dirichlet", 5, 33, 1, 16)])
def test_one_pass_row_half_step(basis, m, N, cs, rb, monkeypatch):
    """csrc/pop_row1.cuh: the register-resident one-pass row half step (one GPU) for every chain class, cluster size
    and row-block height, against the oracle's dense algebra AND against the two-pass row kernels of pop_row.cu."""
    from pop_b200 import _lib as L
    from pop_b200.arrowhead import _bind_stream
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Mo.n
    rng = np.random.default_rng(11)
    Rm, Fm = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    if basis == "c1":
        Kd = Kd + 0.3 * Md
        K = K + 0.3 * M
    Sd, Td = Kd + 3.7 * Md, 0.5 * Md - Kd
    want = np.linalg.solve(Sd, Rm.T).T @ Td - 0.4 * Fm
    dd = P.ElementDecomposition(M)
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    R, Fd = dd.panel(), dd.panel()
    Fd.copy_(torch.from_numpy(Fm).cuda())
    _bind_stream(M.data.ctx)
    got = {}
    monkeypatch.setenv("POP_ROW2", "0")           # this test is about csrc/pop_row1.cuh (k_row2 has its own below)
    for mode in ("one-pass", "one-pass-direct-f", "one-pass-unsplit", "two-pass"):
        monkeypatch.setenv("POP_ROW1", "0" if mode == "two-pass" else "2")    # 2: also for the half step with the product
        monkeypatch.setenv("POP_ROW1_NPAR", "1" if mode == "one-pass-unsplit" else "2")   # 2: even/odd degree chains on two threads
        monkeypatch.setenv("POP_ROW1_CS", str(cs))
        monkeypatch.setenv("POP_ROW1_RB", str(rb))
        monkeypatch.setenv("POP_ROW1_STAGE_F", "0" if mode == "one-pass-direct-f" else "1")
        monkeypatch.setenv("POP_ROW1_STRICT", "1")    # the forced plan must apply: no silent fallback to the two-pass kernels
        R.copy_(torch.from_numpy(Rm).cuda())
        L.check(L.lib().pop_dd_row_half_step(M.data.ctx.handle, dd._h, Fs.factor._h, T.data._h, -0.4, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1)))
        torch.cuda.synchronize()
        got[mode] = host(R)
        assert relerr(got[mode], want) < TOL, mode
        R.copy_(torch.from_numpy(Rm).cuda())
        L.check(L.lib().pop_dd_row_half_step(M.data.ctx.handle, dd._h, Fs.factor._h, None, 0.0, None, 0, R.data_ptr(), R.stride(1)))
        torch.cuda.synchronize()
        assert relerr(host(R), np.linalg.solve(Sd, Rm.T).T) < TOL, mode
    assert relerr(got["one-pass"], got["two-pass"]) < 1e-13
    assert relerr(got["one-pass-unsplit"], got["two-pass"]) < 1e-13
    assert np.array_equal(got["one-pass"], got["one-pass-direct-f"])
    dd.close()


@pytest.mark.parametrize("basis,m,N,cs", [
    ("dirichlet", 8, 12, 1), ("dirichlet", 8, 12, 2), ("dirichlet", 8, 12, 4), ("c1", 8, 9, 4), ("c1", 8, 9, 1), ("dirichlet", 16, 9, 8),
    ("dirichlet", 12, 20, 2), ("c1", 6, 21, 1), ("dirichlet", 8, 40, 4), ("c1", 4, 64, 2), ("dirichlet", 8, 65, 4), ("dirichlet", 6, 33, 1),
    ("dirichlet", 32, 17, 1), ("c1", 32, 5, 2), ("dirichlet", 64, 7, 8), ("c1", 16, 3, 2)])
def test_single_pass_tma_row_half_step(basis, m, N, cs, monkeypatch):
    """csrc/pop_row2.cu (k_row2): the single-pass row half step with TMA tensor traffic and the cluster-wide hat broadcast,
    for every chain class and cluster size: half step with product and gamma F, product alone, plain X / F
    (examples/adi_poisson.jl:54), against the oracle's dense algebra and the two-pass row kernels of pop_row.cu."""
    from pop_b200 import _lib as L
    from pop_b200.arrowhead import _bind_stream
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Mo.n
    rng = np.random.default_rng(12)
    Rm, Fm = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    if basis == "c1":
        Kd = Kd + 0.3 * Md
        K = K + 0.3 * M
    Sd, Td = Kd + 3.7 * Md, 0.5 * Md - Kd
    Xd = np.linalg.solve(Sd, Rm.T).T
    dd = P.ElementDecomposition(M)
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    R, Fd = dd.panel(), dd.panel()
    Fd.copy_(torch.from_numpy(Fm).cuda())
    ctx = M.data.ctx
    _bind_stream(ctx)
    got = {}
    for mode in ("tma", "tma4", "two-pass"):
        monkeypatch.setenv("POP_ROW2", "0" if mode == "two-pass" else "1")
        monkeypatch.setenv("POP_ROW2_DC", "4" if mode == "tma4" else "8")     # degrees per TMA box / region of the row block buffer
        monkeypatch.setenv("POP_ROW1", "0")
        monkeypatch.setenv("POP_ROW2_CS", str(cs))
        monkeypatch.setenv("POP_ROW2_STRICT", "1")    # the forced plan must apply: no silent fallback
        for what, Top, gamma, Fp, want in (("half step", T, -0.4, Fd, Xd @ Td - 0.4 * Fm), ("product", T, 0.0, None, Xd @ Td), ("solve", None, 0.0, None, Xd)):
            R.copy_(torch.from_numpy(Rm).cuda())
            L.check(L.lib().pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, Top.data._h if Top is not None else None, gamma,
                                                 Fp.data_ptr() if Fp is not None else None, Fp.stride(1) if Fp is not None else 0, R.data_ptr(), R.stride(1)))
            torch.cuda.synchronize()
            got[mode, what] = host(R)
            assert relerr(got[mode, what], want) < TOL, (mode, what)
    for what in ("half step", "product", "solve"):
        assert relerr(got["tma", what], got["two-pass", what]) < 1e-13, what
        assert np.array_equal(got["tma", what], got["tma4", what]), what
    # the same kernel behind pop_solve(side = RIGHT) on a row panel with fewer rows than columns (ragged last row block)
    monkeypatch.setenv("POP_ROW2", "1")
    Rp = rng.standard_normal((13, n))
    assert relerr(host(P.rdiv(dev(Rp), Fs, algo=P.SOLVE_FUSED)), np.linalg.solve(Sd, Rp.T).T) < TOL
    dd.close()


@pytest.mark.parametrize("basis,m,N", [("dirichlet", 64, 7), ("c1", 64, 4), ("dirichlet", 128, 5), ("dirichlet", 16, 12), ("c1", 8, 9), ("dirichlet", 16, 64),
                                       ("c1", 16, 33)])
def test_single_pass_row_remote_mailboxes(basis, m, N, monkeypatch):
    """csrc/pop_row2.cu in REMOTE mode -- the CTAs that share a row block meet through mailboxes and flags in global memory,
    the way the GPUs of a multi-GPU run do -- forced on one GPU (POP_ROW2_REMOTE=1; m = 64: 2 CTAs per row block, m = 128: 4;
    m <= 16: blocks of 16 rows, the shape of an 8-GPU run):
    against dense algebra and the cluster mode of the same kernel, repeated calls (mailbox parity and epochs)."""
    from pop_b200 import _lib as L
    from pop_b200.arrowhead import _bind_stream
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Mo.n
    rng = np.random.default_rng(13)
    Rm, Fm = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    if basis == "c1":
        Kd = Kd + 0.3 * Md
        K = K + 0.3 * M
    Sd, Td = Kd + 3.7 * Md, 0.5 * Md - Kd
    Xd = np.linalg.solve(Sd, Rm.T).T
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    ctx = M.data.ctx
    _bind_stream(ctx)
    monkeypatch.setenv("POP_ROW1", "0")
    monkeypatch.setenv("POP_ROW2", "1")
    monkeypatch.setenv("POP_ROW2_STRICT", "1")
    got = {}
    for mode in ("remote", "cluster"):
        monkeypatch.setenv("POP_ROW2_REMOTE", "1" if mode == "remote" else "0")
        dd = P.ElementDecomposition(M)                       # the mailboxes are allocated with the decomposition
        R, Fd = dd.panel(), dd.panel()
        Fd.copy_(torch.from_numpy(Fm).cuda())
        for rep in range(3):                                 # epochs 1, 2, 3: both mailbox parities, reuse of a slot
            for what, Top, gamma, Fp, want in (("half step", T, -0.4, Fd, Xd @ Td - 0.4 * Fm), ("solve", None, 0.0, None, Xd)):
                R.copy_(torch.from_numpy(Rm).cuda())
                L.check(L.lib().pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, Top.data._h if Top is not None else None, gamma,
                                                     Fp.data_ptr() if Fp is not None else None, Fp.stride(1) if Fp is not None else 0, R.data_ptr(), R.stride(1)))
                torch.cuda.synchronize()
                got[mode, what] = host(R)
                assert relerr(got[mode, what], want) < TOL, (mode, what, rep)
        dd.close()
    for what in ("half step", "solve"):
        # same arithmetic; the hat scan splits a row into 4 (16-row blocks) or 8 (8-row blocks) segments: last-bit differences
        assert relerr(got["remote", what], got["cluster", what]) < 1e-14, what


def test_heat_steps_vs_oracle():
    # examples/heat.jl:58-63 (1D) and its factored 2D extension
    m, N = 6, 10
    M, Lp, Mo, Lo = _assemble("c1", m, N)
    K = -Lp
    Ko = O.packed_axpby(-1.0, Lo, 0.0, Mo)
    n = Mo.n
    rng = np.random.default_rng(11)
    u0 = rng.standard_normal((n, 5))
    U = dev(u0)
    P.heat_steps(M, K, U, 1e-3, 7, dims=1)
    want = O.heat_steps_packed(Mo, Ko, u0, 1e-3, 7)
    assert relerr(host(U), want) < TOL
    U0 = rng.standard_normal((n, n))
    U = dev(U0)
    P.heat_steps(M, K, U, 1e-3, 3, dims=2)
    assert relerr(host(U), O.heat2d_steps_packed(Mo, Ko, U0, 1e-3, 3)) < TOL


# --------------------------------------------------------------------------------------
# BASELINE shapes: size-independent properties at (near) full size
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("basis,m,N,ncols", [("c1", 1024, 40, 192), ("dirichlet", 256, 128, 160), ("dirichlet", 128, 64, 300), ("dirichlet", 256, 64, 150)])
def test_baseline_shapes_residual_and_linearity(basis, m, N, ncols):
    """cfg 2/3/4/5 column shapes: fused == staged, oracle agreement on a column sample,
    residual ||S X - B|| through the independent product kernel, linearity."""
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    S = -Lp + M
    F = P.reversecholesky(S)
    n = S.n
    g = torch.Generator(device="cuda").manual_seed(0)
    ld = (n + 1) & ~1
    Bt = torch.randn((ncols, ld), generator=g, dtype=torch.float64, device="cuda").t()[:n]
    Xf = P.ldiv(F, Bt)                                   # auto -> fused
    Xs = P.ldiv(F, Bt, algo=P.SOLVE_STAGED)
    assert float(torch.linalg.norm(Xf - Xs) / torch.linalg.norm(Xs)) < TOL
    Xm = P.ldiv(F, Bt, algo=P.SOLVE_FUSED_SMEM)          # shared-memory-resident single-pass kernel
    assert float(torch.linalg.norm(Xm - Xs) / torch.linalg.norm(Xs)) < TOL
    R = S @ Xf
    v = torch.randn(n, generator=g, dtype=torch.float64, device="cuda")
    for _ in range(30):                                                     # power iteration: ||S||_2 from below
        v = S @ v
        normS = float(torch.linalg.norm(v))
        v = v / normS
    # normwise backward error  ||S X - B|| / (||S|| ||X|| + ||B||)
    assert float(torch.linalg.norm(R - Bt) / (normS * torch.linalg.norm(Xf) + torch.linalg.norm(Bt))) < TOL
    So = O.packed_axpby(-1.0, Lo, 1.0, Mo)
    Uo = O.packed_revchol(So)
    Ug = product_to_packed(F.factor)
    for fld in ("Ad", "Au", "B"):
        assert relerr(getattr(Ug, fld), getattr(Uo, fld)) < TOL, fld           # factor parity
    assert relerr(Ug.D, Uo.D) < TOL
    sample = np.asfortranarray(host(Bt[:, :8]))
    want = O.packed_solve(Uo, sample.copy())
    # FP64 noise floor of this shape: two CPU restatements of the reference algorithm (NumPy and C)
    # on the same inputs; at n=40961 they already differ by 1.4e-12
    from oracle import c_oracle as CO
    floor = relerr(CO.solve(CO.revchol(So), sample.copy(order="F")), want)
    diff = relerr(host(Xf[:, :8]), want)
    print(f"[parity] {basis} m={m} N={N}: gpu-vs-oracle {diff:.3e}, cpu-vs-cpu floor {floor:.3e}")
    assert diff < max(TOL, 3 * floor)
    # and the GPU result is as accurate as the CPU one against an extended-precision solve
    Sl = O.packed_astype(So, np.longdouble)
    xl = O.packed_solve(O.packed_revchol(Sl), sample[:, :2].astype(np.longdouble))
    err_gpu = float(np.linalg.norm((host(Xf[:, :2]) - xl).astype(np.float64)) / np.linalg.norm(xl.astype(np.float64)))
    err_cpu = float(np.linalg.norm((want[:, :2] - xl).astype(np.float64)) / np.linalg.norm(xl.astype(np.float64)))
    assert err_gpu < 1.5 * err_cpu + 1e-13, (err_gpu, err_cpu)
    X2 = P.ldiv(F, 2.0 * Bt[:, :16] - 3.0 * Bt[:, 16:32])
    assert float(torch.linalg.norm(X2 - (2.0 * Xf[:, :16] - 3.0 * Xf[:, 16:32])) / torch.linalg.norm(X2)) < TOL


# --------------------------------------------------------------------------------------
# eigvals (test/test_arrowhead.jl:163-176; examples/adi_poisson.jl:91-94)
# --------------------------------------------------------------------------------------
def _dense_sym(Po):
    return O.unpack(O.packed_symmetrize(Po)).to_dense()


def test_eigvals_reference_testset():
    """test/test_arrowhead.jl:163-176: C1, range(-1,1;length=4), 5 blocks: eigvals(Symmetric(Delta)), eigvals(Symmetric(M)),
    eigvals(Symmetric(Delta), Symmetric(M)) against the dense (LAPACK) values the reference compares with."""
    import scipy.linalg as sla
    M, Lp, Mo, Lo = _assemble("c1", 3, 5)
    Md, Ld = _dense_sym(Mo), _dense_sym(Lo)
    for got, want in ((P.eigvals(Lp), np.linalg.eigvalsh(Ld)), (P.eigvals(M), np.linalg.eigvalsh(Md)),
                      (P.eigvals(Lp, M), sla.eigh(Ld, Md, eigvals_only=True))):
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("basis,m,N", [("dirichlet", 6, 10), ("c1", 5, 7), ("dirichlet", 16, 24)])
def test_eigvals_pencil_full_and_extreme(basis, m, N):
    import scipy.linalg as sla
    M, Lp, Mo, Lo = _assemble(basis, m, N)
    S = -Lp + M                                         # Helmholtz operator, positive definite for both bases
    Sd, Md = _dense_sym(O.packed_axpby(-1.0, Lo, 1.0, Mo)), _dense_sym(Mo)
    want = sla.eigh(Sd, Md, eigvals_only=True)
    got = P.eigvals(S, M)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
    # a short run: the extreme Ritz values with their residual bounds
    th, rs = P.eigvals(S, M, k=min(40, S.n), return_residuals=True)
    assert abs(th[-1] - want[-1]) <= rs[-1] + 1e-12 * want[-1]
    assert th[-1] <= want[-1] * (1 + 1e-12) and th[0] >= want[0] * (1 - 1e-12)


def test_spectrum_bounds_enclose_dense():
    """The ADI interval of examples/adi_poisson.jl:91-94 (dense eigmin / eigmax of U \\ (U \\ M)') from Lanczos."""
    import scipy.linalg as sla
    M, Lp, Mo, Lo = _assemble("dirichlet", 12, 20)
    K = -Lp
    Kd, Md = _dense_sym(O.packed_axpby(-1.0, Lo, 0.0, Mo)), _dense_sym(Mo)
    lam = sla.eigh(Md, Kd, eigvals_only=True)
    a, b = P.spectrum_bounds(M, K)
    assert a <= lam[0] and b >= lam[-1]
    assert a > 0.99 * lam[0] and b < 1.01 * lam[-1]
    assert P.adi_num_iterations(a, b, -b, -a) == P.adi_num_iterations(lam[0], lam[-1], -lam[-1], -lam[0])


# --------------------------------------------------------------------------------------
# race evidence without compute-sanitizer (closed on this pool, profiles/r2_sanitizer_closed.log): the kernels that rest on
# mailbox / mbarrier / flag protocols must return bitwise identical results run after run, for every cluster size
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["fused2-cs1", "fused2-cs4", "fused2-cs8", "fused-smem", "row2-cs1", "row2-cs4", "row2-cs8", "row1-cs4", "two-pass-rows"])
def test_protocol_kernels_are_bitwise_reproducible(kernel, monkeypatch):
    from pop_b200 import _lib as L
    from pop_b200.arrowhead import _bind_stream
    m, N = 16, 24
    M, Lp, Mo, Lo = _assemble("dirichlet", m, N)
    K = -Lp
    n = Mo.n
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    rng = np.random.default_rng(21)
    R0, Fm = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    ctx = M.data.ctx
    _bind_stream(ctx)
    Fd = dev(Fm)
    first = None
    if kernel.startswith("fused"):
        if kernel == "fused-smem":
            monkeypatch.setenv("POP_DISABLE_FUSED2", "1")
        else:
            monkeypatch.setenv("POP_F2_CS", kernel[-1])

        def run():
            R = dev(R0)
            L.check(L.lib().pop_solve_mul_axpy(ctx.handle, Fs.factor._h, T.data._h, -0.4, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1), n))
            return R
    else:
        dd = P.ElementDecomposition(M)
        Fp = dd.panel()
        Fp.copy_(torch.from_numpy(Fm).cuda())
        monkeypatch.setenv("POP_ROW2", "1" if kernel.startswith("row2") else "0")
        monkeypatch.setenv("POP_ROW1", "2" if kernel.startswith("row1") else "0")
        if kernel.startswith("row2"):
            monkeypatch.setenv("POP_ROW2_CS", kernel[-1])
            monkeypatch.setenv("POP_ROW2_STRICT", "1")
        if kernel.startswith("row1"):
            monkeypatch.setenv("POP_ROW1_CS", kernel[-1])
            monkeypatch.setenv("POP_ROW1_STRICT", "1")

        def run():
            R = dd.panel()
            R.copy_(torch.from_numpy(R0).cuda())
            L.check(L.lib().pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, T.data._h, -0.4, Fp.data_ptr(), Fp.stride(1), R.data_ptr(), R.stride(1)))
            return R
    for it in range(25):
        got = host(run())
        torch.cuda.synchronize()
        if first is None:
            first = got
            Xd = np.linalg.solve(O.unpack(O.packed_symmetrize(O.packed_axpby(-1.0, Lo, 3.7, Mo))).to_dense(), R0 if kernel.startswith("fused") else R0.T)
            Td = O.unpack(O.packed_symmetrize(O.packed_axpby(0.5, Mo, 1.0, Lo))).to_dense()
            want = Td @ Xd - 0.4 * Fm if kernel.startswith("fused") else Xd.T @ Td - 0.4 * Fm
            assert relerr(got, want) < TOL
        else:
            assert np.array_equal(got, first), (kernel, it)
