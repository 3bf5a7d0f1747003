"""Right-hand-side pipeline of examples/adi_poisson.jl:60-76,105: per-cell Legendre analysis, the
conversion P\\C and the weak form (Q'P) fp (Q'P)'.  CPU tests pin the oracle (closed forms against quadrature and the
reference's own identities); GPU tests compare the CUDA kernels with the oracle and run the example end to end
against its own `@test u_exact.(z) ≈ Ua`."""
import numpy as np
import pytest

from oracle import arrowhead_oracle as O

TOL = 1e-12


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
@pytest.mark.parametrize("m,N", [(3, 6), (5, 9), (2, 4)])
def test_lowering_closed_form_matches_quadrature(basis, m, N):
    # P \ C (src/continuouspolynomial.jl:226-235) and (P\C)(C\Q) (src/dirichlet.jl:35-53)
    assert np.abs(O.lowering_dense(basis, m, N) - O.lowering_by_quadrature(basis, m, N)).max() < 1e-14


@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
def test_lowering_reproduces_the_mass_matrix(basis):
    # test_continuouspolynomial.jl:64: R[KR,JR]'*((P'P)[KR,KR]*R[KR,JR]) ≈ (C'C)[JR,JR], KR one block longer than JR
    m, N = 4, 8
    for h in (0.5, np.array([0.2, 0.7, 0.4, 0.7])):
        R = O.lowering_dense(basis, m, N + 1)
        G = O.legendre_mass_diag(m, N + 1, h)
        Md = O.unpack(O.packed_symmetrize(O.assemble_mass(basis, m, N, h))).to_dense()
        n = Md.shape[0]
        assert np.abs(((R.T * G) @ R)[:n, :n] - Md).max() < 1e-15


@pytest.mark.parametrize("kind", ["gauss", "chebyshev1"])
def test_cell_analysis_is_exact_on_polynomials(kind):
    n, p = 3, 9
    z, T = O.legendre_grid_transform(p, kind)
    assert np.abs(T @ np.polynomial.legendre.legvander(z, p - 1) - np.eye(p)).max() < 1e-13
    rng = np.random.default_rng(0)
    coef = rng.standard_normal((p, n, p, n))                       # [a, i, b, j]
    Pz = np.polynomial.legendre.legvander(z, p - 1)                 # [s, a]
    vals = np.einsum("sa,aibj,tb->isjt", Pz, coef, Pz).reshape(n * p, n * p)
    F = O.analysis_2d(vals, n, p, T)
    assert relerr(F, coef.reshape(n * p, n * p)) < 1e-13           # row a*n + i, column b*n + j


def _poisson_example(K, p):
    """examples/adi_poisson.jl:78-117 at a small size: returns everything the GPU test needs too."""
    r = np.linspace(-1, 1, K + 1)
    h = 2.0 / K
    f = lambda x, y: -2 * np.sin(np.pi * x) * (2 * np.pi * y * np.cos(np.pi * y) + (1 - np.pi ** 2 * y ** 2) * np.sin(np.pi * y))
    u = lambda x, y: np.sin(np.pi * x) * np.sin(np.pi * y) * y ** 2
    z, T = O.legendre_grid_transform(p)
    xs = ((r[:-1] + r[1:])[:, None] / 2 + h / 2 * z[None, :]).reshape(-1)
    vals = f(xs[:, None], xs[None, :])
    return r, h, f, u, z, T, xs, vals


def test_adi_poisson_example_end_to_end_oracle():
    # the reference's own check: @test u_exact.(z) ≈ Ua  (examples/adi_poisson.jl:117)
    K, p = 4, 22
    r, h, f, u, z, T, xs, vals = _poisson_example(K, p)
    fp = O.analysis_2d(vals, K, p, T)
    F = O.weakform_rhs(O.DIRICHLET, K, p, h, fp)
    Mo = O.assemble_mass(O.DIRICHLET, K, p, h)
    Ko = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.DIRICHLET, K, p, h), 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    lam = np.linalg.eigvals(np.linalg.solve(Kd, Md)).real
    X = O.adi_packed(Mo, Ko, F, lam.min(), lam.max(), -lam.max(), -lam.min())
    Y = np.linalg.solve(Kd, X.T).T                                  # Y = (U' \ (U \ X'))'
    pts = np.linspace(-1, 1, 41)
    E = O.basis_eval_matrix(O.DIRICHLET, r, p, pts)
    Ua = E @ Y @ E.T
    assert np.abs(Ua - u(pts[:, None], pts[None, :])).max() < 1e-10


def test_transform_reproduces_the_function_oracle():
    # test_continuouspolynomial.jl:21-27: C[0.1,Block.(1:10)]' * (F * exp.(g)) ≈ exp(0.1) on r = range(-1,1;length=10)
    r = np.linspace(-1, 1, 10)
    m, N = 9, 11                                  # Block(10) of C <-> 11 Legendre degrees per element
    z, T = O.legendre_grid_transform(N)
    x = ((r[:-1] + r[1:])[:, None] / 2 + np.diff(r)[:, None] / 2 * z[None, :])
    leg = (np.exp(x) @ T.T).T.reshape(-1)         # degree a of cell k at a*m + k
    c = O.c1_from_legendre(O.C1, m, N, leg[:, None])[:, 0]
    assert c.size == 91                           # transform(C,exp)[1:91]
    assert abs(O.c1_evaluate(c, r, N - 1, 0.1) - np.exp(0.1)) < 1e-13
    # round trip F \ (F * vals) ≈ vals: converting back with P\C and synthesising on the grid
    R = O.lowering_dense(O.C1, m, N)[:, :c.size]
    back = (R @ c).reshape(N, m)
    assert np.abs(np.polynomial.legendre.legvander(z, N - 1) @ back - np.exp(x).T).max() < 1e-13


import os  # noqa: E402

RG = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rhs_golden.npz"))


def test_oracle_reproduces_rhs_fixtures():
    """tests/golden/rhs_golden.npz (generator beside it) was produced without the closed forms: quadrature of the basis,
    direct projection, synthesis of known coefficients."""
    for name, basis in (("wf_c1_uniform", O.C1), ("wf_dirichlet_graded", O.DIRICHLET)):
        pts, N = RG[f"{name}/points"], int(RG[f"{name}/N"])
        assert relerr(O.weakform_rhs(basis, len(pts) - 1, N, np.diff(pts), RG[f"{name}/fp"]), RG[f"{name}/F"]) < 1e-13
    r, N = RG["transform_exp/points"], int(RG["transform_exp/N"])
    z, T = O.legendre_grid_transform(N + 1)
    x = (r[:-1] + r[1:])[:, None] / 2 + np.diff(r)[:, None] / 2 * z[None, :]
    leg = (np.exp(x) @ T.T).T.reshape(-1)
    # interpolation at 11 nodes per element against the exact expansion: equal up to the truncation of exp at degree 10
    assert np.abs(O.c1_from_legendre(O.C1, len(r) - 1, N + 1, leg[:, None])[:, 0] - RG["transform_exp/c"]).max() < 1e-12
    n, p = int(RG["analysis/n"]), int(RG["analysis/p"])
    assert relerr(O.analysis_2d(RG["analysis/vals"], n, p, O.legendre_grid_transform(p)[1]), RG["analysis/F"]) < 1e-13


def test_oracle_reproduces_inverse_transform_and_varcoef_fixtures():
    """The fixtures of the inverse transforms come from evaluating the basis functions at the nodes, those of the
    variable-coefficient Laplacian from Gauss quadrature: neither uses the closed forms under test."""
    for name, basis in (("itransform_c1", O.C1), ("itransform_dirichlet", O.DIRICHLET)):
        pts, p = RG[f"{name}/points"], int(RG[f"{name}/p"])
        m = len(pts) - 1
        z, T = O.legendre_grid_transform(p)
        S = np.polynomial.legendre.legvander(z, p - 1)
        leg = O.legendre_from_c1(basis, m, p, RG[f"{name}/c"][:, None])[:, 0].reshape(p, m)
        assert relerr((S @ leg).T.reshape(-1), RG[f"{name}/vals"]) < 1e-13
        L = O.lowering_dense(basis, m, p)[:, :RG[f"{name}/c"].size]
        assert relerr(O.synthesis_2d(L @ RG[f"{name}/C2"] @ L.T, m, p, S), RG[f"{name}/vals2"]) < 1e-13
    for name, basis in (("varcoef_c1", O.C1), ("varcoef_dirichlet", O.DIRICHLET)):
        pts, a, N = RG[f"{name}/points"], RG[f"{name}/a"], int(RG[f"{name}/N"])
        Ld = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(basis, len(pts) - 1, N, np.diff(pts), a=a))).to_dense()
        assert np.abs(Ld - RG[f"{name}/L"]).max() < 1e-12 * np.abs(RG[f"{name}/L"]).max()


# ------------------------------------------------------------------------------------------------ CUDA kernels
def _gpu():
    import torch
    import pop_b200 as P
    return torch, P


def _dev(torch, A):
    A = np.asarray(A, dtype=np.float64)
    ld = (A.shape[0] + 1) & ~1
    buf = torch.zeros((A.shape[1], ld), dtype=torch.float64, device="cuda")
    buf[:, :A.shape[0]] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    return buf.t()[:A.shape[0]]


@pytest.mark.gpu
@pytest.mark.parametrize("n,p", [(1, 1), (1, 5), (3, 16), (7, 33), (2, 64), (5, 62), (1, 65), (3, 100), (2, 129)])
def test_analysis_2d_kernel_vs_oracle(n, p):
    torch, P = _gpu()
    z, T = P.legendre_grid_transform(p, "gauss" if p % 2 else "chebyshev1")
    assert np.abs(T - O.legendre_grid_transform(p, "gauss" if p % 2 else "chebyshev1")[1]).max() == 0
    vals = np.random.default_rng(n * 100 + p).standard_normal((n * p, n * p))
    F = P.analysis_2d(_dev(torch, vals), n, p, T)
    want = O.analysis_2d(vals, n, p, T)
    assert relerr(F.cpu().numpy(), want) < TOL


@pytest.mark.gpu
def test_analysis_2d_rejects_what_it_cannot_hold():
    torch, P = _gpu()
    z, T = P.legendre_grid_transform(4)
    with pytest.raises(P.DimensionMismatch):
        P.analysis_2d(_dev(torch, np.zeros((9, 8))), 2, 4, T)


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["c1", "dirichlet"])
@pytest.mark.parametrize("m,N,uniform", [(2, 1, True), (3, 2, True), (5, 9, True), (8, 20, False), (33, 7, False)])
def test_weakform_kernel_vs_oracle(basis, m, N, uniform):
    torch, P = _gpu()
    if basis == "dirichlet" and m < 2:
        pytest.skip("no interior node")
    rng = np.random.default_rng(m * 10 + N)
    pts = np.linspace(-1, 1, m + 1) if uniform else np.concatenate([[-1.0], np.sort(rng.uniform(-0.9, 0.9, m - 1)), [1.0]])
    Q = P.ContinuousPolynomial(1, pts) if basis == "c1" else P.DirichletPolynomial(pts)
    bo = O.C1 if basis == "c1" else O.DIRICHLET
    A = O.weakform_dense(bo, m, N, np.diff(pts))
    KR = P.Block.oneto(N)
    X = rng.standard_normal((m * N, 13))
    assert relerr(P.weakform_apply(Q, KR, _dev(torch, X), P.LEFT).cpu().numpy(), A @ X) < TOL
    Xr = rng.standard_normal((37, m * N))
    assert relerr(P.weakform_apply(Q, KR, _dev(torch, Xr), P.RIGHT).cpu().numpy(), Xr @ A.T) < TOL
    fp = rng.standard_normal((m * N, m * N))
    assert relerr(P.weakform_rhs(Q, KR, _dev(torch, fp)).cpu().numpy(), A @ fp @ A.T) < TOL
    with pytest.raises(P.DimensionMismatch):
        P.weakform_apply(Q, KR, _dev(torch, np.zeros((m * N + 1, 2))), P.LEFT)


@pytest.mark.gpu
def test_adi_poisson_example_end_to_end_gpu():
    """examples/adi_poisson.jl:78-117 through the CUDA path: analysis_2D -> (Q'P) fp (Q'P)' -> ADI -> X / Delta ->
    evaluation, against the exact solution (the reference's `@test u_exact.(z) ≈ Ua`) and against the oracle pipeline."""
    torch, P = _gpu()
    K, p = 4, 22
    r, h, f, u, z, T, xs, vals = _poisson_example(K, p)
    Q = P.DirichletPolynomial(r)
    KR = P.Block.oneto(p)
    assert np.allclose(P.cell_grid(r, z), xs)
    fp = P.analysis_2d(_dev(torch, vals), K, p, T)
    F = P.weakform_rhs(Q, KR, fp)
    Fo = O.weakform_rhs(O.DIRICHLET, K, p, h, O.analysis_2d(vals, K, p, T))
    assert relerr(F.cpu().numpy(), Fo) < TOL
    M = P.grammatrix(Q)[KR, KR]
    Kh = -P.weaklaplacian(Q)[KR, KR]
    a, b = P.spectrum_bounds(M, Kh)
    X = P.adi_poisson(M, Kh, F, a, b)
    Y = P.rdiv(X, P.reversecholesky(Kh))                            # Y = (U' \ (U \ X'))'
    pts = np.linspace(-1, 1, 41)
    E = O.basis_eval_matrix(O.DIRICHLET, r, p, pts)
    Ua = E @ Y.cpu().numpy() @ E.T
    assert np.abs(Ua - u(pts[:, None], pts[None, :])).max() < 1e-10


@pytest.mark.gpu
def test_adi_poisson_example_at_its_own_size():
    """examples/adi_poisson.jl:78-117 with the example's OWN parameters: r = range(-1, 1, 5) (four cells), p = 300 -- cells of more
    than 64 nodes go through the tiled analysis, chains of 299 degrees through the shared-memory column kernel and the two-pass
    row kernels; checked like the example does, `u_exact.(z) ≈ Ua` on its 201 x 201 grid."""
    torch, P = _gpu()
    K, p = 4, 300
    r, h, f, u, z, T, xs, vals = _poisson_example(K, p)
    Q = P.DirichletPolynomial(r)
    KR = P.Block.oneto(p)
    fp = P.analysis_2d(_dev(torch, vals), K, p, T)
    F = P.weakform_rhs(Q, KR, fp)
    M = P.grammatrix(Q)[KR, KR]
    Kh = -P.weaklaplacian(Q)[KR, KR]
    a, b = P.spectrum_bounds(M, Kh)
    X = P.adi_poisson(M, Kh, F, a, b)
    Y = P.rdiv(X, P.reversecholesky(Kh))
    pts = np.arange(-1, 1.0001, 0.01)                              # z = SVector.(-1:0.01:1, (-1:0.01:1)')
    E = O.basis_eval_matrix(O.DIRICHLET, r, p, pts)
    Ua = E @ Y.cpu().numpy() @ E.T
    exact = u(pts[:, None], pts[None, :])
    assert np.allclose(Ua, exact, rtol=1e-8, atol=1e-10)            # Julia's `≈` is rtol = sqrt(eps)
    assert np.abs(Ua - exact).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["c1", "dirichlet"])
@pytest.mark.parametrize("m,N", [(2, 2), (3, 3), (4, 7), (9, 11), (33, 40)])
def test_legendre_to_c1_kernel_vs_oracle(basis, m, N):
    torch, P = _gpu()
    bo = O.C1 if basis == "c1" else O.DIRICHLET
    Q = P.ContinuousPolynomial(1, np.linspace(0, 1, m + 1)) if basis == "c1" else P.DirichletPolynomial(np.linspace(0, 1, m + 1))
    n = Q.nh + m * (N - 2)
    rng = np.random.default_rng(m + N)
    c = rng.standard_normal((n, 6))
    R = O.lowering_dense(bo, m, N)[:, :n]
    Yl = R @ c                                    # Legendre coefficients of continuous piecewise polynomials
    want = O.c1_from_legendre(bo, m, N, Yl)
    assert relerr(want, c) < 1e-12
    got = P.legendre_to_c1(Q, N, _dev(torch, Yl), P.LEFT)
    assert relerr(got.cpu().numpy(), c) < TOL
    gotr = P.legendre_to_c1(Q, N, _dev(torch, Yl.T.copy()), P.RIGHT)
    assert relerr(gotr.cpu().numpy(), c.T) < TOL
    # a jump at an interior node is reported like the reference's ArgumentError (src/continuouspolynomial.jl:105-107)
    if m >= 2:
        bad = Yl.copy()
        bad[0, 0] += 0.5                           # degree 0 of element 0 only
        with pytest.raises(P.DiscontinuityError):
            P.legendre_to_c1(Q, N, _dev(torch, bad), P.LEFT)
        P.legendre_to_c1(Q, N, _dev(torch, bad), P.LEFT, check=False)


@pytest.mark.gpu
def test_transform_and_trapezoid_heat_like_the_example():
    """examples/heat.jl:26-47: c0 = transform(C, f)[K] for the piecewise initial condition, then 100 steps of the trapezoid
    rule A \\ (B u), A = M - h/2 Delta, B = M + h/2 Delta, against the dense oracle; and backward Euler (:58-63)."""
    torch, P = _gpu()
    r = np.linspace(-1, 1, 4)
    C1 = P.ContinuousPolynomial(1, r)
    f = lambda x: np.exp(x + 1 / 3) if x < -1 / 3 else (9 * x * x if abs(x) <= 1 / 3 else np.exp(x - 1 / 3))
    p = 41                                        # K = Block.(1:40)
    c0 = P.transform(C1, f, p)
    n = c0.numel()
    assert n == 4 + 3 * 39
    for xv in (-0.9, -0.2, 0.1, 0.77):
        assert abs(O.c1_evaluate(c0.cpu().numpy(), r, p - 1, xv) - f(xv)) < 1e-12
    KR = P.Block.oneto(p - 1)
    M, K = P.grammatrix(C1)[KR, KR], -P.weaklaplacian(C1)[KR, KR]
    Mo = O.assemble_mass(O.C1, 3, p - 1, 2 / 3)
    Ko = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.C1, 3, p - 1, 2 / 3), 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    for theta, steps in ((0.5, 100), (1.0, 100), (0.75, 7)):
        U = c0.clone().reshape(-1, 1)
        P.heat_steps(M, K, U, 1e-3, steps, theta=theta)
        want = O.heat_theta_dense(Md, Kd, c0.cpu().numpy(), 1e-3, theta, steps)
        assert relerr(U[:, 0].cpu().numpy(), want) < 1e-11, theta


@pytest.mark.gpu
def test_transform_2d_like_the_reference_test():
    # test_continuouspolynomial.jl:29-33: C[0.1, Block.(1:10)]' * V * C[0.2, Block.(1:11)] ≈ exp(0.1*cos(0.2)) (square here)
    torch, P = _gpu()
    r = np.linspace(-1, 1, 10)
    C1 = P.ContinuousPolynomial(1, r)
    p = 12
    z, T = P.legendre_grid_transform(p)
    x = P.cell_grid(r, z)
    vals = np.exp(x[:, None] * np.cos(x[None, :]))
    V = P.transform_2d(C1, _dev(torch, vals), p, T).cpu().numpy()
    E = O.basis_eval_matrix(O.C1, r, p - 1, np.array([0.1, 0.2]))
    assert abs(E[0] @ V @ E[1] - np.exp(0.1 * np.cos(0.2))) < 1e-10


@pytest.mark.gpu
def test_kernels_reproduce_rhs_fixtures():
    torch, P = _gpu()
    for name, basis in (("wf_c1_uniform", "c1"), ("wf_dirichlet_graded", "dirichlet")):
        pts, N = RG[f"{name}/points"], int(RG[f"{name}/N"])
        Q = P.ContinuousPolynomial(1, pts) if basis == "c1" else P.DirichletPolynomial(pts)
        F = P.weakform_rhs(Q, P.Block.oneto(N), _dev(torch, RG[f"{name}/fp"]))
        assert relerr(F.cpu().numpy(), RG[f"{name}/F"]) < TOL
    r, N = RG["transform_exp/points"], int(RG["transform_exp/N"])
    c = P.transform(P.ContinuousPolynomial(1, r), np.exp, N + 1)
    assert np.abs(c.cpu().numpy() - RG["transform_exp/c"]).max() < 1e-12
    n, p = int(RG["analysis/n"]), int(RG["analysis/p"])
    F = P.analysis_2d(_dev(torch, RG["analysis/vals"]), n, p, P.legendre_grid_transform(p)[1])
    assert relerr(F.cpu().numpy(), RG["analysis/F"]) < TOL


@pytest.mark.gpu
def test_kernels_reproduce_inverse_transform_and_varcoef_fixtures():
    torch, P = _gpu()
    for name, basis in (("itransform_c1", "c1"), ("itransform_dirichlet", "dirichlet")):
        pts, p = RG[f"{name}/points"], int(RG[f"{name}/p"])
        Q = P.ContinuousPolynomial(1, pts) if basis == "c1" else P.DirichletPolynomial(pts)
        x, vals = P.itransform(Q, _dev(torch, RG[f"{name}/c"][:, None])[:, 0], p)
        assert relerr(vals, RG[f"{name}/vals"]) < TOL
        z, _ = P.legendre_grid_transform(p)
        V2 = P.itransform_2d(Q, _dev(torch, RG[f"{name}/C2"]), p, P.legendre_synthesis_matrix(z, p))
        assert relerr(V2.cpu().numpy(), RG[f"{name}/vals2"]) < TOL
    for name, basis in (("varcoef_c1", "c1"), ("varcoef_dirichlet", "dirichlet")):
        pts, a, N = RG[f"{name}/points"], RG[f"{name}/a"], int(RG[f"{name}/N"])
        Q = P.ContinuousPolynomial(1, pts) if basis == "c1" else P.DirichletPolynomial(pts)
        Lv = P.weaklaplacian(Q, a=a)[P.Block.oneto(N), P.Block.oneto(N)]
        assert np.abs(Lv.to_dense() - RG[f"{name}/L"]).max() < 1e-12 * np.abs(RG[f"{name}/L"]).max()


@pytest.mark.gpu
def test_theta_scheme_two_dimensions():
    """pop_heat_steps_theta with dims = 2: U <- G U G', G = (M + theta dt K)^-1 (M - (1-theta) dt K) (the factored 2D extension
    of examples/heat.jl:26-45), against the dense oracle."""
    torch, P = _gpu()
    m, N = 5, 9
    Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
    KR = P.Block.oneto(N)
    M, K = P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]
    Mo = O.assemble_mass(O.DIRICHLET, m, N, 2 / m)
    Ko = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.DIRICHLET, m, N, 2 / m), 0.0, Mo)
    Md, Kd = O.unpack(O.packed_symmetrize(Mo)).to_dense(), O.unpack(O.packed_symmetrize(Ko)).to_dense()
    n = Md.shape[0]
    U0 = np.random.default_rng(5).standard_normal((n, n))
    for theta in (0.5, 1.0):
        G = np.linalg.solve(Md + theta * 1e-2 * Kd, Md - (1 - theta) * 1e-2 * Kd)
        want = np.linalg.matrix_power(G, 6) @ U0 @ np.linalg.matrix_power(G, 6).T
        U = _dev(torch, U0)
        P.heat_steps(M, K, U, 1e-2, 6, dims=2, theta=theta)
        assert relerr(U.cpu().numpy(), want) < 1e-11, theta


# ------------------------------------------------------------------------------------------------ inverse transforms
@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
def test_inverse_transform_oracle_round_trip_and_pointwise(basis):
    """`F \\ (F * vals) ≈ vals` (test/test_continuouspolynomial.jl:19,26; test_dirichlet.jl:70) for the oracle's restatement of
    `Pl \\ cfs` (src/continuouspolynomial.jl:129-146): conversion to Legendre coefficients + synthesis; the synthesised
    values are the expansion evaluated at the nodes (independent evaluation through basis_eval_matrix)."""
    m, p = 5, 9
    r = np.linspace(-1, 1, m + 1)
    nh = m + 1 if basis == O.C1 else m - 1
    n = nh + m * (p - 2)
    rng = np.random.default_rng(11)
    c = rng.standard_normal((n, 3))
    leg = O.legendre_from_c1(basis, m, p, c)                              # [degree a of element k at a m + k]
    assert relerr(O.c1_from_legendre(basis, m, p, leg), c) < 1e-12        # F * (F \ c) = c
    z, T = O.legendre_grid_transform(p)
    S = np.polynomial.legendre.legvander(z, p - 1)
    assert np.allclose(S @ T, np.eye(p), atol=1e-13)
    x = ((r[:-1] + r[1:])[:, None] / 2 + np.diff(r)[:, None] / 2 * z[None, :]).reshape(-1)
    vals = np.einsum("sa,akc->ksc", S, leg.reshape(p, m, 3)).reshape(m * p, 3)
    E = O.basis_eval_matrix(basis, r, p - 1, x)                           # blocks 1..p-1 of the basis: hats and W_2..W_{p-1}
    assert E.shape[1] == n
    assert relerr(vals, E @ c) < 1e-12


def test_synthesis_2d_oracle_inverts_analysis_2d():
    n, p = 3, 7
    z, T = O.legendre_grid_transform(p)
    S = np.polynomial.legendre.legvander(z, p - 1)
    rng = np.random.default_rng(5)
    V = rng.standard_normal((n * p, n * p))
    assert relerr(O.synthesis_2d(O.analysis_2d(V, n, p, T), n, p, S), V) < 1e-12


@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
def test_variable_coefficient_laplacian_closed_form_matches_quadrature(basis):
    """-(D*C)' * (a .* (D*C)) of examples/heat.jl:77-79 for a piecewise-constant a: element integrals scaled by a_k."""
    m, N = 6, 7
    rng = np.random.default_rng(2)
    hk = 0.2 + rng.random(m)
    a = 0.5 + rng.random(m)
    Lp = O.assemble_weaklaplacian(basis, m, N, hk, a=a)
    Ld = O.unpack(O.packed_symmetrize(Lp)).to_dense()
    _, Lq = O.quadrature_matrices(basis, m, N, hk, a=a)
    assert np.abs(Ld - Lq).max() < 1e-12 * np.abs(Lq).max()
    # a = 1 is the plain weak Laplacian
    L1 = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(basis, m, N, hk, a=np.ones(m)))).to_dense()
    L0 = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(basis, m, N, hk))).to_dense()
    assert np.array_equal(L1, L0)


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["c1", "dirichlet"])
@pytest.mark.parametrize("m,N", [(2, 2), (3, 3), (4, 7), (9, 11), (33, 40)])
def test_c1_to_legendre_kernel_vs_oracle(basis, m, N):
    torch, P = _gpu()
    bo = O.C1 if basis == "c1" else O.DIRICHLET
    Q = P.ContinuousPolynomial(1, np.linspace(0, 1, m + 1)) if basis == "c1" else P.DirichletPolynomial(np.linspace(0, 1, m + 1))
    n = Q.nh + m * (N - 2)
    rng = np.random.default_rng(3 * m + N)
    c = rng.standard_normal((n, 5))
    want = O.legendre_from_c1(bo, m, N, c)
    got = P.c1_to_legendre(Q, N, _dev(torch, c), P.LEFT)
    assert relerr(got.cpu().numpy(), want) < TOL
    gotr = P.c1_to_legendre(Q, N, _dev(torch, c.T.copy()), P.RIGHT)
    assert relerr(gotr.cpu().numpy(), want.T) < TOL
    # the two conversions are inverses of each other (F * (F \ c) = c)
    back = P.legendre_to_c1(Q, N, got, P.LEFT)
    assert relerr(back.cpu().numpy(), c) < TOL
    with pytest.raises(P.DimensionMismatch):
        P.c1_to_legendre(Q, N + 1, _dev(torch, c), P.LEFT)


@pytest.mark.gpu
@pytest.mark.parametrize("n,p", [(1, 1), (1, 5), (3, 16), (7, 33), (2, 64), (5, 62), (1, 65), (3, 100), (2, 129)])
def test_synthesis_2d_kernel_vs_oracle(n, p):
    torch, P = _gpu()
    z, T = P.legendre_grid_transform(p)
    S = P.legendre_synthesis_matrix(z, p)
    rng = np.random.default_rng(p + n)
    F = rng.standard_normal((n * p, n * p))
    got = P.synthesis_2d(_dev(torch, F), n, p, S).cpu().numpy()
    assert relerr(got, O.synthesis_2d(F, n, p, S)) < TOL
    # F \ (F * vals) ≈ vals on the device (test_continuouspolynomial.jl:33)
    V = rng.standard_normal((n * p, n * p))
    back = P.synthesis_2d(P.analysis_2d(_dev(torch, V), n, p, T), n, p, S).cpu().numpy()
    assert relerr(back, V) < 1e-11


@pytest.mark.gpu
def test_inverse_transforms_like_the_reference_tests():
    """test_continuouspolynomial.jl:19-33 and test_dirichlet.jl:70: `F \\ (F * exp.(g)) ≈ exp.(g)` in one dimension and
    `F \\ V ≈ vals` for f(x,y) = exp(x cos(y)) in two."""
    torch, P = _gpu()
    r = np.linspace(-1, 1, 10)
    C1 = P.ContinuousPolynomial(1, r)
    p = 11
    c = P.transform(C1, np.exp, p)
    x, vals = P.itransform(C1, c, p)
    assert relerr(vals, np.exp(x)) < 1e-12
    z, T = P.legendre_grid_transform(p)
    S = P.legendre_synthesis_matrix(z, p)
    xg = P.cell_grid(r, z)
    v2 = np.exp(xg[:, None] * np.cos(xg[None, :]))
    V = P.transform_2d(C1, _dev(torch, v2), p, T)
    back = P.itransform_2d(C1, V, p, S).cpu().numpy()
    assert relerr(back, v2) < 1e-11
    # Dirichlet basis: a field that vanishes on the boundary
    Q = P.DirichletPolynomial(r)
    w2 = np.outer((1 - xg ** 2) * np.exp(xg), (1 - xg ** 2) * np.cos(xg))
    W = P.transform_2d(Q, _dev(torch, w2), p, T)
    assert relerr(P.itransform_2d(Q, W, p, S).cpu().numpy(), w2) < 1e-11


@pytest.mark.gpu
def test_dirichlet_transform_testset_of_the_reference():
    """test/test_dirichlet.jl:58-71 ("transform"): r = range(-1, 1; length=3), Block(20): `Q[0.1,1:39]' * (F*f.(x)) ≈ f(0.1)`,
    `F \\ (F*f.(x)) ≈ f.(x)`, and in two dimensions `Q[0.1,KR]' * V * Q[0.2,KR] ≈ f(0.1,0.2)`, `F \\ V ≈ vals` (square here: the
    reference's second dimension has 21 blocks)."""
    torch, P = _gpu()
    r = np.linspace(-1, 1, 3)
    Q = P.DirichletPolynomial(r)
    p = 21                                          # Block(20): hats + bubbles up to degree 20
    f = lambda x: (1 - x * x) * np.exp(x)
    c = P.transform(Q, f, p)
    assert c.numel() == 39
    E = O.basis_eval_matrix(O.DIRICHLET, r, p - 1, np.array([0.1, 0.2]))
    assert abs(E[0] @ c.cpu().numpy() - f(0.1)) < 1e-12
    x, vals = P.itransform(Q, c, p)
    assert relerr(vals, f(x)) < 1e-12
    z, T = P.legendre_grid_transform(p)
    xg = P.cell_grid(r, z)
    f2 = lambda x, y: (1 - x ** 2) * (1 - y ** 2) * np.exp(x * np.cos(y))
    v2 = f2(xg[:, None], xg[None, :])
    V = P.transform_2d(Q, _dev(torch, v2), p, T)
    assert abs(E[0] @ V.cpu().numpy() @ E[1] - f2(0.1, 0.2)) < 1e-11
    assert relerr(P.itransform_2d(Q, V, p, P.legendre_synthesis_matrix(z, p)).cpu().numpy(), v2) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("basis", ["c1", "dirichlet"])
def test_variable_coefficient_heat(basis):
    """examples/heat.jl:77-83 with a piecewise-constant a (0.5 on the middle third, 1 elsewhere, as in the example): the
    operator against the oracle's closed form and Gauss quadrature, then theta-scheme steps against the dense oracle (the
    reference integrates this operator with an ODE solver, which stays out of scope)."""
    torch, P = _gpu()
    bo = O.C1 if basis == "c1" else O.DIRICHLET
    m, N = 9, 12
    r = np.linspace(-1, 1, m + 1)
    Q = P.ContinuousPolynomial(1, r) if basis == "c1" else P.DirichletPolynomial(r)
    mid = (r[:-1] + r[1:]) / 2
    a = np.where(np.abs(mid) <= 1 / 3, 0.5, 1.0)
    KR = P.Block.oneto(N)
    Lv = P.weaklaplacian(Q, a=a)[KR, KR]
    M = P.grammatrix(Q)[KR, KR]
    hk = np.diff(r)
    Lo = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(bo, m, N, hk, a=a))).to_dense()
    _, Lq = O.quadrature_matrices(bo, m, N, hk, a=a)
    assert np.abs(Lv.to_dense() - Lo).max() < 1e-13 * np.abs(Lo).max()
    assert np.abs(Lv.to_dense() - Lq).max() < 1e-12 * np.abs(Lq).max()
    Md = M.to_dense()
    n = Md.shape[0]
    rng = np.random.default_rng(8)
    u0 = rng.standard_normal(n)
    for theta, steps in ((1.0, 20), (0.5, 20)):
        U = _dev(torch, u0.reshape(-1, 1).copy())
        P.heat_steps(M, -Lv, U, 1e-3, steps, theta=theta)
        want = O.heat_theta_dense(Md, -Lo, u0, 1e-3, theta, steps)
        assert relerr(U[:, 0].cpu().numpy(), want) < 1e-11, theta
    # M + dt (-L_a) combines lazily as well
    S = (P.grammatrix(Q) - 1e-3 * P.weaklaplacian(Q, a=a))[KR, KR]
    assert np.abs(S.to_dense() - (Md - 1e-3 * Lo)).max() < 1e-13 * np.abs(Md).max()
