"""Pins the oracle: reference golden value, structural goldens of
/root/reference/test/test_arrowhead.jl, dense-equivalence testsets restated, survey goldens."""
import numpy as np
import pytest

from oracle import arrowhead_oracle as O

rng = np.random.default_rng(0)


def random_general(n, p, nB=2, nC=3, lower_B=True, lD=1, uD=2, diag1=False):
    """Arrowheads of the shapes in test/test_arrowhead.jl:9-48 (n hats, n-1 elements)."""
    m = n - 1
    A = np.diag(rng.standard_normal(n) + 10) + np.diag(rng.standard_normal(n - 1), 1) + np.diag(rng.standard_normal(n - 1), -1)
    def Bblk():
        Bm = np.zeros((n, m))
        Bm[np.arange(m), np.arange(m)] = rng.standard_normal(m)
        Bm[np.arange(m) + 1, np.arange(m)] = rng.standard_normal(m)
        return Bm
    B = [Bblk() for _ in range(nB)]
    C = [Bblk().T.copy() for _ in range(nC)]
    Dk = np.diag(rng.standard_normal(p) + 10) + np.diag(rng.standard_normal(p - 2), 2)
    if lD >= 1:
        Dk += np.diag(rng.standard_normal(p - 1), -1)
    if diag1:
        Dk += np.diag(rng.standard_normal(p - 1), 1)
    D = [Dk.copy() for _ in range(m)]     # `fill(...)`: same block for every element
    return O.BlockArrowhead(A, B, C, D)


def test_structure_goldens():
    # test_arrowhead.jl:148-161
    A = random_general(5, 5)
    assert O.block_bandwidths(A) == (3, 2)
    assert O.banded_bandwidths(A) == (13, 9)
    # test_arrowhead.jl:163-172 : bandwidths(M) == (7,7), bandwidths(Delta) == (1,1) at m=3, N=5
    M = O.unpack(O.packed_symmetrize(O.assemble_mass(O.C1, 3, 5, 2 / 3)))
    L = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(O.C1, 3, 5, 2 / 3)))
    assert O._bandwidths(M.to_dense()) == (7, 7)
    assert O._bandwidths(L.to_dense()) == (1, 1)


def test_getindex_and_pack_roundtrip():
    A = random_general(4, 5)
    S = A.to_dense()
    nh, m, q = A.nh, A.m, A.q
    for K in range(q + 1):
        for J in range(q + 1):
            for k in range(nh if K == 0 else m):
                for j in range(nh if J == 0 else m):
                    r = k if K == 0 else nh + (K - 1) * m + k
                    c = j if J == 0 else nh + (J - 1) * m + j
                    assert S[r, c] == A.getindex(K, k, J, j)
    P = O.pack(A)
    assert np.array_equal(O.unpack(P).to_dense(), S)
    assert np.array_equal(O.unpack(O.packed_transpose(P)).to_dense(), S.T)


def test_algebra():
    # test_arrowhead.jl:17-37
    A = random_general(4, 5)
    S = A.to_dense()
    assert np.array_equal(O.block_scale(2, A).to_dense(), 2 * S)
    assert np.array_equal(O.block_add(A, A).to_dense(), 2 * S)
    assert np.all(O.block_add(A, A, -1).to_dense() == 0)
    assert np.array_equal(O.block_add(A, A.transpose()).to_dense(), S + S.T)
    assert np.array_equal(O.block_add(A, A.transpose(), -1).to_dense(), S - S.T)
    P = O.pack(A)
    assert np.array_equal(O.unpack(O.packed_axpby(1.0, P, -1.0, O.packed_transpose(P))).to_dense(), S - S.T)


def test_mul():
    # test_arrowhead.jl:39-54
    A = random_general(4, 5)
    S = A.to_dense()
    n = A.n
    x = rng.standard_normal(n)
    X = rng.standard_normal((n, n))
    Xt = rng.standard_normal((n, 5))
    np.testing.assert_allclose(O.ref_mul_vec(1.0, A, x, 0.0, np.zeros(n)), S @ x, rtol=1e-13)
    np.testing.assert_allclose(O.ref_mul_vec(1.0, A.transpose(), x, 0.0, np.zeros(n)), S.T @ x, rtol=1e-13)
    np.testing.assert_allclose(O.ref_mul_mat_left(1.0, A, X, 0.0, np.zeros((n, n))), S @ X, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.ref_mul_mat_left(1.0, A, Xt, 0.0, np.zeros((n, 5))), S @ Xt, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.ref_mul_mat_right(1.0, X, A, 0.0, np.zeros((n, n))), X @ S, rtol=1e-12, atol=1e-12)
    Y0 = rng.standard_normal((n, 5))
    np.testing.assert_allclose(O.ref_mul_mat_left(0.7, A, Xt, -1.3, Y0.copy()), 0.7 * S @ Xt - 1.3 * Y0, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.packed_mul_left(0.7, O.pack(A), Xt, -1.3, Y0.copy()), 0.7 * S @ Xt - 1.3 * Y0, rtol=1e-12, atol=1e-12)


def test_triangular():
    # test_arrowhead.jl:56-81
    A = random_general(4, 5, nB=2, nC=2)
    S = A.to_dense()
    n = A.n
    c = rng.standard_normal(n)
    P = O.pack(A)
    for upper in (True, False):
        for unit in (False, True):
            T = np.triu(S) if upper else np.tril(S)
            if unit:
                T = T - np.diag(np.diag(T)) + np.eye(n)
            want = np.linalg.solve(T, c)
            got = (O.ref_ldiv_upper if upper else O.ref_ldiv_lower)(A, c.copy(), unit)
            np.testing.assert_allclose(got, want, rtol=1e-11)
            np.testing.assert_allclose(O.packed_trsm(P, c.copy(), upper, unit), want, rtol=1e-11)
            # c'/T  ==  (T' \ c)'
            wantT = np.linalg.solve(T.T, c)
            gotT = (O.ref_ldiv_lower if upper else O.ref_ldiv_upper)(A.transpose(), c.copy(), unit)
            np.testing.assert_allclose(gotT, wantT, rtol=1e-11)


@pytest.mark.parametrize("diag1", [False, True])
def test_reversecholesky_lower_B(diag1):
    # test_arrowhead.jl:83-100  (n=3, p=5, no C, D bands (0,2) without / with first superdiagonal)
    A = random_general(3, 5, nB=2, nC=0, lD=0, diag1=diag1)
    Ssym = O.symmetric_from_upper(A).to_dense()
    Uw = O.dense_reversecholesky(Ssym)
    Ug = O.ref_reversecholesky_layout(A.copy())
    Ug = O.BlockArrowhead(np.triu(Ug.A), Ug.B, [], [np.triu(d) for d in Ug.D]).to_dense()
    np.testing.assert_allclose(Ug, Uw, rtol=1e-12, atol=1e-14)
    Up = O.unpack(O.packed_revchol(O.pack(A, lD=0))).to_dense()
    np.testing.assert_allclose(Up, Uw, rtol=1e-12, atol=1e-14)


def test_reversecholesky_upper_B_dirichlet():
    # test_arrowhead.jl:137-146  (n=5: n-2 hats, n-1 elements, B upper-bidiagonal)
    n, p = 5, 5
    nh, m = n - 2, n - 1
    A = np.diag(rng.standard_normal(nh) + 10) + np.diag(rng.standard_normal(nh - 1), 1) + np.diag(rng.standard_normal(nh - 1), -1)
    def Bblk():
        Bm = np.zeros((nh, m))
        Bm[np.arange(nh), np.arange(nh)] = rng.standard_normal(nh)
        Bm[np.arange(nh), np.arange(nh) + 1] = rng.standard_normal(nh)
        return Bm
    Dk = np.diag(rng.standard_normal(p) + 10) + np.diag(rng.standard_normal(p - 2), 2)
    Ah = O.BlockArrowhead(A, [Bblk(), Bblk()], [], [Dk.copy() for _ in range(m)])
    Ssym = O.symmetric_from_upper(Ah).to_dense()
    Uw = O.dense_reversecholesky(Ssym)
    Ug = O.ref_reversecholesky_layout(Ah.copy())
    Ug = O.BlockArrowhead(np.triu(Ug.A), Ug.B, [], [np.triu(d) for d in Ug.D]).to_dense()
    np.testing.assert_allclose(Ug, Uw, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(Ug @ Ug.T, Ssym, rtol=1e-12, atol=1e-13)
    Up = O.unpack(O.packed_revchol(O.pack(Ah, lD=0))).to_dense()
    np.testing.assert_allclose(Up, Uw, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
def test_assembly_vs_quadrature(basis):
    # test_continuouspolynomial.jl:49-81,116-132 ; test_dirichlet.jl:31-42 restated as quadrature
    m, N, h = 4, 9, 0.5
    Mq, Lq = O.quadrature_matrices(basis, m, N, h)
    M = O.unpack(O.packed_symmetrize(O.assemble_mass(basis, m, N, h))).to_dense()
    L = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(basis, m, N, h))).to_dense()
    np.testing.assert_allclose(M, Mq, atol=2e-16 * 50)
    np.testing.assert_allclose(L, Lq, atol=1e-13)


def test_helmholtz_reference_golden():
    """test/test_arrowhead.jl:117-135: the reference's only known-answer value on this path."""
    r = np.linspace(-1, 1, 4)
    m, N, h = 3, 100, 2 / 3
    Mp = O.assemble_mass(O.C1, m, N, h)
    Lp = O.assemble_weaklaplacian(O.C1, m, N, h)
    S = O.packed_axpby(-1.0, Lp, 1.0, Mp)                       # parent(-Delta + M)[KR,KR]
    U = O.packed_revchol(S)
    cexp = O.c1_expand(np.exp, np.exp, r, N)
    x = np.zeros_like(cexp)
    O.packed_mul_left(1.0, O.packed_symmetrize(Mp), cexp, 0.0, x)
    c = O.packed_solve(U, x.copy())
    val = O.c1_evaluate(c, r, N, 0.1)
    assert abs(val - 1.1952730862177243) < 5e-13
    # L1 (literal block algorithm) gives the same solution vector
    Ub = O.ref_reversecholesky_layout(O.unpack(S))
    Ub = O.BlockArrowhead(np.triu(Ub.A), Ub.B, [], [np.triu(d) for d in Ub.D])
    c1 = O.ref_revchol_solve(Ub, x)
    np.testing.assert_allclose(c1, c, rtol=1e-12, atol=1e-15)
    # analytic solution of -u'' + u = e^x, u'(+-1) = 0
    e = np.e
    u = lambda t: -t * np.exp(t) / 2 + (e ** 4 * np.exp(t) + e ** 2 * np.exp(-t)) / (e ** 4 - 1)
    assert abs(val - u(0.1)) < 5e-13


def test_survey_goldens_readme_case():
    """README case (C1, range(-1,1;length=4), N=10, S=-Delta+M, n=31); values from SURVEY 8c
    (independent dense NumPy at survey time)."""
    m, N, h = 3, 10, 2 / 3
    S = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.C1, m, N, h), 1.0, O.assemble_mass(O.C1, m, N, h))
    U = O.unpack(O.packed_revchol(S)).to_dense()
    L = U.T
    g = {(1, 1): 0.9818490617583832, (2, 1): -0.8671084871570119, (2, 2): 1.6080940178727772, (4, 4): 1.309925459167635,
         (5, 1): 0.07687846187879602, (5, 2): 0.07687846187879602, (5, 5): 1.4452827020171803,
         (8, 1): -0.020179587355387243, (8, 2): 0.020179587355387243, (8, 8): 1.101222826357233,
         (11, 5): -0.006841062695774555, (11, 11): 0.9281023477723708, (31, 25): -0.0002447828624173438,
         (31, 31): 0.5621263590512574}
    for (i, j), v in g.items():
        assert abs(L[i - 1, j - 1] - v) < 1e-14, (i, j)
    assert abs(2 * np.sum(np.log(np.diag(U))) - (-10.033633705327116)) < 1e-12
    assert np.count_nonzero(U) == 67
    b = np.arange(1, 32) / 31
    x = O.packed_solve(O.packed_revchol(S), b.copy())
    for i, v in {1: 0.11098908536056581, 4: 0.18024989943421074, 5: 0.06638120377479541, 31: 3.1655746467356325}.items():
        assert abs(x[i - 1] - v) < 1e-13 * max(1, abs(v))
    Sd = O.unpack(O.packed_symmetrize(S)).to_dense()
    np.testing.assert_allclose(Sd @ x, b, atol=1e-14)


def test_survey_goldens_dirichlet():
    m, N, h = 4, 12, 0.5
    K = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.DIRICHLET, m, N, h), 0.0, O.assemble_mass(O.DIRICHLET, m, N, h))
    S = O.packed_axpby(1.0, K, 0.5, O.assemble_mass(O.DIRICHLET, m, N, h))
    U = O.unpack(O.packed_revchol(S)).to_dense()
    assert U.shape == (47, 47)
    for (i, j), v in {(1, 1): 1.7269538792779693, (1, 2): -1.087623283546823, (1, 4): 0.025357549071001476,
                      (3, 3): 2.0409052178056677, (47, 47): 0.5898029288025082}.items():
        assert abs(U[i - 1, j - 1] - v) < 1e-14
    Sd = O.unpack(O.packed_symmetrize(S)).to_dense()
    np.testing.assert_allclose(U, O.dense_reversecholesky(Sd), atol=1e-14)


def test_adi_small():
    """examples/adi_poisson.jl:42-59 on Dirichlet m=4, N=12 with manufactured Y (SURVEY 8c:
    J=42, a=1.9262e-05, b=4/pi^2, error 8e-15)."""
    m, N, h = 4, 12, 0.5
    Mp = O.assemble_mass(O.DIRICHLET, m, N, h)
    Kp = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.DIRICHLET, m, N, h), 0.0, Mp)
    Kp.B = np.zeros((0, 3, m))
    Md = O.unpack(O.packed_symmetrize(Mp)).to_dense()
    Kd = O.unpack(O.packed_symmetrize(Kp)).to_dense()
    n = Mp.n
    Y = np.random.default_rng(1).standard_normal((n, n))
    F = Kd @ Y @ Md + Md @ Y @ Kd
    lam = np.linalg.eigvals(np.linalg.solve(Kd, Md)).real
    a, b = lam.min(), lam.max()
    assert abs(a - 1.9262e-05) < 1e-8 and abs(b - 4 / np.pi ** 2) < 1e-5
    J = O.adi_num_iterations(a, b, -b, -a)
    assert J == 42
    X = O.adi_packed(Mp, Kp, F, a, b, -b, -a)
    Xd = O.adi_dense(Md, -Md, Kd, F, a, b, -b, -a)
    np.testing.assert_allclose(X, Xd, rtol=0, atol=1e-11 * np.abs(Xd).max())
    Yg = np.linalg.solve(Kd, X.T).T                              # Y = X K^{-1}
    assert np.linalg.norm(Yg - Y) / np.linalg.norm(Y) < 1e-12


def test_heat_1d():
    m, N, h = 3, 20, 2 / 3
    Mp = O.assemble_mass(O.C1, m, N, h)
    Kp = O.packed_axpby(-1.0, O.assemble_weaklaplacian(O.C1, m, N, h), 0.0, Mp)
    Md = O.unpack(O.packed_symmetrize(Mp)).to_dense()
    Kd = O.unpack(O.packed_symmetrize(Kp)).to_dense()
    u0 = np.random.default_rng(2).standard_normal(Mp.n)
    u = O.heat_steps_packed(Mp, Kp, u0, 1e-3, 10)
    w = u0.copy()
    for _ in range(10):
        w = np.linalg.solve(Md + 1e-3 * Kd, Md @ w)
    np.testing.assert_allclose(u, w, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("basis", [O.C1, O.DIRICHLET])
def test_nonuniform_mesh_closed_forms_match_quadrature(basis):
    """the per-element closed forms on a non-uniform mesh (what the reference's generic L'*G*L,
    src/continuouspolynomial.jl:259-265,293-297, evaluates to) against Gauss quadrature of the basis; a uniform
    array of element lengths reproduces the range closed forms exactly (test_continuouspolynomial.jl:69-80)."""
    pts = np.array([-1.0, -0.7, -0.55, 0.1, 0.2, 1.0])
    hk = np.diff(pts)
    m, N = hk.size, 9
    Mq, Lq = O.quadrature_matrices(basis, m, N, hk)
    Mc = O.unpack(O.packed_symmetrize(O.assemble_mass(basis, m, N, hk))).to_dense()
    Lc = O.unpack(O.packed_symmetrize(O.assemble_weaklaplacian(basis, m, N, hk))).to_dense()
    assert np.abs(Mc - Mq).max() < 1e-14
    assert np.abs(Lc - Lq).max() < 1e-12
    hu = np.full(m, 0.4)
    for f in (O.assemble_mass, O.assemble_weaklaplacian):
        a, b = f(basis, m, N, hu), f(basis, m, N, 0.4)
        for x, y in ((a.Ad, b.Ad), (a.Au, b.Au), (a.B, b.B), (a.D, b.D)):
            assert np.allclose(x, y, rtol=4e-16, atol=0)
    # the graded-mesh Helmholtz operator factors and solves like the uniform one
    So = O.packed_axpby(-1.0, O.assemble_weaklaplacian(basis, m, N, hk), 1.0, O.assemble_mass(basis, m, N, hk))
    Uo = O.packed_revchol(So)
    b = np.random.default_rng(4).standard_normal(So.n)
    x = O.packed_solve(Uo, b.copy())
    Sd = Mc - Lc
    assert np.linalg.norm(Sd @ x - b) / np.linalg.norm(b) < 1e-12
