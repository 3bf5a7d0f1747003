"""The C restatement (oracle/pop_oracle.c) against the NumPy oracle it scales up."""
import numpy as np
import pytest

from oracle import arrowhead_oracle as O
from oracle import c_oracle as CO


@pytest.mark.parametrize("basis,m,N", [(O.C1, 3, 10), (O.DIRICHLET, 4, 12), (O.C1, 33, 17), (O.DIRICHLET, 64, 24), (O.C1, 5, 2)])
def test_c_oracle_matches_numpy_oracle(basis, m, N):
    h = 2.0 / m
    Mo, Lo = O.assemble_mass(basis, m, N, h), O.assemble_weaklaplacian(basis, m, N, h)
    S = O.packed_axpby(-1.0, Lo, 2.5, Mo)
    U1, U2 = O.packed_revchol(S), CO.revchol(S)
    np.testing.assert_allclose(O.unpack(U2).to_dense(), O.unpack(U1).to_dense(), rtol=1e-13, atol=1e-15)
    rng = np.random.default_rng(0)
    B = np.asfortranarray(rng.standard_normal((S.n, 9)))
    X1 = O.packed_solve(U1, B.copy())
    X2 = CO.solve(U2, B.copy(order="F"))
    np.testing.assert_allclose(X2, X1, rtol=1e-12, atol=1e-14)
    Y1 = O.packed_mul_left(0.7, O.packed_symmetrize(S), B, 0.0, np.zeros_like(B))
    Y2 = CO.mul_sym(S, B, alpha=0.7)
    np.testing.assert_allclose(Y2, Y1, rtol=1e-13, atol=1e-14)
    assert CO.num_threads() >= 1


def test_c_oracle_general_bands_and_not_pd():
    from tests.util import random_general
    rng = np.random.default_rng(1)
    A = random_general(rng, 7, 6, nB=2, nC=0, lD=0, diag1=True, same_D=False)
    for k in range(A.m):
        A.D[k] += 5 * np.eye(A.q)
    pk = O.pack(A, lD=0)
    U = CO.revchol(pk)
    np.testing.assert_allclose(O.unpack(U).to_dense(), O.dense_reversecholesky(O.symmetric_from_upper(A).to_dense()), rtol=1e-12, atol=1e-14)
    A.D[2][1, 1] = -3.0
    with pytest.raises(np.linalg.LinAlgError):
        CO.revchol(O.pack(A, lD=0))


def test_reference_loop_structure_variant_is_the_same_arithmetic():
    """oracle_solve_refloop (columns serial, elements threaded above / serial below, as src/arrowhead.jl:439,480 run) performs
    the operations of oracle_solve in the same order per entry: bit-identical results."""
    for basis, m, N in ((O.C1, 9, 7), (O.DIRICHLET, 12, 10)):
        h = 2.0 / m
        S = O.packed_axpby(-1.0, O.assemble_weaklaplacian(basis, m, N, h), 1.0, O.assemble_mass(basis, m, N, h))
        U = CO.revchol(S)
        X = np.asfortranarray(np.random.default_rng(3).standard_normal((S.n, 5)))
        a = CO.solve(U, X.copy(order="F"))
        b = CO.solve_refloop(U, X.copy(order="F"))
        assert np.array_equal(a, b)
