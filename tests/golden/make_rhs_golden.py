#!/usr/bin/env python
"""Generates tests/golden/rhs_golden.npz: fixtures of the right-hand-side pipeline, each produced by a
route that is INDEPENDENT of the closed forms the oracle and the CUDA kernels use:
  * weak form (Q'P)[KR,KR]: the conversion P\\Q from Gauss quadrature of the basis functions against the Legendre
    polynomials (not from the closed-form bands), times the Legendre mass;
  * transform(C, exp)[Block.(1:10)] on r = range(-1,1;length=10) (test/test_continuouspolynomial.jl:21-27): hats by
    interpolation at the nodes, bubbles from the Legendre coefficients of the derivative (oracle.c1_expand);
  * analysis_2D: random tensor Legendre coefficients synthesised on the cell grids (the coefficients are the answer);
  * inverse transforms: random coefficients in the C1 / Dirichlet bases, values from evaluating the basis functions at the nodes;
  * variable-coefficient weak Laplacian: Gauss quadrature with the coefficient inside.
Run from the repo root:  python tests/golden/make_rhs_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import arrowhead_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    rng = np.random.default_rng(2024)
    # weak form on a uniform and on a graded mesh
    for name, basis, pts, N in (("wf_c1_uniform", O.C1, np.linspace(-1, 1, 6), 8), ("wf_dirichlet_graded", O.DIRICHLET, np.array([-1, -0.6, -0.5, 0.2, 0.45, 1.0]), 7)):
        m = len(pts) - 1
        R = O.lowering_by_quadrature(basis, m, N)
        A = R.T * O.legendre_mass_diag(m, N, np.diff(pts))[None, :]
        fp = rng.standard_normal((m * N, m * N))
        out[f"{name}/points"], out[f"{name}/N"] = pts, np.array(N)
        out[f"{name}/fp"], out[f"{name}/F"] = fp, A @ fp @ A.T
    # transform(C, exp)
    r = np.linspace(-1, 1, 10)
    N = 10
    out["transform_exp/points"], out["transform_exp/N"] = r, np.array(N)
    out["transform_exp/c"] = O.c1_expand(np.exp, np.exp, r, N)
    # analysis_2D
    n, p = 3, 12
    z, T = O.legendre_grid_transform(p)
    coef = rng.standard_normal((p, n, p, n))                        # [a, i, b, j]
    Pz = np.polynomial.legendre.legvander(z, p - 1)
    out["analysis/n"], out["analysis/p"] = np.array(n), np.array(p)
    out["analysis/vals"] = np.einsum("sa,aibj,tb->isjt", Pz, coef, Pz).reshape(n * p, n * p)
    out["analysis/F"] = coef.reshape(n * p, n * p)
    # inverse transforms (`Pl \\ cfs`, src/continuouspolynomial.jl:129-184): random coefficients, values by EVALUATING the basis
    # functions at the cell nodes (basis_eval_matrix: hats and bubbles W_n themselves, no conversion to Legendre coefficients)
    for name, basis, pts, p in (("itransform_c1", O.C1, np.linspace(-1, 1, 6), 9), ("itransform_dirichlet", O.DIRICHLET, np.linspace(0, 2, 5), 12)):
        m = len(pts) - 1
        nh = m + 1 if basis == O.C1 else m - 1
        n1 = nh + m * (p - 2)
        z, _ = O.legendre_grid_transform(p)
        x = ((pts[:-1] + pts[1:])[:, None] / 2 + np.diff(pts)[:, None] / 2 * z[None, :]).reshape(-1)
        E = O.basis_eval_matrix(basis, pts, p - 1, x)
        c1 = rng.standard_normal(n1)
        c2 = rng.standard_normal((n1, n1))
        out[f"{name}/points"], out[f"{name}/p"] = pts, np.array(p)
        out[f"{name}/c"], out[f"{name}/vals"] = c1, E @ c1
        out[f"{name}/C2"], out[f"{name}/vals2"] = c2, E @ c2 @ E.T
    # weak Laplacian with a piecewise-constant diffusion coefficient (examples/heat.jl:77-79) on a graded mesh: Gauss quadrature
    # of a(x) phi_i' phi_j' (not the closed-form bands)
    pts = np.array([-1, -0.7, -0.2, 0.1, 0.6, 1.0])
    a = np.array([1.0, 0.5, 2.0, 0.25, 1.5])
    for name, basis in (("varcoef_c1", O.C1), ("varcoef_dirichlet", O.DIRICHLET)):
        _, Lq = O.quadrature_matrices(basis, len(pts) - 1, 7, np.diff(pts), a=a)
        out[f"{name}/points"], out[f"{name}/a"], out[f"{name}/N"], out[f"{name}/L"] = pts, a, np.array(7), Lq
    np.savez_compressed(os.path.join(HERE, "rhs_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
