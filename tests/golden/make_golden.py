#!/usr/bin/env python
"""Generates tests/golden/arrowhead_golden.npz and golden_meta.json.

The reference is Julia and cannot run in this container, so the fixtures are
  (i)  the known answers the reference's own tests hold for this path, copied as literals
       (test/test_arrowhead.jl:134 and the structural numbers of :156-159,171-172), and
  (ii) seeded input/output vectors produced by the DENSE level of the oracle (NumPy/LAPACK on the
       dense Matrix(A), which is exactly the oracle of every reference testset): reverse Cholesky
       via flip(cholesky(flip(S))), dense solves, dense products, dense ADI.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import arrowhead_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dense_case(basis, m, N, sigma, ncols, seed):
    h = 2.0 / m
    Mp, Lp = O.assemble_mass(basis, m, N, h), O.assemble_weaklaplacian(basis, m, N, h)
    Sp = O.packed_axpby(-1.0, Lp, sigma, Mp)
    S = O.unpack(O.packed_symmetrize(Sp)).to_dense()
    U = O.dense_reversecholesky(S)
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((S.shape[0], ncols))
    X = np.linalg.solve(S, B)
    Mden = O.unpack(O.packed_symmetrize(Mp)).to_dense()
    return {"B": B, "X": X, "U": U, "MB": Mden @ B, "BtM": (B.T @ Mden)}


def main():
    out = {}
    meta = {"reference_goldens": {"helmholtz_value_at_0.1": 1.1952730862177243, "source": "test/test_arrowhead.jl:134",
                                  "blockbandwidths": [3, 2], "bandwidths": [13, 9], "source_structure": "test/test_arrowhead.jl:156-159",
                                  "bandwidths_M_m3_N5": [7, 7], "bandwidths_Delta_m3_N5": [1, 1], "source_bw": "test/test_arrowhead.jl:171-172"},
            "cases": {}}
    cases = {"readme_c1_m3_N10": (O.C1, 3, 10, 1.0, 3, 11), "dirichlet_m4_N12_s0.5": (O.DIRICHLET, 4, 12, 0.5, 4, 12),
             "c1_m16_N9_s1e4": (O.C1, 16, 9, 1e4, 5, 13), "dirichlet_m12_N40_s3": (O.DIRICHLET, 12, 40, 3.0, 6, 14),
             "dirichlet_m6_N70_s1": (O.DIRICHLET, 6, 70, 1.0, 3, 15)}
    for name, (basis, m, N, sigma, ncols, seed) in cases.items():
        d = dense_case(basis, m, N, sigma, ncols, seed)
        for k, v in d.items():
            out[f"{name}/{k}"] = v
        meta["cases"][name] = {"basis": "c1" if basis == O.C1 else "dirichlet", "m": m, "N": N, "sigma": sigma, "ncols": ncols, "seed": seed}
    # dense ADI (examples/adi_poisson.jl:42-59) at m=4, N=8
    m, N = 4, 8
    h = 2.0 / m
    Mp, Lp = O.assemble_mass(O.DIRICHLET, m, N, h), O.assemble_weaklaplacian(O.DIRICHLET, m, N, h)
    Md = O.unpack(O.packed_symmetrize(Mp)).to_dense()
    Kd = -O.unpack(O.packed_symmetrize(Lp)).to_dense()
    rng = np.random.default_rng(21)
    Y = rng.standard_normal(Md.shape)
    F = Kd @ Y @ Md + Md @ Y @ Kd
    lam = np.linalg.eigvals(np.linalg.solve(Kd, Md)).real
    a, b = float(lam.min() * (1 - 1e-9)), float(lam.max() * (1 + 1e-9))
    X = O.adi_dense(Md, -Md, Kd, F, a, b, -b, -a, tol=1e-15)
    out["adi_dirichlet_m4_N8/F"] = F
    out["adi_dirichlet_m4_N8/X"] = X
    meta["cases"]["adi_dirichlet_m4_N8"] = {"m": m, "N": N, "a": a, "b": b, "J": int(O.adi_num_iterations(a, b, -b, -a, 1e-15)), "seed": 21}
    np.savez_compressed(os.path.join(HERE, "arrowhead_golden.npz"), **out)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
