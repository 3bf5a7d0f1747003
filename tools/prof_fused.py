"""Runs a few fused solves of one shape (target program for `ncu --set full`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402

shapes = {"cfg2": ("c1", 1024, 40, 4096), "cfg3": ("dirichlet", 256, 128, 1024), "cfg4": ("dirichlet", 128, 64, 8191), "cfg5": ("dirichlet", 256, 64, 4096)}
basis, m, N, nrhs = shapes[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
if len(sys.argv) > 2:
    nrhs = int(sys.argv[2])
Bs = P.ContinuousPolynomial(1, np.linspace(-1, 1, m + 1)) if basis == "c1" else P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
KR = P.Block.oneto(N)
S = (-P.weaklaplacian(Bs) + P.grammatrix(Bs))[KR, KR]
F = P.reversecholesky(S)
n = S.n
ld = (n + 1) & ~1
Bb = torch.randn((nrhs, ld), dtype=torch.float64, device="cuda")
Xb = torch.empty_like(Bb)
for _ in range(4):
    Xb.copy_(Bb)
    P.ldiv(F, Xb.t()[:n], overwrite=True, algo=P.SOLVE_FUSED)
torch.cuda.synchronize()
print("ok", n, nrhs)
