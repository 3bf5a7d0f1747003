"""Runs a few fused ADI half steps (solve + product + gamma*F) at config-4 shape: target for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402
from pop_b200 import _lib as L  # noqa: E402
from pop_b200.arrowhead import _bind_stream  # noqa: E402

m, N = 128, 64
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
KR = P.Block.oneto(N)
M = P.grammatrix(Q)[KR, KR]
K = -P.weaklaplacian(Q)[KR, KR]
n = M.n
ld = (n + 1) & ~1
ctx = M.data.ctx
_bind_stream(ctx)
F1 = P.reversecholesky(K + 3.7 * M)
T = (0.5 * M - K)
R = torch.randn((ncols, ld), dtype=torch.float64, device="cuda").t()[:n]
Fm = torch.randn((ncols, ld), dtype=torch.float64, device="cuda").t()[:n]
for _ in range(4):
    L.check(L.lib().pop_solve_mul_axpy(ctx.handle, F1.factor._h, T.data._h, -0.4, Fm.data_ptr(), Fm.stride(1), R.data_ptr(), R.stride(1), ncols))
torch.cuda.synchronize()
print("ok", n, ncols)
