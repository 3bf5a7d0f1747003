"""e2e solve through pop_solve_host (pinned host panel): the flat address-aligned pipeline against the column-chunk one."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P
m, N, nrhs = 1024, 40, 4096
C = P.ContinuousPolynomial(1, np.linspace(-1, 1, m + 1))
KR = P.Block.oneto(N)
F = P.reversecholesky((-P.weaklaplacian(C) + P.grammatrix(C))[KR, KR])
n = F.factor.n
hB = torch.randn((nrhs, n), dtype=torch.float64).pin_memory()
hX = torch.empty((nrhs, n), dtype=torch.float64).pin_memory()
Xh = hX.numpy().T
for flat in ("1", "0", "1"):
    os.environ["POP_HOST_FLAT"] = flat
    ts = []
    for _ in range(6):
        hX.copy_(hB); torch.cuda.synchronize()
        t0 = time.perf_counter(); P.ldiv(F, Xh, overwrite=True); ts.append(time.perf_counter() - t0)
    print(f"flat={flat}: " + " ".join(f"{t * 1e3:.1f}" for t in ts) + f" ms -> median of last 4 {np.median(ts[2:]) * 1e3:.2f} ms, {n * nrhs / np.median(ts[2:]) / 1e9:.2f} G unknown*RHS/s", flush=True)
