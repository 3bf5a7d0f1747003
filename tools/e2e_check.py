"""e2e solve through pop_solve_host (pinned host panel) for a few chunk sizes."""
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
for nbuf, mb in ((3, 32), (4, 32), (6, 32), (4, 16), (6, 16), (8, 8), (4, 64), (3, 16)):
    os.environ["POP_HOST_CHUNK_MB"] = str(mb)
    os.environ["POP_HOST_NBUF"] = str(nbuf)
    best = 1e9
    for _ in range(4):
        hX.copy_(hB); torch.cuda.synchronize()
        t0 = time.perf_counter(); P.ldiv(F, Xh, overwrite=True); best = min(best, time.perf_counter() - t0)
    print(f"{nbuf} buffers, chunk {mb} MB: {best * 1e3:.1f} ms  {n * nrhs / best / 1e9:.2f} G unknown*RHS/s", flush=True)
