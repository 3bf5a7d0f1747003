"""e2e solve through pop_solve_host on PAGEABLE host memory (what a Julia Array is): the staged pipeline (worker threads,
pinned pieces) against the driver's own pageable path, checked against the pinned result; and the ADI through
pop_adi_poisson_host with pinned and pageable F / X."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P
from pop_b200 import _lib as L
m, N, nrhs = 1024, 40, 4096
C = P.ContinuousPolynomial(1, np.linspace(-1, 1, m + 1))
KR = P.Block.oneto(N)
F = P.reversecholesky((-P.weaklaplacian(C) + P.grammatrix(C))[KR, KR])
n = F.factor.n
hB = torch.randn((nrhs, n), dtype=torch.float64).pin_memory()
hX = torch.empty((nrhs, n), dtype=torch.float64).pin_memory()
hX.copy_(hB); P.ldiv(F, hX.numpy().T, overwrite=True)
want = hX.clone()
pg = torch.empty((nrhs, n), dtype=torch.float64)           # pageable
Xp = pg.numpy().T
for stage, thr in (("1", "6"), ("1", "4"), ("1", "8"), ("1", "3"), ("0", "6")):
    os.environ["POP_HOST_STAGE"] = stage
    os.environ["POP_HOST_THREADS"] = thr
    ts = []
    for _ in range(5):
        pg.copy_(hB); torch.cuda.synchronize()
        t0 = time.perf_counter(); P.ldiv(F, Xp, overwrite=True); ts.append(time.perf_counter() - t0)
    err = float((pg - want).abs().max())
    print(f"pageable stage={stage} threads={thr}: " + " ".join(f"{t * 1e3:.1f}" for t in ts) + f" ms -> median of last 3 {np.median(ts[2:]) * 1e3:.2f} ms, "
          f"{n * nrhs / np.median(ts[2:]) / 1e9:.2f} G unknown*RHS/s, max |diff| vs pinned {err:.1e}", flush=True)
os.environ["POP_HOST_STAGE"] = "1"
os.environ.pop("POP_HOST_THREADS")
del hB, hX, pg, want

# ADI, host F in, host X out
m, N = 128, 64
Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
KR = P.Block.oneto(N)
M, K = P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]
n = M.n
a, b = 3.27e-11, 4 / np.pi ** 2 * (1 + 1e-9)
J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
ref = None
for kind in ("pinned", "pageable"):
    Fh = torch.randn((n, n), dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    Xh = torch.empty((n, n), dtype=torch.float64)
    if kind == "pinned":
        Fh, Xh = Fh.pin_memory(), Xh.pin_memory()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        P.adi_poisson(M, K, Fh.numpy().T, a, b, J=J, out=Xh.numpy().T)
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = Xh.clone()
    print(f"ADI host ({kind}) n={n} J={J}: " + " ".join(f"{t * 1e3:.1f}" for t in ts) + f" ms, max |diff| vs pinned {float((Xh - ref).abs().max()):.1e}", flush=True)
