"""Times the pieces of one ADI half step at BASELINE config 4 (Dirichlet m=128, p=64, n=8191):
fused solve+product kernel, transpose, batched shifted factorisation, and the whole ADI solve."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402
from pop_b200 import _lib as L  # noqa: E402
from pop_b200.arrowhead import _bind_stream  # noqa: E402

PEAK = 6549.8


def ev_time(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps + 2):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


def main():
    m, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
    Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
    KR = P.Block.oneto(N)
    M = P.grammatrix(Q)[KR, KR]
    K = -P.weaklaplacian(Q)[KR, KR]
    n = M.n
    ld = (n + 1) & ~1
    lib = L.lib()
    ctx = M.data.ctx
    _bind_stream(ctx)          # kernels on torch's current stream, so torch events time them
    F1 = P.reversecholesky(K + 3.7 * M)
    T = (0.5 * M - K)
    R = torch.randn((n, ld), dtype=torch.float64, device="cuda").t()[:n]
    Fm = torch.randn((n, ld), dtype=torch.float64, device="cuda").t()[:n]
    V = torch.empty((n, ld), dtype=torch.float64, device="cuda").t()[:n]
    R0 = R.clone()
    ent = float(n) * n

    def half():
        L.check(lib.pop_solve_mul_axpy(ctx.handle, F1.factor._h, T.data._h, -0.4, Fm.data_ptr(), Fm.stride(1), R.data_ptr(), R.stride(1), n))
    for name, env in (("fused-reg", {}), ("fused-reg npar=2", {"POP_F2_NPAR": "2"}), ("fused-reg npar=2 occ=1", {"POP_F2_NPAR": "2", "POP_F2_OCC": "1"}),
                      ("fused-smem", {"POP_DISABLE_FUSED2": "1"})):
        for k_ in ("POP_DISABLE_FUSED2", "POP_F2_NPAR", "POP_F2_OCC"):
            os.environ.pop(k_, None)
        os.environ.update(env)
        R.copy_(R0)
        ms = ev_time(half)
        print(f"n={n} half step solve+mul+axpy [{name}]: {ms:.3f} ms  {24 * ent / ms / 1e6:.0f} GB/s ({24 * ent / ms / 1e6 / PEAK * 100:.1f}% of measured peak, 24 B/entry)", flush=True)
    for k_ in ("POP_DISABLE_FUSED2", "POP_F2_NPAR", "POP_F2_OCC"):
        os.environ.pop(k_, None)
    Rchk = torch.empty((n, ld), dtype=torch.float64, device="cuda").t()[:n]; Rchk.copy_(R0)
    L.check(lib.pop_solve_mul_axpy(ctx.handle, F1.factor._h, T.data._h, -0.4, Fm.data_ptr(), Fm.stride(1), Rchk.data_ptr(), Rchk.stride(1), 64))
    os.environ["POP_F2_NPAR"] = "2"
    Rchk2 = torch.empty((n, ld), dtype=torch.float64, device="cuda").t()[:n]; Rchk2.copy_(R0)
    L.check(lib.pop_solve_mul_axpy(ctx.handle, F1.factor._h, T.data._h, -0.4, Fm.data_ptr(), Fm.stride(1), Rchk2.data_ptr(), Rchk2.stride(1), 64))
    os.environ.pop("POP_F2_NPAR", None)
    print("npar=2 vs npar=1 half step rel diff:", float(torch.linalg.norm(Rchk2[:, :64] - Rchk[:, :64]) / torch.linalg.norm(Rchk[:, :64])), flush=True)

    def solve_only():
        L.check(lib.pop_solve(ctx.handle, F1.factor._h, 0, R.data_ptr(), R.stride(1), n, 0))
    ms = ev_time(solve_only)
    print(f"n={n} plain fused solve: {ms:.3f} ms  {16 * ent / ms / 1e6:.0f} GB/s ({16 * ent / ms / 1e6 / PEAK * 100:.1f}%)", flush=True)
    os.environ["POP_F2_NPAR"] = "2"
    ms = ev_time(solve_only)
    print(f"n={n} plain fused solve npar=2: {ms:.3f} ms  {16 * ent / ms / 1e6:.0f} GB/s ({16 * ent / ms / 1e6 / PEAK * 100:.1f}%)", flush=True)
    os.environ.pop("POP_F2_NPAR", None)

    def tr():
        L.check(lib.pop_transpose(ctx.handle, R.data_ptr(), R.stride(1), n, n, V.data_ptr(), V.stride(1)))
    ms = ev_time(tr)
    print(f"n={n} transpose: {ms:.3f} ms  {16 * ent / ms / 1e6:.0f} GB/s ({16 * ent / ms / 1e6 / PEAK * 100:.1f}%, 16 B/entry)", flush=True)
    ms = ev_time(lambda: V.copy_(R))
    print(f"n={n} torch copy (reference point): {ms:.3f} ms  {16 * ent / ms / 1e6:.0f} GB/s", flush=True)

    a, b = 3.1e-10, 4 / np.pi ** 2 * (1 + 1e-9)
    J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
    p, q = P.adi_shifts(J, a, b, -b, -a)
    sg = np.empty(2 * J)
    sg[0::2] = 1.0 / p
    sg[1::2] = -1.0 / q
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    facs = P.reversecholesky_shifted(K, M, np.ones(2 * J), sg)
    torch.cuda.synchronize()
    print(f"n={n} J={J}: batched factorisation of {2 * J} shifts: {(time.perf_counter() - t0) * 1e3:.2f} ms (wall)", flush=True)
    del facs
    for it in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        X = P.adi_poisson(M, K, Fm, a, b, J=J)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"n={n} J={J}: ADI solve {dt * 1e3:.1f} ms  -> {48 * ent * J / dt / 1e9:.0f} GB/s algorithmic (48 B/entry/iteration), {2 * J * ent / dt / 1e9:.1f} G unknown*RHS/s", flush=True)


if __name__ == "__main__":
    main()
