"""Times the fused / staged solve kernels for a few shapes and cluster sizes (CUDA events).
Scratch tool for kernel tuning; not part of the bench contract."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402


def time_solve(F, X, B, reps=5, algo=P.SOLVE_AUTO):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps + 2):
        X.copy_(B)
        e0.record()
        P.ldiv(F, X, overwrite=True, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


def main():
    shapes = [("c1", 1024, 40, 4096), ("dirichlet", 256, 128, 1024), ("dirichlet", 128, 64, 8191), ("dirichlet", 256, 64, 4096)]
    if len(sys.argv) > 1:
        shapes = [shapes[int(a)] for a in sys.argv[1].split(",")]
    for basis, m, N, nrhs in shapes:
        Bs = P.ContinuousPolynomial(1, np.linspace(-1, 1, m + 1)) if basis == "c1" else P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
        KR = P.Block.oneto(N)
        S = (-P.weaklaplacian(Bs) + P.grammatrix(Bs))[KR, KR]
        F = P.reversecholesky(S)
        n = S.n
        ld = (n + 1) & ~1
        Bb = torch.randn((nrhs, ld), dtype=torch.float64, device="cuda")
        Xb = torch.empty_like(Bb)
        B, X = Bb.t()[:n], Xb.t()[:n]
        gb = 16.0 * n * nrhs / 1e9
        for cs in os.environ.get("TUNE_CS", "1,2,4,8").split(","):
            os.environ["POP_FUSED_CS"] = cs
            try:
                ms = time_solve(F, X, B, algo=P.SOLVE_FUSED)
                print(f"{basis} m={m} N={N} n={n} nrhs={nrhs} fused cs={cs}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb / ms * 1e3 / 6549.8 * 100:.1f}% of measured peak)", flush=True)
            except Exception as ex:
                print(f"{basis} m={m} N={N} cs={cs}: {type(ex).__name__}: {str(ex)[:80]}", flush=True)
        os.environ.pop("POP_FUSED_CS", None)
        ms = time_solve(F, X, B, algo=P.SOLVE_STAGED)
        print(f"{basis} m={m} N={N} n={n} nrhs={nrhs} staged: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s (16 B/unknown accounting)", flush=True)


if __name__ == "__main__":
    main()
