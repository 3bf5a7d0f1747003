"""Times the register-resident fused solve (pop_fused2.cu) over cluster sizes / occupancies, next to
the shared-memory fused kernel and the staged kernels (CUDA events).  Scratch tool for kernel tuning."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402

PEAK = 6549.8


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
        Xref = None
        tag = f"{basis} m={m} N={N} n={n} nrhs={nrhs}"
        ms = time_solve(F, X, B, algo=P.SOLVE_STAGED)
        Xref = X.clone()
        print(f"{tag} staged: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s (16 B/unknown accounting)", flush=True)
        for cs in os.environ.get("TUNE_CS", "1,2,4,8").split(","):
            for occ in os.environ.get("TUNE_OCC", "0,1,2,3").split(","):
                os.environ["POP_F2_CS"] = cs
                os.environ["POP_F2_OCC"] = occ
                try:
                    ms = time_solve(F, X, B, algo=P.SOLVE_FUSED)
                    err = float(torch.linalg.norm(X - Xref) / torch.linalg.norm(Xref))
                    print(f"{tag} fused-reg cs={cs} occ={occ}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb / ms * 1e3 / PEAK * 100:.1f}% of measured peak)  diff-vs-staged {err:.1e}", flush=True)
                except Exception as ex:
                    print(f"{tag} fused-reg cs={cs} occ={occ}: {type(ex).__name__}: {str(ex)[:90]}", flush=True)
        os.environ.pop("POP_F2_CS", None)
        os.environ.pop("POP_F2_OCC", None)
        try:
            ms = time_solve(F, X, B, algo=P.SOLVE_FUSED_SMEM)
            print(f"{tag} fused-smem (auto): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb / ms * 1e3 / PEAK * 100:.1f}% of measured peak)", flush=True)
        except Exception as ex:
            print(f"{tag} fused-smem: {type(ex).__name__}: {str(ex)[:90]}", flush=True)


if __name__ == "__main__":
    main()
