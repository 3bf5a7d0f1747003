"""Phase breakdown of the register-resident fused solve (debug build only):
  make -C piecewiseorthogonalpolynomials.jl_b200 EXTRA_NVFLAGS=-DF2_PROF  &&  python tools/prof_f2_phases.py [shape indices]
Block 0's warp 0 (a scan warp) and last warp accumulate SM cycles per phase of the column loop (clock64)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402


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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(4):
            X.copy_(B)
            if rep == 3:
                os.environ["POP_F2_PROF_PRINT"] = "1"
                print(f"--- {basis} m={m} N={N} n={n} nrhs={nrhs}", file=sys.stderr, flush=True)
            e0.record()
            P.ldiv(F, X, overwrite=True, algo=P.SOLVE_FUSED)
            e1.record()
            torch.cuda.synchronize()
            os.environ.pop("POP_F2_PROF_PRINT", None)
        print(f"{basis} m={m} N={N}: {e0.elapsed_time(e1):.3f} ms (profiling build, includes the print on the last repetition)", flush=True)


if __name__ == "__main__":
    main()
