"""Times the right-hand-side pipeline (csrc/pop_rhs.cu) at the BASELINE config-4 shape: n = 128 cells, p = 64 nodes
per cell (8192 x 8192 values), Dirichlet basis.  python tools/prof_rhs.py [n p]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402


def timeit(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]) * 1e-3, out


def main():
    n, p = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
    torch.cuda.set_device(0)
    P.default_context(0)
    N = n * p
    z, T = P.legendre_grid_transform(p)
    g = torch.Generator(device="cuda").manual_seed(0)
    vals = torch.randn((N, N), generator=g, dtype=torch.float64, device="cuda").t()      # column-major
    t, fp = timeit(lambda: P.analysis_2d(vals, n, p, T))
    flops = 4.0 * p * N * N
    print(f"analysis_2D n={n} p={p} ({N}x{N}): {t * 1e3:.3f} ms, {flops / t / 1e12:.2f} TFLOP/s FP64, {16.0 * N * N / t / 1e9:.0f} GB/s algorithmic", flush=True)
    pts = np.linspace(-1, 1, n + 1)
    Q = P.DirichletPolynomial(pts)
    KR = P.Block.oneto(p)
    nq = Q.nh + Q.m * (p - 1)
    t1, Z = timeit(lambda: P.weakform_apply(Q, KR, fp, P.LEFT))
    t2, F = timeit(lambda: P.weakform_apply(Q, KR, Z, P.RIGHT))
    print(f"(Q'P) X  left : {t1 * 1e3:.3f} ms, {8.0 * (N + nq) * N / t1 / 1e9:.0f} GB/s", flush=True)
    print(f"X (Q'P)' right: {t2 * 1e3:.3f} ms, {8.0 * (N + nq) * nq / t2 / 1e9:.0f} GB/s", flush=True)
    # CPU reference of the analysis step on a sample of cell columns (NumPy einsum, all host cores through BLAS)
    if os.environ.get("RHS_CPU", "1") == "1":
        from oracle import arrowhead_oracle as O
        ns = min(n, 8)
        v = np.random.default_rng(0).standard_normal((ns * p, ns * p))
        t0 = time.perf_counter()
        O.analysis_2d(v, ns, p, T)
        dt = time.perf_counter() - t0
        print(f"oracle analysis_2d on {ns}x{ns} cells: {dt * 1e3:.1f} ms -> {dt * (n / ns) ** 2 * 1e3:.0f} ms for the full grid (NumPy)", flush=True)


if __name__ == "__main__":
    main()
