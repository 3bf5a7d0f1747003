"""Single-pass TMA row half step (csrc/pop_row2.cu) at full size on one GPU: agreement with the two-pass row kernels of
csrc/pop_row.cu and timing of both, then the whole ADI.
  python tools/row2_check.py [m N] [--adi]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402
from pop_b200 import _lib as L  # noqa: E402
from pop_b200.arrowhead import _bind_stream  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    m, N = (int(args[0]), int(args[1])) if len(args) >= 2 else (128, 64)
    basis = os.environ.get("ROW_BASIS", "dirichlet")
    torch.cuda.set_device(0)
    P.default_context(0)
    pts = np.linspace(-1, 1, m + 1)
    Q = P.DirichletPolynomial(pts) if basis == "dirichlet" else P.ContinuousPolynomial(1, pts)
    KR = P.Block.oneto(N)
    M, K = P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]
    if basis != "dirichlet":
        K = K + 0.3 * M
    n = M.n
    dd = P.ElementDecomposition(M)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    R0, Fd, R = dd.panel(), dd.panel(), dd.panel()
    R0.copy_(torch.randn((n, n), generator=g, dtype=torch.float64, device="cuda:0"))
    Fd.copy_(torch.randn((n, n), generator=g, dtype=torch.float64, device="cuda:0"))
    Fs = P.reversecholesky(K + 3.7 * M)
    T = 0.5 * M - K
    ctx = M.data.ctx
    _bind_stream(ctx)
    lib = L.lib()
    os.environ["POP_ROW1"] = "0"

    def half(mul):
        if mul:
            L.check(lib.pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, T.data._h, -0.4, Fd.data_ptr(), Fd.stride(1), R.data_ptr(), R.stride(1)))
        else:
            L.check(lib.pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, None, 0.0, None, 0, R.data_ptr(), R.stride(1)))

    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for mul in (True, False):
        for mode in ("0", "1"):                      # 0: two-pass kernels, 1: k_row2
            os.environ["POP_ROW2"] = mode
            ts = []
            for it in range(6):
                R.copy_(R0)
                torch.cuda.synchronize()
                e0.record()
                half(mul)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out[(mul, mode)] = R.clone()
            t = min(ts[1:])
            byt = (24 if mul else 16) * n * n
            print(f"{basis} m={m} N={N} n={n} row half step ({'solve+product+F' if mul else 'plain solve'}) "
                  f"{'k_row2' if mode == '1' else 'two-pass'}: {t * 1e3:.1f} us, {byt / t / 1e6:.0f} GB/s algorithmic ({byt / t / 1e6 / 6549.8:.3f} of measured peak)", flush=True)
        d = float(torch.linalg.norm(out[(mul, "1")] - out[(mul, "0")]) / torch.linalg.norm(out[(mul, "0")]))
        print(f"   k_row2 vs two-pass: rel diff {d:.2e}", flush=True)
        assert d < 1e-12, d
    if "--adi" in sys.argv:
        Kd = -P.weaklaplacian(Q)[KR, KR]
        a, b = P.spectrum_bounds(M, Kd)
        J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
        res = {}
        for mode in ("0", "1"):
            os.environ["POP_ROW2"] = mode
            for it in range(3):
                torch.cuda.synchronize()
                e0.record()
                X = P.adi_poisson_dd(M, Kd, Fd, a, b, J=J, dd=dd)
                e1.record()
                torch.cuda.synchronize()
            s = e0.elapsed_time(e1) * 1e-3
            res[mode] = X.clone()
            print(f"ADI n={n} J={J} [{'k_row2' if mode == '1' else 'two-pass'} rows]: {s * 1e3:.1f} ms/solve, {48.0 * n * n * J / s / 1e9:.0f} GB/s algorithmic", flush=True)
        d = float(torch.linalg.norm(res["1"] - res["0"]) / torch.linalg.norm(res["0"]))
        print(f"   ADI k_row2 vs two-pass rows: rel diff {d:.2e}", flush=True)
    dd.close()


if __name__ == "__main__":
    main()
