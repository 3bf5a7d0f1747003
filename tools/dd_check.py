"""Element-decomposed ADI (csrc/pop_row.cu): correctness against the single-GPU driver and timing.
  python tools/dd_check.py                                   # 1 GPU
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/dd_check.py [m N]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402
from pop_b200 import _lib as L  # noqa: E402
from pop_b200.arrowhead import _bind_stream  # noqa: E402


def colmajor(rows, cols, dev, gen=None):
    ld = (rows + 1) & ~1
    buf = torch.zeros((cols, ld), dtype=torch.float64, device=dev)
    if gen is not None:
        buf.copy_(torch.randn((cols, ld), generator=gen, dtype=torch.float64, device=dev))
    return buf.t()[:rows]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    P.default_context(lr)

    def problem(basis, m, N):
        pts = np.linspace(-1, 1, m + 1)
        Q = P.DirichletPolynomial(pts) if basis == "dirichlet" else P.ContinuousPolynomial(1, pts)
        KR = P.Block.oneto(N)
        return P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]

    worst = 0.0
    small = () if os.environ.get("DD_SKIP_SMALL") == "1" else (("dirichlet", 8, 12), ("dirichlet", 16, 9))
    for basis, m, N in small:
        M, K = problem(basis, m, N)
        n = M.n
        g = torch.Generator(device=dev).manual_seed(7)
        Ffull = colmajor(n, n, dev, g)                    # same on every rank
        dd = P.ElementDecomposition(M)
        cols = torch.from_numpy(dd.columns).to(dev)
        # one row half step against the generic kernels:  R' = (R S^-1) T + gamma F
        S = K + 3.7 * M
        T = 0.5 * M - K
        Fs = P.reversecholesky(S)
        Rfull = colmajor(n, n, dev, g)
        want = (P.rdiv(Rfull, Fs) @ T) - 0.4 * Ffull
        Rl, Fl = dd.panel(), dd.panel()
        Rl.copy_(Rfull[:, cols])
        Fl.copy_(Ffull[:, cols])
        ctx = M.data.ctx
        _bind_stream(ctx)
        L.check(L.lib().pop_dd_row_half_step(ctx.handle, dd._h, Fs.factor._h, T.data._h, -0.4, Fl.data_ptr(), Fl.stride(1), Rl.data_ptr(), Rl.stride(1)), "row half step")
        torch.cuda.synchronize()
        e_half = float(torch.linalg.norm(Rl - want[:, cols]) / torch.linalg.norm(want))
        a, b = P.spectrum_bounds(M, K)
        J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
        Xref = P.adi_poisson(M, K, Ffull, a, b, J=J)
        Xl = P.adi_poisson_dd(M, K, Fl, a, b, J=J, dd=dd)
        e_adi = float(torch.linalg.norm(Xl - Xref[:, cols]) / torch.linalg.norm(Xref))
        errs = torch.tensor([e_half, e_adi], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} {basis} m={m} N={N} n={n}: row half step err {errs[0]:.2e}, ADI (element decomposition) vs single-GPU {errs[1]:.2e}", flush=True)
        worst = max(worst, float(errs.max()))
        dd.close()
    assert worst < 1e-11, worst

    m, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
    M, K = problem("dirichlet", m, N)
    n = M.n
    dd = P.ElementDecomposition(M)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    Fl = dd.panel()
    Fl.copy_(torch.randn((n, dd.nloc), generator=g, dtype=torch.float64, device=dev))
    b = 4 / np.pi ** 2 * (1 + 1e-9)
    a = 3.27e-11 if (m, N) == (128, 64) else 1e-9     # 1 / lambda_max(K, M) of BASELINE config 4: J = 90
    J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(label, reps=3):
        for it in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            P.adi_poisson_dd(M, K, Fl, a, b, J=J, dd=dd)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            s = float(t.item())
            print(f"world={world} n={n} J={J} ADI [element decomposition]{label}: {s * 1e3:.2f} ms/solve, {2.0 * J * n * n / s / 1e9:.1f} G unknown*RHS/s, "
                  f"{48.0 * n * n * J / world / s / 1e9:.0f} GB/s algorithmic per GPU ({48.0 * n * n * J / world / s / 1e9 / 6549.8:.3f} of the measured HBM peak)", flush=True)

    timed("")
    # DD_VARIANTS="A=1,B=2;C=3": the same solve under other environment switches of the kernels (read at every launch)
    for var in filter(None, os.environ.get("DD_VARIANTS", "").split(";")):
        kv = dict(x.split("=") for x in var.split(","))
        os.environ.update(kv)
        timed(f" [{var}]")
        for k in kv:
            os.environ.pop(k)
    if os.environ.get("DD_TRACE") == "1":
        os.environ["POP_ROW2_TRACE"] = "1"
        if rank in (0, world // 2):
            sys.stderr.write(f"rank {rank}:\n")
        P.adi_poisson_dd(M, K, Fl, a, b, J=3, dd=dd)
        torch.cuda.synchronize()
        os.environ.pop("POP_ROW2_TRACE")
    dd.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
