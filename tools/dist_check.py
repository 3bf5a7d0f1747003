"""Multi-GPU check + timing of the distributed ADI (run under torchrun, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [m N]
1. distributed transpose (fused transpose+exchange kernel over peer memory) against a gathered reference,
2. distributed ADI (p2p and NCCL exchange) against the single-GPU ADI on rank 0 (small problem),
3. timing of the BASELINE config 4 problem (Dirichlet m=128, p=64)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402


def colmajor(rows, cols, dev, gen=None):
    ld = (rows + 1) & ~1
    buf = torch.zeros((cols, ld), dtype=torch.float64, device=dev)
    if gen is not None:
        buf.copy_(torch.randn((cols, ld), generator=gen, dtype=torch.float64, device=dev))
    return buf.t()[:rows]


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = f"cuda:{lr}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    ctx = P.default_context(lr)

    def problem(m, N):
        Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
        KR = P.Block.oneto(N)
        return P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]

    # ---- 1+2: correctness on a small problem ------------------------------------------------------
    m, N = 8, 12
    M, K = problem(m, N)
    n = M.n
    g = torch.Generator(device=dev).manual_seed(7)
    Ffull = colmajor(n, n, dev, g)                     # same on every rank (same seed)
    grid = P.DistMatrix(n, M.data.ctx)
    lo, hi = grid.lo, grid.hi
    W = grid.buffer(0)
    W.copy_(Ffull[:, lo:hi])
    grid.transpose(0, 1)
    torch.cuda.synchronize()
    err_t = float(torch.linalg.norm(grid.buffer(1) - Ffull.t()[:, lo:hi]) / torch.linalg.norm(Ffull))
    a, b = P.spectrum_bounds(M, K)
    J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
    Xref = P.adi_poisson(M, K, Ffull, a, b, J=J)
    Xp = P.adi_poisson_p2p(M, K, _cm(Ffull[:, lo:hi], dev), a, b, J=J, grid=grid)
    err_p = float(torch.linalg.norm(Xp - Xref[:, lo:hi]) / torch.linalg.norm(Xref))
    Xn = P.adi_poisson_distributed(M, K, _cm(Ffull[:, lo:hi], dev), _cm(Ffull.t()[:, lo:hi], dev), a, b, J=J)
    err_n = float(torch.linalg.norm(Xn - Xref[:, lo:hi]) / torch.linalg.norm(Xref))
    errs = torch.tensor([err_t, err_p, err_n], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} n={n}: transpose err {errs[0]:.2e}, ADI p2p vs single-GPU {errs[1]:.2e}, ADI nccl vs single-GPU {errs[2]:.2e}", flush=True)
    assert errs[0] == 0.0 and errs[1] < 1e-12 and errs[2] < 1e-12, errs
    grid.close()

    # ---- 3: timing at config-4 size ------------------------------------------------------------------
    if len(sys.argv) > 2:
        m, N = int(sys.argv[1]), int(sys.argv[2])
    else:
        m, N = 128, 64
    M, K = problem(m, N)
    n = M.n
    grid = P.DistMatrix(n, M.data.ctx)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    Fl = colmajor(n, grid.w, dev, g)
    b = 4 / np.pi ** 2 * (1 + 1e-9)
    a = 3.1e-10 if (m, N) == (128, 64) else 1e-9
    J = P.adi_num_iterations(a, b, -b, -a, 1e-15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    quick = os.environ.get("DIST_QUICK", "0") == "1"
    cases = [("p2p", lambda: P.adi_poisson_p2p(M, K, Fl, a, b, J=J, grid=grid))]
    if not quick:
        cases.append(("nccl", lambda: P.adi_poisson_distributed(M, K, Fl, Fl, a, b, J=J)))
    for name, fn in cases:
        for it in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            s = float(t.item())
            print(f"world={world} n={n} J={J} ADI [{name}]: {s * 1e3:.1f} ms/solve, {2.0 * J * n * n / s / 1e9:.1f} G unknown*RHS/s, "
                  f"{48.0 * n * n * J / world / s / 1e9:.0f} GB/s algorithmic per GPU", flush=True)
    # one transpose alone
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        grid.transpose(0, 1)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        s = float(t.item())
        sent = 8.0 * n * grid.w * (world - 1) / world
        print(f"world={world} distributed transpose: {s * 1e3:.3f} ms, {sent / s / 1e9:.0f} GB/s sent per GPU over NVLink, {16.0 * n * grid.w / s / 1e9:.0f} GB/s local+remote", flush=True)
    grid.close()
    # reference point: copy-engine peer copy of the same number of bytes one rank sends per transpose
    if world > 1 and not quick:
        peer = f"cuda:{(lr + 1) % world}"
        nbytes = int(8 * n * ((n + world - 1) // world) * (world - 1) / world)
        a_ = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
        b_ = torch.empty(nbytes // 8, dtype=torch.float64, device=peer)
        for it in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            b_.copy_(a_, non_blocking=True)
            torch.cuda.synchronize()
            torch.cuda.synchronize(peer)
            dt = time.perf_counter() - t0
        if rank == 0:
            print(f"world={world} torch peer copy of {nbytes / 1e6:.0f} MB (all ranks at once, ring): {dt * 1e3:.3f} ms, {nbytes / dt / 1e9:.0f} GB/s per GPU", flush=True)
    dist.destroy_process_group()


def _cm(view, dev):
    """contiguous column-major copy (even leading dimension) of a 2-D view"""
    rows, cols = view.shape
    out = colmajor(rows, cols, dev)
    out.copy_(view)
    return out


if __name__ == "__main__":
    main()
