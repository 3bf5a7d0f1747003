"""Per-kernel timing of the element-decomposed ADI at the shapes one GPU sees in a `world`-GPU run, on ONE GPU:
POP_DD_FAKE_PEERS=1 makes every peer alias this rank's own mailboxes (waits pass at once, results are meaningless),
so the numbers are the kernels' own time without NVLink latency or skew.
  POP_DD_FAKE_PEERS=1 python tools/prof_dd.py [world rank J]"""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("POP_DD_FAKE_PEERS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pop_b200 as P  # noqa: E402


def main():
    world, rank, J = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 3, 82)
    dev = "cuda:0"
    torch.cuda.set_device(0)
    P.default_context(0)
    m, N = 128, 64
    Q = P.DirichletPolynomial(np.linspace(-1, 1, m + 1))
    KR = P.Block.oneto(N)
    M, K = P.grammatrix(Q)[KR, KR], -P.weaklaplacian(Q)[KR, KR]
    n = M.n
    dd = P.ElementDecomposition(M, _standin=(rank, world) if world > 1 else None)
    Fl = dd.panel()
    Fl.copy_(torch.randn((n, dd.nloc), dtype=torch.float64, device=dev))
    b = 4 / np.pi ** 2 * (1 + 1e-9)
    a = 3.1e-10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        torch.cuda.synchronize()
        e0.record()
        P.adi_poisson_dd(M, K, Fl, a, b, J=J, dd=dd)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) * 1e-3
    print(f"stand-in for rank {rank} of {world}: n={n} nloc={dd.nloc} J={J}: {s * 1e3:.2f} ms/solve ({s / J * 1e6:.1f} us per iteration)", flush=True)
    dd.close()


if __name__ == "__main__":
    main()
