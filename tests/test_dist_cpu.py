"""Host-side logic of the multi-GPU paths on CPU: world-size-2 `gloo` process groups."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pop_b200 as P
    # distributed transpose of an n x n matrix held by column blocks (CPU tensors: the local tile
    # transposes run in torch, the exchange is the same all_to_all_single the NCCL path uses)
    g = torch.Generator().manual_seed(3)
    A = torch.randn((n, n), generator=g, dtype=torch.float64)          # same on both ranks
    tr = P.DistTranspose(n, device="cpu")
    lo, hi = tr.lo, tr.hi
    src = A[:, lo:hi].t().contiguous().t()                               # column-major (n x w)
    dst = torch.empty((hi - lo, n), dtype=torch.float64).t()
    tr(src, dst)
    ok_t = bool(torch.equal(dst, A.t()[:, lo:hi]))
    # shards cover the columns exactly once and differ by at most one column
    b = P.shard_bounds(n, world)
    ok_b = b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    # bench-style reduction: max over ranks of a per-rank time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (ok_t, ok_b, float(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 31])
def test_distributed_transpose_world2_gloo(n):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok_t, ok_b, tmax = out[r]
        assert ok_t and ok_b
        assert tmax == 2.0


def test_shard_formula_matches_library_convention():
    """pop_dist (csrc/pop_dist.cu) and drivers.shard_bounds use the same split: first n % world ranks one larger."""
    import pop_b200 as P
    for n, w in ((8191, 8), (8191, 2), (95, 2), (5, 4)):
        base, rem = divmod(n, w)
        lo = [r * base + min(r, rem) for r in range(w + 1)]
        assert P.shard_bounds(n, w) == [(lo[r], lo[r + 1]) for r in range(w)]
