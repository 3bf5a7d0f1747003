"""Drivers of the reference's examples on the accelerated path.

  adi_poisson : examples/adi_poisson.jl:42-59   ADI(A=M, B=-M, C=K, F, a, b, c, d, tol)
  heat_steps  : examples/heat.jl:58-63          A = factorize(M - h*Delta); u <- A \\ (M*u)
  spectrum_bounds : replaces the dense eigmin/eigmax of examples/adi_poisson.jl:91-94 by power /
                    inverse iteration on the arrowhead kernels.

Multi-GPU (one process per GPU, torch.distributed): X is distributed by column blocks; the
x<->y switch between the two half steps is one all-to-all (NCCL over NVLink) of locally
transposed tiles.  Shift computation and the 2J small factorisations are replicated.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .arrowhead import BBBArrowheadMatrix, Symmetric, _Panel, _bind_stream, _new_like


def _base(S):
    return S.data if isinstance(S, Symmetric) else S


def adi_num_iterations(a, b, c, d, tol=1e-15) -> int:
    return int(L.lib().pop_adi_num_iterations(a, b, c, d, tol))


def adi_shifts(J, a, b, c, d):
    p = np.zeros(J)
    q = np.zeros(J)
    L.check(L.lib().pop_adi_shifts(int(J), a, b, c, d, p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p)), "ADI_shifts")
    return p, q


def adi_poisson(M, K, F, a, b, c=None, d=None, tol=1e-15, J=0, out=None):
    """Solve K X M + M X K = F by the reference's ADI iteration on ONE GPU.  F: CUDA torch tensor
    (n x n, column-major) or NumPy array; returns X of the same kind.  (a,b) enclose the spectrum
    of K^{-1}M; c,d default to -b,-a as in examples/adi_poisson.jl:107."""
    Mb, Kb = _base(M), _base(K)
    c = -b if c is None else c
    d = -a if d is None else d
    if not hasattr(F, "data_ptr"):
        # host arrays: pop_adi_poisson_host (one copy in, the iteration on the device, one copy out)
        Fh = np.asfortranarray(F, dtype=np.float64)
        X = np.empty(Fh.shape, dtype=np.float64, order="F") if out is None else out   # out: a Fortran-ordered array (e.g. a view of pinned memory)
        if X.shape != Fh.shape or not X.flags.f_contiguous or X.dtype != np.float64:
            raise ValueError("adi_poisson: out must be a Fortran-ordered float64 array of the shape of F")
        L.check(L.lib().pop_adi_poisson_host(Mb.ctx.handle, Kb._h, Mb._h, Fh.ctypes.data, Fh.shape[0], X.ctypes.data, X.shape[0], a, b, c, d, tol, int(J)), "ADI")
        return X
    pf = _Panel(F, False)
    X = _new_like(F, tuple(F.shape))
    px = _Panel(X, True)
    _bind_stream(Mb.ctx)
    L.check(L.lib().pop_adi_poisson(Mb.ctx.handle, Kb._h, Mb._h, pf.ptr, pf.ld, px.ptr, px.ld, a, b, c, d, tol, int(J)), "ADI")
    return X


def heat_steps(M, K, U, dt, nsteps, dims=1, theta=1.0):
    """In place: nsteps of the theta scheme (M + theta dt K) u' = (M - (1-theta) dt K) u with one factorisation (K = -Delta):
    theta = 1 backward Euler (examples/heat.jl:58-63), theta = 1/2 the trapezoid rule (examples/heat.jl:26-45)."""
    Mb, Kb = _base(M), _base(K)
    p = _Panel(U, True)
    if not p.torch:
        raise TypeError("heat_steps works on device-resident (CUDA torch) fields")
    _bind_stream(Mb.ctx)
    L.check(L.lib().pop_heat_steps_theta(Mb.ctx.handle, Kb._h, Mb._h, float(dt), float(theta), int(nsteps), int(dims), p.ptr, p.ld, p.cols), "heat")
    return U


def eigvals(A, B=None, k=None, seed=0, return_residuals=False):
    """eigvals(Symmetric(A)) / eigvals(Symmetric(A), Symmetric(B)) (src/arrowhead.jl:536-537) by Lanczos with full
    re-orthogonalisation on the arrowhead kernels (csrc/pop_eig.cu).  k=None: all n eigenvalues; k << n: k Ritz values
    (the extreme ones converge first) with residual bounds |lambda - theta_i| <= resid_i."""
    from .arrowhead import reversecholesky
    Ab = _base(A)
    n = Ab.n
    kk = n if k is None else int(min(k, n))
    UB = None
    if B is not None:
        UB = B.factor if hasattr(B, "factor") else reversecholesky(B if isinstance(B, Symmetric) else Symmetric(B)).factor
    theta = np.zeros(kk)
    resid = np.zeros(kk)
    kout = C.c_int(0)
    L.check(L.lib().pop_eigvals_lanczos(Ab.ctx.handle, Ab._h, UB._h if UB is not None else None, kk, int(seed),
                                        theta.ctypes.data_as(C.c_void_p), resid.ctypes.data_as(C.c_void_p), C.byref(kout)), "eigvals")
    theta, resid = theta[:kout.value], resid[:kout.value]
    return (theta, resid) if return_residuals else theta


def spectrum_bounds(M, K, k=160, seed=0, rtol=1e-6, iters=None):
    """(a, b) with a <= lambda(K^{-1} M) <= b, the ADI interval the reference takes from a dense eigmin/eigmax
    (examples/adi_poisson.jl:91-94).  b = lambda_max of the pencil (M, K) and 1/a = lambda_max of the pencil (K, M), both
    from Lanczos on the arrowhead kernels; each Ritz value is moved outwards by its residual bound (a Ritz value lies
    inside the spectrum, an eigenvalue lies within the residual of it) and by `rtol`, so the interval encloses the
    spectrum.  The number of Lanczos steps doubles until the residual of the extreme Ritz value is below 1e-3 of it."""
    n = _base(M).n

    def lam_max(A, B):
        from .arrowhead import reversecholesky
        FB = reversecholesky(B if isinstance(B, Symmetric) else Symmetric(_base(B)))
        kk = min(k, n)
        while True:
            th, rs = eigvals(A, FB, k=kk, seed=seed, return_residuals=True)
            if rs[-1] <= 1e-3 * abs(th[-1]) or kk >= n:
                return (th[-1] + rs[-1]) * (1 + rtol)
            kk = min(2 * kk, n)

    b = lam_max(M, K)
    lmax = lam_max(K, M)
    return 1.0 / lmax, b


# --------------------------------------------------------------------------------------
# multi-GPU ADI: column-block distribution, all-to-all transposes (torch.distributed)
# --------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int):
    """Column ranges [lo, hi) per rank: near-equal split, first n % world ranks one larger."""
    base, rem = divmod(n, world)
    lo = [r * base + min(r, rem) for r in range(world)]
    return [(lo[r], lo[r] + base + (1 if r < rem else 0)) for r in range(world)]


class DistTranspose:
    """Distributed transpose of an n x n matrix held as column blocks: rank r owns columns
    [lo_r, hi_r).  Afterwards rank r owns the same column range of the TRANSPOSED matrix.
    Tile (rows of block s, my columns) is transposed locally by pop_transpose into the send
    buffer; one all_to_all_single exchanges the tiles; received tiles are already column-major
    slices of the destination and are placed with one strided copy each."""

    def __init__(self, n, group=None, ctx=None, device=None):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n = n
        self.bounds = shard_bounds(n, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        self.w = self.hi - self.lo
        self.ctx = ctx
        self.device = device
        self.send_counts = [(hi - lo) * self.w for lo, hi in self.bounds]
        self.recv_counts = list(self.send_counts)
        self.sendbuf = torch.empty(sum(self.send_counts), dtype=torch.float64, device=device)
        self.recvbuf = torch.empty(sum(self.recv_counts), dtype=torch.float64, device=device)

    def __call__(self, src, dst):
        """src, dst: column-major (n x w) CUDA tensors (views with stride(0) == 1)."""
        import torch
        if src.is_cuda:
            ps = _Panel(src, False)
            if self.ctx is not None:
                _bind_stream(self.ctx)
        off = 0
        for s, (lo, hi) in enumerate(self.bounds):
            rows = hi - lo
            # tile = src[lo:hi, :] (rows x w) -> its transpose (w x rows), column-major, packed
            if src.is_cuda:
                L.check(L.lib().pop_transpose(self.ctx.handle, ps.ptr + 8 * lo, ps.ld, rows, self.w,
                                              self.sendbuf.data_ptr() + 8 * off, self.w), "transpose")
            else:
                self.sendbuf[off: off + rows * self.w].view(rows, self.w).copy_(src[lo:hi, :])
            off += rows * self.w
        if self.world > 1:
            self.dist.all_to_all_single(self.recvbuf, self.sendbuf, self.recv_counts, self.send_counts, group=self.group)
            rb = self.recvbuf
        else:
            rb = self.sendbuf
        # the tile received from rank s is (w_s x w) column-major = dst[lo_s:hi_s, :]
        off = 0
        for s, (lo, hi) in enumerate(self.bounds):
            rows = hi - lo
            tile = rb[off: off + rows * self.w].view(self.w, rows).t()
            dst[lo:hi, :].copy_(tile)
            off += rows * self.w
        return dst


class _DevView:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr, rows, cols, ld):
        self.__cuda_array_interface__ = {"shape": (cols, rows), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": (ld * 8, 8)}


class DistMatrix:
    """n x n matrix distributed by column blocks over the ranks of a torch.distributed group, with the
    x<->y switch done by the library's fused transpose+exchange kernel over NVLink peer memory
    (csrc/pop_dist.cu).  One process per GPU; the CUDA IPC handles of the library-owned buffers are
    exchanged once through the process group (all_gather_object)."""

    def __init__(self, n, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.n = int(n)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        lib = L.lib()
        h = C.c_void_p()
        L.check(lib.pop_dist_create(ctx.handle, self.rank, self.world, self.n, C.byref(h)), "pop_dist_create")
        self._h = h
        if self.world > 1:
            nb = lib.pop_dist_handle_bytes()
            mine = C.create_string_buffer(nb)
            L.check(lib.pop_dist_get_handle(h, mine), "pop_dist_get_handle")
            allh = [None] * self.world
            dist.all_gather_object(allh, bytes(mine.raw), group=group)
            blob = C.create_string_buffer(b"".join(allh), nb * self.world)
            L.check(lib.pop_dist_open(h, blob), "pop_dist_open")
            dist.barrier(group=group)
        self.ld = int(lib.pop_dist_ld(h))
        self.lo = int(lib.pop_dist_lo(h, self.rank))
        self.hi = int(lib.pop_dist_lo(h, self.rank + 1))
        self.w = self.hi - self.lo

    def buffer(self, which):
        """Column-major (n x w) torch view of local buffer `which` (0, 1 work panels, 2 F, 3 F^T)."""
        import torch
        ptr = L.lib().pop_dist_buffer(self._h, int(which))
        t = torch.as_tensor(_DevView(ptr, self.n, self.w, self.ld), device=f"cuda:{self.ctx.device}")
        return t.t()

    def transpose(self, src, dst):
        _bind_stream(self.ctx)
        L.check(L.lib().pop_dist_transpose(self.ctx.handle, self._h, int(src), int(dst)), "pop_dist_transpose")

    def close(self):
        if self._h is not None:
            L.lib().pop_dist_destroy(self.ctx.handle, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def adi_poisson_p2p(M, K, F_local, a, b, c=None, d=None, tol=1e-15, J=0, grid: "DistMatrix" = None, group=None):
    """Multi-GPU ADI with the fused transpose+exchange kernel.  F_local = F[:, lo:hi] (column-major CUDA
    tensor, n x w); returns X[:, lo:hi].  Shift computation and the 2J small factorisations are replicated."""
    Mb, Kb = _base(M), _base(K)
    ctx = Mb.ctx
    c = -b if c is None else c
    d = -a if d is None else d
    own = grid is None
    if own:
        grid = DistMatrix(Mb.n, ctx, group=group)
    pf = _Panel(F_local, False)
    X = _new_like(F_local, tuple(F_local.shape))
    px = _Panel(X, True)
    _bind_stream(ctx)
    L.check(L.lib().pop_adi_poisson_dist(ctx.handle, grid._h, Kb._h, Mb._h, pf.ptr, pf.ld, px.ptr, px.ld, a, b, c, d, tol, int(J)), "ADI (distributed)")
    if own:
        grid.close()
    return X


def dd_columns(m, nh, q, rank, world):
    """Global column indices of rank's local panel in the element decomposition (csrc/pop_row.cu): its hats
    [hA,hB), then for every degree j the bubbles of its elements [k0,k1) -- arrowhead order of the local mesh."""
    base, rem = divmod(m, world)
    kb = [r * base + min(r, rem) for r in range(world + 1)]
    k0, k1 = kb[rank], kb[rank + 1]
    hA, hB = k0, (nh if rank == world - 1 else min(nh, k1))
    cols = list(range(hA, max(hB, hA)))
    for j in range(q):
        cols.extend(nh + j * m + k for k in range(k0, k1))
    return np.asarray(cols, dtype=np.int64)


class ElementDecomposition:
    """X (n x n) distributed by mesh elements of the column direction (csrc/pop_row.cu)."""

    def __init__(self, M, group=None, _standin=None):
        import torch.distributed as dist
        Mb = _base(M)
        self.ctx = Mb.ctx
        dsc = Mb.desc
        self.m, self.nh, self.q = int(dsc.m), int(dsc.nh), int(dsc.q)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if _standin is not None:   # profiling aid: (rank, world) of the shard this single GPU stands in for (POP_DD_FAKE_PEERS=1)
            self.rank, self.world = _standin
        lib = L.lib()
        h = C.c_void_p()
        L.check(lib.pop_dd_create(self.ctx.handle, self.rank, self.world, self.m, self.nh, self.q, C.byref(h)), "pop_dd_create")
        self._h = h
        if self.world > 1 and _standin is None:
            nb = lib.pop_dist_handle_bytes()
            mine = C.create_string_buffer(nb)
            L.check(lib.pop_dd_get_handle(h, mine), "pop_dd_get_handle")
            allh = [None] * self.world
            dist.all_gather_object(allh, bytes(mine.raw), group=group)
            blob = C.create_string_buffer(b"".join(allh), nb * self.world)
            L.check(lib.pop_dd_open(h, blob), "pop_dd_open")
            dist.barrier(group=group)
        self.nloc = int(lib.pop_dd_nloc(h))
        self.ld = int(lib.pop_dd_ld(h))
        self.n = self.nh + self.m * self.q
        self.columns = dd_columns(self.m, self.nh, self.q, self.rank, self.world)
        assert len(self.columns) == self.nloc

    def panel(self, device=None):
        """Uninitialised local panel (n x nloc, column-major with the library's leading dimension)."""
        import torch
        buf = torch.empty((self.nloc, self.ld), dtype=torch.float64, device=device or f"cuda:{self.ctx.device}")
        return buf.t()[:self.n]

    def close(self):
        if self._h is not None:
            L.lib().pop_dd_destroy(self.ctx.handle, self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def adi_poisson_dd(M, K, F_local, a, b, c=None, d=None, tol=1e-15, J=0, dd: "ElementDecomposition" = None, group=None):
    """ADI without transposes on the element-decomposed matrix.  F_local: the columns dd.columns of F as an
    (n x nloc) panel from dd.panel(); returns the same columns of X."""
    Mb, Kb = _base(M), _base(K)
    ctx = Mb.ctx
    c = -b if c is None else c
    d = -a if d is None else d
    own = dd is None
    if own:
        dd = ElementDecomposition(M, group=group)
    pf = _Panel(F_local, False)
    X = dd.panel()
    px = _Panel(X, True)
    _bind_stream(ctx)
    L.check(L.lib().pop_adi_poisson_dd(ctx.handle, dd._h, Kb._h, Mb._h, pf.ptr, pf.ld, px.ptr, px.ld, a, b, c, d, tol, int(J)), "ADI (element decomposition)")
    if own:
        dd.close()
    return X


def heat_steps_p2p(M, K, U_local, dt, nsteps, grid: "DistMatrix" = None, group=None):
    """In place: nsteps of the factored 2D backward Euler on the column-block distributed field U_local = U[:, lo:hi]."""
    Mb, Kb = _base(M), _base(K)
    ctx = Mb.ctx
    own = grid is None
    if own:
        grid = DistMatrix(Mb.n, ctx, group=group)
    p = _Panel(U_local, True)
    _bind_stream(ctx)
    L.check(L.lib().pop_heat_steps_dist(ctx.handle, grid._h, Kb._h, Mb._h, float(dt), int(nsteps), p.ptr, p.ld), "heat (distributed)")
    if own:
        grid.close()
    return U_local


def adi_poisson_distributed(M, K, F_local, Ft_local, a, b, c=None, d=None, tol=1e-15, J=0, group=None):
    """Multi-GPU ADI.  F_local = F[:, lo:hi], Ft_local = F^T[:, lo:hi] (column blocks of F and of its
    transpose, column-major CUDA tensors).  Returns X[:, lo:hi].  Same operator sequence as
    pop_adi_poisson (csrc/pop_host.cu) with the local transposes replaced by DistTranspose."""
    import torch
    from .arrowhead import reversecholesky_shifted
    Mb, Kb = _base(M), _base(K)
    ctx = Mb.ctx
    c = -b if c is None else c
    d = -a if d is None else d
    n = Mb.n
    if J <= 0:
        J = adi_num_iterations(a, b, c, d, tol)
    p, q = adi_shifts(J, a, b, c, d)
    al = np.ones(2 * J)
    sg = np.empty(2 * J)
    sg[0::2] = 1.0 / p
    sg[1::2] = -1.0 / q
    facs = reversecholesky_shifted(Kb, Mb, al, sg)
    tops = [Mb._axpby(s, Kb, -1.0) for s in sg]          # T1_j = M/p_j - K (even), T2_j = -M/q_j - K (odd)
    tr = DistTranspose(n, group=group, ctx=ctx, device=F_local.device)
    w = tr.w
    W = _new_like(F_local, (n, w))
    V = _new_like(F_local, (n, w))
    lib = L.lib()
    pF, pFt = _Panel(F_local, False), _Panel(Ft_local, False)
    pW, pV = _Panel(W, True), _Panel(V, True)
    W.copy_(Ft_local)
    W.mul_(-1.0 / p[0])
    _bind_stream(ctx)
    for j in range(J):
        L.check(lib.pop_solve_mul_axpy(ctx.handle, facs[2 * j].factor._h, tops[2 * j + 1]._h, -1.0 / q[j], pFt.ptr, pFt.ld, pW.ptr, pW.ld, w), "ADI half step")
        tr(W, V)
        if j + 1 < J:
            L.check(lib.pop_solve_mul_axpy(ctx.handle, facs[2 * j + 1].factor._h, tops[2 * j + 2]._h, -1.0 / p[j + 1], pF.ptr, pF.ld, pV.ptr, pV.ld, w), "ADI half step")
            tr(V, W)
        else:
            L.check(lib.pop_solve(ctx.handle, facs[2 * j + 1].factor._h, L.LEFT, pV.ptr, pV.ld, w, L.SOLVE_AUTO), "ADI final solve")
    return V
