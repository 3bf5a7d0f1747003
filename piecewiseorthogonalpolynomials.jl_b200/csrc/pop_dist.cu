// Multi-GPU support for the 2D drivers (examples/adi_poisson.jl:53-56 alternates sweeps along rows
// and columns of X): one process per GPU, X distributed by column blocks, and the x<->y switch done
// by ONE kernel that transposes 32x32 tiles through shared memory and stores them straight into the
// owning peer's buffer over NVLink (CUDA IPC mappings of buffers this library allocates).  Flags in
// peer memory order the exchange on the device: no host synchronisation inside the iteration.
//
//   rank r owns columns [lo_r, hi_r) of the n x n matrix, stored n x w_r column-major (ld even).
//   transpose: dst_all = src_all^T, i.e. tile src[lo_r:hi_r, :] of rank s lands, transposed, in rows
//   [lo_s, hi_s) of rank r's dst.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pop_internal.h"

#define POP_DIST_MAXWORLD 16
#define POP_DIST_NBUF 4   // W, V, F, Ft

struct pop_dist {
    int rank, world;
    int64_t n, ld, wmax;
    int64_t lo[POP_DIST_MAXWORLD + 1];
    double *buf[POP_DIST_NBUF];                          // local buffers (n x wmax, leading dimension ld)
    unsigned long long *flags;                           // [world] epochs written by the peers
    double *peer_buf[POP_DIST_MAXWORLD][POP_DIST_NBUF];  // IPC mappings (own rank: local pointers)
    unsigned long long *peer_flags[POP_DIST_MAXWORLD];
    unsigned long long epoch;
    bool opened;
    void *slab;                                          // one allocation: buffers + flags
    void *peer_slab[POP_DIST_MAXWORLD];
    size_t buf_bytes;
    cudaStream_t xstream;                                // exchange stream (overlaps the sweeps of later column chunks)
    cudaEvent_t ev_chunk[16], ev_done;
};

struct DistPeers {
    double *dst[POP_DIST_MAXWORLD];
    int64_t lo[POP_DIST_MAXWORLD + 1];
};
struct DistFlags {
    unsigned long long *f[POP_DIST_MAXWORLD];
};
static int pop_scale_inplace(pop_ctx *ctx, double *x, int64_t n, double s);

static void dist_bounds(int64_t n, int world, int64_t *lo) {
    const int64_t base = n / world, rem = n % world;
    for (int r = 0; r <= world; ++r) lo[r] = r * base + std::min<int64_t>(r, rem);
}

extern "C" int pop_dist_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int pop_dist_create(pop_ctx *ctx, int rank, int world, int64_t n, pop_dist **out) {
    POP_REQUIRE(ctx && out, POP_ERR_ARG, "pop_dist_create: null argument");
    POP_REQUIRE(world >= 1 && world <= POP_DIST_MAXWORLD && rank >= 0 && rank < world && n >= 1, POP_ERR_ARG, "pop_dist_create: bad rank/world/n");
    pop_dist *d = new pop_dist();
    memset(d, 0, sizeof(*d));
    d->rank = rank; d->world = world; d->n = n;
    d->ld = (n + 1) & ~(int64_t)1;
    dist_bounds(n, world, d->lo);
    d->wmax = d->lo[1] - d->lo[0];
    d->buf_bytes = ((size_t)d->ld * d->wmax * sizeof(double) + 255) & ~(size_t)255;
    const size_t total = d->buf_bytes * POP_DIST_NBUF + 256;
    cudaError_t e = cudaMalloc(&d->slab, total);
    if (e != cudaSuccess) {
        pop_set_error("pop_dist_create: %zu bytes: %s", total, cudaGetErrorString(e));
        delete d;
        return POP_ERR_ALLOC;
    }
    POP_CUDA(cudaMemsetAsync(d->slab, 0, total, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < POP_DIST_NBUF; ++b) d->buf[b] = (double *)((char *)d->slab + b * d->buf_bytes);
    d->flags = (unsigned long long *)((char *)d->slab + POP_DIST_NBUF * d->buf_bytes);
    for (int b = 0; b < POP_DIST_NBUF; ++b) d->peer_buf[rank][b] = d->buf[b];
    d->peer_flags[rank] = d->flags;
    d->opened = world == 1;
    POP_CUDA(cudaStreamCreateWithFlags(&d->xstream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) POP_CUDA(cudaEventCreateWithFlags(&d->ev_chunk[i], cudaEventDisableTiming));
    POP_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
    *out = d;
    return POP_OK;
}

extern "C" int pop_dist_get_handle(pop_dist *d, void *handle_out) {
    POP_REQUIRE(d && handle_out, POP_ERR_ARG, "pop_dist_get_handle: null argument");
    cudaIpcMemHandle_t h;
    POP_CUDA(cudaIpcGetMemHandle(&h, d->slab));
    memcpy(handle_out, &h, sizeof(h));
    return POP_OK;
}

// all_handles: world handles in rank order (this rank's own entry is ignored)
extern "C" int pop_dist_open(pop_dist *d, const void *all_handles) {
    POP_REQUIRE(d && all_handles, POP_ERR_ARG, "pop_dist_open: null argument");
    for (int r = 0; r < d->world; ++r) {
        if (r == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            pop_set_error("pop_dist_open: cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
            return POP_ERR_CUDA;
        }
        d->peer_slab[r] = p;
        for (int b = 0; b < POP_DIST_NBUF; ++b) d->peer_buf[r][b] = (double *)((char *)p + b * d->buf_bytes);
        d->peer_flags[r] = (unsigned long long *)((char *)p + POP_DIST_NBUF * d->buf_bytes);
    }
    d->opened = true;
    return POP_OK;
}

extern "C" int pop_dist_destroy(pop_ctx *ctx, pop_dist *d) {
    if (!d) return POP_OK;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < d->world; ++r)
        if (r != d->rank && d->peer_slab[r]) cudaIpcCloseMemHandle(d->peer_slab[r]);
    if (d->xstream) {
        cudaStreamSynchronize(d->xstream);
        cudaStreamDestroy(d->xstream);
        for (int i = 0; i < 16; ++i) cudaEventDestroy(d->ev_chunk[i]);
        cudaEventDestroy(d->ev_done);
    }
    if (d->slab) cudaFree(d->slab);
    delete d;
    return POP_OK;
}

extern "C" double *pop_dist_buffer(pop_dist *d, int which) { return (d && which >= 0 && which < POP_DIST_NBUF) ? d->buf[which] : nullptr; }
extern "C" int64_t pop_dist_ld(const pop_dist *d) { return d ? d->ld : 0; }
extern "C" int64_t pop_dist_lo(const pop_dist *d, int r) { return (d && r >= 0 && r <= d->world) ? d->lo[r] : -1; }

// dst_r[(lo_me + c) + (rr - lo_r) * ld] = src[rr + c * ld]   for rr in [lo_r, hi_r), c in [0, w_me)
// blockIdx.z = destination rank; tiles of 32 source rows x 64 source columns, 32x8 threads.  Reads are
// 256 B runs along the source rows; writes are 512 B runs along the destination rows (16 B per lane
// when the destination run is 16 B aligned), which is what the NVLink path to a peer likes.
__global__ void __launch_bounds__(256) k_dist_transpose(const double *__restrict__ src, int64_t ld, int64_t wme, int64_t lome, DistPeers peers,
                                                        int me, int world) {
    __shared__ double tile[64][33];
    // blocks are dispatched z-major: rotate the destinations so that at any moment the ranks write to
    // DIFFERENT peers (no incast on one peer's ingress links while the others idle)
    const int r = (blockIdx.z + me + 1) % world;
    const int64_t lor = peers.lo[r], hir = peers.lo[r + 1];
    const int64_t r0 = lor + (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 64;
    if (r0 >= hir) return;
    for (int cc = threadIdx.y; cc < 64; cc += 8) {
        const int64_t rr = r0 + threadIdx.x, c = c0 + cc;
        if (rr < hir && c < wme) tile[cc][threadIdx.x] = src[c * ld + rr];
    }
    __syncthreads();
    double *dst = peers.dst[r];
    const bool vec = ((lome + c0) & 1) == 0;   // ld is even: every destination run then starts 16 B aligned
    for (int rrl = threadIdx.y; rrl < 32; rrl += 8) {
        const int64_t rr = r0 + rrl;
        if (rr >= hir) continue;
        double *drow = dst + (rr - lor) * ld + lome + c0;
        const int cl = 2 * threadIdx.x;
        if (vec && c0 + cl + 1 < wme) {
            *reinterpret_cast<double2 *>(drow + cl) = make_double2(tile[cl][rrl], tile[cl + 1][rrl]);
        } else {
            if (c0 + cl < wme) drow[cl] = tile[cl][rrl];
            if (c0 + cl + 1 < wme) drow[cl + 1] = tile[cl + 1][rrl];
        }
    }
}

// PULL variant: dst_me[(lo_s + c) + (rr - lo_me) * ld] = src_s[rr + c * ld] for rr in [lo_me, hi_me), c in [0, w_s):
// blockIdx.z = source rank; the 256 B runs along rr are read from the peer's buffer over NVLink, the
// transposed 512 B runs are written locally.
__global__ void __launch_bounds__(256) k_dist_transpose_pull(double *__restrict__ dst, int64_t ld, int64_t lome, int64_t hime, DistPeers peers,
                                                             int me, int world) {
    __shared__ double tile[64][33];
    const int s = (blockIdx.z + me + 1) % world;
    const int64_t los = peers.lo[s], ws = peers.lo[s + 1] - los;
    const int64_t r0 = lome + (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 64;
    if (r0 >= hime || c0 >= ws) return;
    const double *__restrict__ src = peers.dst[s];
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t rr = r0 + threadIdx.x, c = c0 + threadIdx.y + 8 * i;
        v[i] = (rr < hime && c < ws) ? src[c * ld + rr] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tile[threadIdx.y + 8 * i][threadIdx.x] = v[i];
    __syncthreads();
    const bool vec = ((los + c0) & 1) == 0;
    for (int rrl = threadIdx.y; rrl < 32; rrl += 8) {
        const int64_t rr = r0 + rrl;
        if (rr >= hime) continue;
        double *drow = dst + (rr - lome) * ld + los + c0;
        const int cl = 2 * threadIdx.x;
        if (vec && c0 + cl + 1 < ws) {
            *reinterpret_cast<double2 *>(drow + cl) = make_double2(tile[cl][rrl], tile[cl + 1][rrl]);
        } else {
            if (c0 + cl < ws) drow[cl] = tile[cl][rrl];
            if (c0 + cl + 1 < ws) drow[cl + 1] = tile[cl + 1][rrl];
        }
    }
}

// after the transpose kernel has completed: tell every peer that my tiles of epoch `ep` have landed
__global__ void k_dist_signal(DistFlags peers, int world, int me, unsigned long long ep) {
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        volatile unsigned long long *p = peers.f[r] + me;
        *p = ep;
        __threadfence_system();
    }
}

// wait until every peer's tiles of epoch `ep` have landed in my buffer
// (bounded: a peer that died must not hang the device; flags[POP_DIST_MAXWORLD] records a timeout)
__global__ void k_dist_wait(volatile unsigned long long *flags, int world, unsigned long long ep) {
    const int r = threadIdx.x;
    if (r < world) {
        long long spins = 0;
        while (flags[r] < ep) {
            __nanosleep(256);
            if (++spins > 40000000LL) {   // ~10 s
                flags[POP_DIST_MAXWORLD] = ep;
                break;
            }
        }
    }
    __threadfence_system();
}

// After the final stream synchronisation of a distributed driver: did a spin-wait on a peer run into its bound?
// (the kernels then carried on with incomplete data: the result must not be handed back as POP_OK)
static int dist_check_timeout(pop_ctx *ctx, pop_dist *d, const char *who) {
    if (d->world == 1) return POP_OK;
    unsigned long long mark = 0;
    POP_CUDA(cudaMemcpyAsync(&mark, d->flags + POP_DIST_MAXWORLD, 8, cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mark == 0) return POP_OK;
    cudaMemsetAsync(d->flags + POP_DIST_MAXWORLD, 0, 8, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    pop_set_error("%s: rank %d timed out waiting for a peer at exchange epoch %llu; the result is incomplete", who, d->rank, mark);
    return POP_ERR_CUDA;
}

// `handshake`: before the first tile is pushed, every rank announces that its kernels queued so far -- the ones that
// read or write the destination buffer -- have completed, and waits for the same announcement of all peers.  Needed
// whenever the peers' destination buffer was last touched by work that no earlier exchange epoch orders (the first
// exchange of a driver call, the heat driver's passes); inside the ADI loop the epoch of the previous transpose already
// says that the peer has finished reading the buffer, and the handshake is skipped.
static int dist_transpose(pop_ctx *ctx, pop_dist *d, int src_which, int dst_which, bool handshake);
static void dist_handshake(pop_ctx *ctx, pop_dist *d) {
    if (d->world == 1) return;
    const unsigned long long ep0 = ++d->epoch;   // "the kernels I queued so far are done: my buffers are free"
    DistFlags fl0 = {};
    for (int r = 0; r < d->world; ++r) fl0.f[r] = d->peer_flags[r];
    k_dist_signal<<<1, 32, 0, ctx->stream>>>(fl0, d->world, d->rank, ep0);
    k_dist_wait<<<1, 32, 0, ctx->stream>>>(d->flags, d->world, ep0);
    ctx->launches += 2;
}
extern "C" int pop_dist_transpose(pop_ctx *ctx, pop_dist *d, int src_which, int dst_which) {
    return dist_transpose(ctx, d, src_which, dst_which, true);
}

static int dist_transpose(pop_ctx *ctx, pop_dist *d, int src_which, int dst_which, bool handshake) {
    POP_REQUIRE(ctx && d, POP_ERR_ARG, "pop_dist_transpose: null argument");
    POP_REQUIRE(d->opened, POP_ERR_ARG, "pop_dist_transpose: peers not opened (pop_dist_open)");
    POP_REQUIRE(src_which >= 0 && src_which < POP_DIST_NBUF && dst_which >= 0 && dst_which < POP_DIST_NBUF && src_which != dst_which,
                POP_ERR_ARG, "pop_dist_transpose: bad buffer ids");
    DistPeers peers;
    for (int r = 0; r <= d->world; ++r) peers.lo[r] = d->lo[r];
    const int64_t wme = d->lo[d->rank + 1] - d->lo[d->rank];
    static int pull = -1;
    if (pull < 0) {
        const char *e = getenv("POP_DIST_PULL");
        pull = e ? atoi(e) : 0;
    }
    dim3 block(32, 8);
    if (!pull || d->world == 1) {
        // PUSH: transpose my tiles and store them into the owners' buffers; then tell everybody and wait for everybody
        if (handshake) dist_handshake(ctx, d);
        for (int r = 0; r < d->world; ++r) peers.dst[r] = d->peer_buf[r][dst_which];
        dim3 grid((unsigned)((d->wmax + 31) / 32), (unsigned)((wme + 63) / 64), (unsigned)d->world);
        k_dist_transpose<<<grid, block, 0, ctx->stream>>>(d->buf[src_which], d->ld, wme, d->lo[d->rank], peers, d->rank, d->world);
        ctx->launches++;
        if (d->world > 1) {
            const unsigned long long ep = ++d->epoch;
            DistFlags fl = {};
            for (int r = 0; r < d->world; ++r) fl.f[r] = d->peer_flags[r];
            k_dist_signal<<<1, 32, 0, ctx->stream>>>(fl, d->world, d->rank, ep);
            k_dist_wait<<<1, 32, 0, ctx->stream>>>(d->flags, d->world, ep);
            ctx->launches += 2;
        }
    } else {
        // PULL: tell everybody my source panel is complete, wait for everybody's, then read the peers' tiles.
        // A peer's "ready" of the next exchange also says it has finished reading my panel of this one.
        const unsigned long long ep = ++d->epoch;
        DistFlags fl = {};
        for (int r = 0; r < d->world; ++r) fl.f[r] = d->peer_flags[r];
        k_dist_signal<<<1, 32, 0, ctx->stream>>>(fl, d->world, d->rank, ep);
        k_dist_wait<<<1, 32, 0, ctx->stream>>>(d->flags, d->world, ep);
        for (int r = 0; r < d->world; ++r) peers.dst[r] = d->peer_buf[r][src_which];
        dim3 grid((unsigned)((wme + 31) / 32), (unsigned)((d->wmax + 63) / 64), (unsigned)d->world);
        k_dist_transpose_pull<<<grid, block, 0, ctx->stream>>>(d->buf[dst_which], d->ld, d->lo[d->rank], d->lo[d->rank + 1], peers, d->rank, d->world);
        ctx->launches += 3;
    }
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

// One ADI half step on the distributed panel with the exchange OVERLAPPED: the local columns are
// processed in chunks; while the sweep kernel works on chunk c+1 (compute stream), the tiles of chunk c
// are already transposed and pushed to their owners (exchange stream).  Chunks of <= 256 columns leave
// every SM room for transpose blocks next to the persistent sweep CTAs.
static int dist_half_step(pop_ctx *ctx, pop_dist *d, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *Fp,
                          int src_which, int dst_which, int nchunks) {
    const int64_t w = d->lo[d->rank + 1] - d->lo[d->rank];
    const int64_t ld = d->ld;
    double *src = d->buf[src_which];
    if (nchunks <= 1 || d->world == 1) {
        int rc = pop_solve_mul_axpy(ctx, U, T, gamma, Fp, ld, src, ld, w);
        if (rc) return rc;
        return pop_dist_transpose(ctx, d, src_which, dst_which);
    }
    DistPeers peers;
    for (int r = 0; r <= d->world; ++r) peers.lo[r] = d->lo[r];
    for (int r = 0; r < d->world; ++r) peers.dst[r] = d->peer_buf[r][dst_which];
    const int64_t cw = ((w + nchunks - 1) / nchunks + 63) & ~(int64_t)63;   // multiples of the tile width keep the 16 B stores
    int c = 0;
    for (int64_t c0 = 0; c0 < w; c0 += cw, ++c) {
        const int64_t cols = std::min(cw, w - c0);
        int rc = pop_solve_mul_axpy(ctx, U, T, gamma, Fp + c0 * ld, ld, src + c0 * ld, ld, cols);
        if (rc) return rc;
        POP_CUDA(cudaEventRecord(d->ev_chunk[c & 15], ctx->stream));
        POP_CUDA(cudaStreamWaitEvent(d->xstream, d->ev_chunk[c & 15], 0));
        dim3 block(32, 8), grid((unsigned)((d->wmax + 31) / 32), (unsigned)((cols + 63) / 64), (unsigned)d->world);
        k_dist_transpose<<<grid, block, 0, d->xstream>>>(src + c0 * ld, ld, cols, d->lo[d->rank] + c0, peers, d->rank, d->world);
        ctx->launches++;
    }
    const unsigned long long ep = ++d->epoch;
    DistFlags fl = {};
    for (int r = 0; r < d->world; ++r) fl.f[r] = d->peer_flags[r];
    k_dist_signal<<<1, 32, 0, d->xstream>>>(fl, d->world, d->rank, ep);
    k_dist_wait<<<1, 32, 0, d->xstream>>>(d->flags, d->world, ep);
    ctx->launches += 2;
    POP_CUDA(cudaEventRecord(d->ev_done, d->xstream));
    POP_CUDA(cudaStreamWaitEvent(ctx->stream, d->ev_done, 0));
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// distributed ADI (examples/adi_poisson.jl:42-59), same operator sequence as pop_adi_poisson
// ------------------------------------------------------------------------------------
extern "C" int pop_adi_poisson_dist(pop_ctx *ctx, pop_dist *d, const pop_arrowhead *K, const pop_arrowhead *M, const double *dF_local,
                                    int64_t ldf, double *dX_local, int64_t ldx, double a, double b, double c, double dd, double tol, int J) {
    POP_REQUIRE(ctx && d && K && M && dF_local && dX_local, POP_ERR_ARG, "pop_adi_poisson_dist: null argument");
    const int64_t n = K->n();
    POP_REQUIRE(n == d->n, POP_ERR_DIM, "pop_adi_poisson_dist: operator size %lld != distributed size %lld", (long long)n, (long long)d->n);
    POP_REQUIRE(ldf >= n && ldx >= n, POP_ERR_DIM, "pop_adi_poisson_dist: leading dimensions too small");
    if (J <= 0) J = pop_adi_num_iterations(a, b, c, dd, tol);
    POP_REQUIRE(J >= 1 && J <= 2000, POP_ERR_ARG, "pop_adi_poisson_dist: iteration count %d out of range", J);
    std::vector<double> p(J), q(J);
    int rc = pop_adi_shifts(J, a, b, c, dd, p.data(), q.data());
    if (rc) return rc;
    std::vector<double> al(2 * J), sg(2 * J), ta(2 * J), tb(2 * J);
    for (int j = 0; j < J; ++j) {
        al[2 * j] = 1.0; sg[2 * j] = 1.0 / p[j];
        al[2 * j + 1] = 1.0; sg[2 * j + 1] = -1.0 / q[j];
        ta[2 * j] = -1.0; tb[2 * j] = 1.0 / p[j];
        ta[2 * j + 1] = -1.0; tb[2 * j + 1] = -1.0 / q[j];
    }
    std::vector<pop_arrowhead *> fac(2 * J, nullptr), top(2 * J, nullptr);
    auto cleanup = [&]() {
        for (auto *h : fac) pop_arrowhead_free(ctx, h);
        for (auto *h : top) pop_arrowhead_free(ctx, h);
    };
    rc = pop_revchol_shifted_batch(ctx, K, M, al.data(), sg.data(), 2 * J, fac.data(), nullptr);
    if (rc) { cleanup(); return rc; }
    {
        pop_arrowhead Ku = *K, Mu = *M;   // upper data (Symmetric views)
        Ku.d.nC = 0; Mu.d.nC = 0;
        Ku.p.D = K->p.D + (size_t)K->d.lD * K->d.q * K->mD; Ku.d.lD = 0;
        Mu.p.D = M->p.D + (size_t)M->d.lD * M->d.q * M->mD; Mu.d.lD = 0;
        rc = pop_launch_axpby_batch(ctx, &Ku, &Mu, ta.data(), tb.data(), 2 * J, top.data());
        if (rc) { cleanup(); return rc; }
    }
    const int64_t ld = d->ld;
    const int64_t w = d->lo[d->rank + 1] - d->lo[d->rank];
    double *W = d->buf[0], *V = d->buf[1], *Fl = d->buf[2], *Ft = d->buf[3];
    cudaError_t e = cudaSuccess;
    // column chunks per half step (exchange of chunk c overlaps the sweep of chunk c+1).  Measured at 8 GPUs
    // (n = 8191, 1024 columns per GPU): 1 chunk 28.1 ms/solve, 4 chunks 31.0 ms, 8 chunks 39.1 ms -- the persistent
    // sweep kernel needs several columns per CTA to pipeline its loads, so chunking costs more than the overlap gains.
    // Default: no chunking; POP_DIST_CHUNKS opts in.
    int nchunks = 1;
    if (const char *e_ = getenv("POP_DIST_CHUNKS")) nchunks = std::max(1, std::min(16, atoi(e_)));
    // POP_DIST_TIMING=1: per-phase device times (events on the stream), printed by every rank
    const bool timing = getenv("POP_DIST_TIMING") && atoi(getenv("POP_DIST_TIMING"));
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() {
        if (!timing) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, ctx->stream);
        evs.push_back(ev);
    };
    do {
        // F (my columns) and F^T (my columns of the transpose)
        e = cudaMemcpy2DAsync(Fl, ld * 8, dF_local, ldf * 8, n * 8, w, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) break;
        if ((rc = pop_dist_transpose(ctx, d, 2, 3))) break;
        // X0 = 0  =>  R1 = -F/p_1 ; W holds my columns of R1^T
        {
            const double s = -1.0 / p[0];
            // W <- s * Ft
            e = cudaMemcpyAsync(W, Ft, (size_t)ld * w * 8, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) break;
            if ((rc = pop_scale_inplace(ctx, W, ld * w, s))) break;
        }
        for (int j = 0; j < J && rc == POP_OK; ++j) {
            // W <- T2_j (S1_j^{-1} W) - F^T/q_j   (my columns of R2^T), then V <- transpose (my columns of R2)
            mark();
            if (nchunks > 1) {
                if (j == 0) dist_handshake(ctx, d);
                if ((rc = dist_half_step(ctx, d, fac[2 * j], top[2 * j + 1], -1.0 / q[j], Ft, 0, 1, nchunks))) break;
                mark();
            } else {
                if ((rc = pop_solve_mul_axpy(ctx, fac[2 * j], top[2 * j + 1], -1.0 / q[j], Ft, ld, W, ld, w))) break;
                mark();
                if ((rc = dist_transpose(ctx, d, 0, 1, j == 0))) break;
            }
            mark();
            if (j + 1 < J) {
                if (nchunks > 1) {
                    if ((rc = dist_half_step(ctx, d, fac[2 * j + 1], top[2 * j + 2], -1.0 / p[j + 1], Fl, 1, 0, nchunks))) break;
                    mark();
                    continue;
                }
                if ((rc = pop_solve_mul_axpy(ctx, fac[2 * j + 1], top[2 * j + 2], -1.0 / p[j + 1], Fl, ld, V, ld, w))) break;
                mark();
                if ((rc = dist_transpose(ctx, d, 1, 0, j == 0))) break;
            } else {
                if ((rc = pop_solve(ctx, fac[2 * j + 1], POP_LEFT, V, ld, w, POP_SOLVE_AUTO))) break;
            }
        }
        if (rc) break;
        e = cudaMemcpy2DAsync(dX_local, ldx * 8, V, ld * 8, n * 8, w, cudaMemcpyDeviceToDevice, ctx->stream);
    } while (0);
    if (rc == POP_OK && e != cudaSuccess) { pop_set_error("pop_adi_poisson_dist: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    cudaStreamSynchronize(ctx->stream);
    if (timing && evs.size() >= 5) {
        double tf = 0, tt = 0;
        for (size_t i = 0; i + 1 < evs.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
            if ((i % 4) == 0 || (i % 4) == 2) tf += ms; else tt += ms;
        }
        float tot = 0;
        cudaEventElapsedTime(&tot, evs.front(), evs.back());
        fprintf(stderr, "[pop_dist rank %d] J=%d: fused half steps %.2f ms, transposes+exchange %.2f ms, loop total %.2f ms\n", d->rank, J, tf, tt, tot);
        for (auto ev : evs) cudaEventDestroy(ev);
    }
    cleanup();
    if (rc == POP_OK) rc = dist_check_timeout(ctx, d, "pop_adi_poisson_dist");
    return rc;
}

// ------------------------------------------------------------------------------------
// distributed 2D heat stepping (2D extension of examples/heat.jl:58-63): nsteps of
// U <- S^{-1} M U M S^{-1}, S = M + dt K, on the column-block distributed field.  Column and row
// operators commute: all column sweeps are local, then one distributed transpose, the same sweeps
// again, and one transpose back -- 2 exchanges per call.
// ------------------------------------------------------------------------------------
extern "C" int pop_heat_steps_dist(pop_ctx *ctx, pop_dist *d, const pop_arrowhead *K, const pop_arrowhead *M, double dt, int nsteps,
                                   double *dU_local, int64_t ldu) {
    POP_REQUIRE(ctx && d && K && M && dU_local, POP_ERR_ARG, "pop_heat_steps_dist: null argument");
    const int64_t n = K->n();
    POP_REQUIRE(n == d->n && ldu >= n, POP_ERR_DIM, "pop_heat_steps_dist: DimensionMismatch");
    POP_REQUIRE(nsteps >= 0 && dt > 0, POP_ERR_ARG, "pop_heat_steps_dist: need nsteps >= 0, dt > 0");
    if (nsteps == 0) return POP_OK;
    pop_arrowhead *U = nullptr;
    double one = 1.0;
    int rc = pop_revchol_shifted_batch(ctx, M, K, &one, &dt, 1, &U, nullptr);   // S = M + dt*K
    if (rc) { pop_arrowhead_free(ctx, U); return rc; }
    const int64_t ld = d->ld;
    const int64_t w = d->lo[d->rank + 1] - d->lo[d->rank];
    MulView mv;
    pop_make_mulview(M, POP_SYM_UPPER, 0, &mv);
    cudaError_t e = cudaMemcpy2DAsync(d->buf[0], ld * 8, dU_local, ldu * 8, n * 8, w, cudaMemcpyDeviceToDevice, ctx->stream);
    int cur = 0;
    for (int pass = 0; pass < 2 && rc == POP_OK && e == cudaSuccess; ++pass) {
        double *src = d->buf[cur], *dst = d->buf[cur ^ 1];
        if ((rc = pop_launch_mul(ctx, mv, 1.0, src, 1, ld, 0.0, dst, 1, ld, w))) break;
        for (int s = 0; s + 1 < nsteps && rc == POP_OK; ++s) rc = pop_solve_mul_axpy(ctx, U, M, 0.0, dst, ld, dst, ld, w);
        if (rc) break;
        if ((rc = pop_solve(ctx, U, POP_LEFT, dst, ld, w, POP_SOLVE_AUTO))) break;
        if ((rc = pop_dist_transpose(ctx, d, cur ^ 1, cur))) break;     // back into buf[cur], transposed
    }
    if (rc == POP_OK && e == cudaSuccess)
        e = cudaMemcpy2DAsync(dU_local, ldu * 8, d->buf[cur], ld * 8, n * 8, w, cudaMemcpyDeviceToDevice, ctx->stream);
    if (rc == POP_OK && e != cudaSuccess) { pop_set_error("pop_heat_steps_dist: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    cudaStreamSynchronize(ctx->stream);
    pop_arrowhead_free(ctx, U);
    if (rc == POP_OK) rc = dist_check_timeout(ctx, d, "pop_heat_steps_dist");
    return rc;
}

__global__ void k_scale(double *x, int64_t n, double s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= s;
}
static int pop_scale_inplace(pop_ctx *ctx, double *x, int64_t n, double s) {
    k_scale<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(x, n, s);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}
