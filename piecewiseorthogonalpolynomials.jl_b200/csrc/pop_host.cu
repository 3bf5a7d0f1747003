// Host-buffer variants (chunked, copy/compute overlapped) and the drivers of
// examples/adi_poisson.jl and examples/heat.jl on top of the device kernels.
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <memory>
#include <thread>
#include <vector>

#include "pop_internal.h"

// ------------------------------------------------------------------------------------
// chunked host panels: H2D (copy stream 0) -> kernels (ctx stream) -> D2H (copy stream 1)
// ------------------------------------------------------------------------------------
struct PanelOp {
    // run on a device panel: `vec` x `cnt` region, leading dimension ldd
    virtual int run(pop_ctx *ctx, double *d, int64_t ldd, int64_t cnt) = 0;
    virtual ~PanelOp() {}
};

// dst (rows x cols, leading dimension ldd) <- src (leading dimension lds); device to device re-pitch
__global__ void k_repitch(const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / rows, r = i - c * rows;
        dst[c * ldd + r] = src[c * lds + r];
    }
}

// ------------------------------------------------------------------------------------
// PAGEABLE host panels (a Julia Array, a plain NumPy array).  cudaMemcpyAsync on pageable memory is staged by the driver
// through one small bounce buffer, synchronously: 13 GB/s and no overlap of the two directions on this platform, a sixth
// of what pinned memory gets.  Here worker threads move the panel through pinned pieces instead: each in-worker copies a
// piece of the panel into one of its two private pinned slots and queues the H2D copy of the slot, each out-worker
// queues the D2H copy of a finished piece into a private slot and copies it home.  Pieces are taken in order from
// shared counters; the device streams order everything else.  The slots live with the context (allocated on first use).
// ------------------------------------------------------------------------------------
struct HostStage {
    char *base = nullptr;          // [2 directions][nthr][2 slots][piece]
    size_t piece = 0;
    int nthr = 0;
    std::vector<cudaEvent_t> ev;   // same indexing as the slots
};

static int host_stage_threads() {
    if (const char *e = getenv("POP_HOST_THREADS")) return std::min(std::max(atoi(e), 1), 16);
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(8u, std::max(2u, hw / 3));   // per direction; measured (16 cores): 3 -> 68 ms, 4 -> 58.5, 6 -> 59, 8 -> 55.5 per 1.34 GB panel
}

void pop_host_stage_free(pop_ctx *ctx) {
    HostStage *hs = (HostStage *)ctx->host_stage;
    if (!hs) return;
    for (auto e : hs->ev) cudaEventDestroy(e);
    if (hs->base) cudaFreeHost(hs->base);
    delete hs;
    ctx->host_stage = nullptr;
}

static HostStage *host_stage_get(pop_ctx *ctx, size_t piece) {
    HostStage *hs = (HostStage *)ctx->host_stage;
    const int nthr = host_stage_threads();
    if (hs && hs->piece >= piece && hs->nthr == nthr) return hs;
    pop_host_stage_free(ctx);
    hs = new HostStage();
    hs->piece = piece;
    hs->nthr = nthr;
    if (cudaHostAlloc((void **)&hs->base, (size_t)4 * nthr * piece, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        delete hs;
        return nullptr;
    }
    hs->ev.resize((size_t)4 * nthr);
    for (auto &e : hs->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ctx->host_stage = hs;
    return hs;
}

static bool host_is_pageable(const void *p) {
    if (const char *e = getenv("POP_HOST_STAGE")) {
        if (atoi(e) == 0) return false;   // 0: never stage (the driver's own pageable path), 2: always
        if (atoi(e) == 2) return true;
    }
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// One transfer of a contiguous host range [host, host + pb.back()) to / from the device range starting at dev, in the
// pieces [pb[i], pb[i+1]).  in: start_in() launches the workers; wait_in(bytes) returns once every piece below `bytes` has
// its H2D copy queued on s_copy[0] (the caller then records an event there).  out: start_out(); release_out(bytes) lets
// the pieces below `bytes` go home -- the caller has made s_copy[1] wait for the kernels that produce them before.
struct StagedXfer {
    pop_ctx *ctx;
    HostStage *hs;
    char *host, *host_out, *dev;   // host: what goes to the device; host_out: where the results go (the same for in-place panels)
    std::vector<int64_t> pb;
    std::atomic<int> next_in{0}, next_out{0}, failed{0}, stop{0};
    std::unique_ptr<std::atomic<int>[]> issued;
    std::atomic<int64_t> out_ready{0};
    std::vector<std::thread> thr;

    StagedXfer(pop_ctx *c, HostStage *h, void *hostp, void *devp, std::vector<int64_t> bounds, void *host_outp = nullptr)
        : ctx(c), hs(h), host((char *)hostp), host_out((char *)(host_outp ? host_outp : hostp)), dev((char *)devp), pb(std::move(bounds)),
          issued(new std::atomic<int>[pb.size()]) {
        for (size_t i = 0; i < pb.size(); ++i) issued[i].store(0);
    }
    int npieces() const { return (int)pb.size() - 1; }
    char *slot(int dir, int w, int s) const { return hs->base + (((size_t)dir * hs->nthr + w) * 2 + s) * hs->piece; }
    cudaEvent_t event(int dir, int w, int s) const { return hs->ev[((size_t)dir * hs->nthr + w) * 2 + s]; }

    void in_worker(int w) {
        cudaSetDevice(ctx->device);
        for (int i = 0;; ++i) {
            const int p = next_in.fetch_add(1);
            if (p >= npieces() || stop.load()) break;
            const int s = i & 1;
            const size_t len = (size_t)(pb[p + 1] - pb[p]);
            if (i >= 2 && cudaEventSynchronize(event(0, w, s)) != cudaSuccess) failed.store(1);   // the slot's previous copy has left
            memcpy(slot(0, w, s), host + pb[p], len);
            if (cudaMemcpyAsync(dev + pb[p], slot(0, w, s), len, cudaMemcpyHostToDevice, ctx->s_copy[0]) != cudaSuccess) failed.store(1);
            cudaEventRecord(event(0, w, s), ctx->s_copy[0]);
            issued[p].store(1, std::memory_order_release);
        }
    }
    void out_worker(int w) {
        cudaSetDevice(ctx->device);
        int pend = -1, pend_slot = 0;
        auto finish = [&]() {
            if (pend < 0) return;
            if (cudaEventSynchronize(event(1, w, pend_slot)) != cudaSuccess) failed.store(1);
            memcpy(host_out + pb[pend], slot(1, w, pend_slot), (size_t)(pb[pend + 1] - pb[pend]));
            pend = -1;
        };
        for (int i = 0;; ++i) {
            const int p = next_out.fetch_add(1);
            if (p >= npieces()) break;
            bool go = true;
            if (out_ready.load(std::memory_order_acquire) < pb[p + 1]) {
                finish();                                     // nothing to overlap with: bring the pending piece home first
                while (out_ready.load(std::memory_order_acquire) < pb[p + 1]) {
                    if (stop.load()) { go = false; break; }
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                }
            }
            if (!go) break;
            const int s = i & 1;
            if (pend >= 0 && pend_slot == s) finish();
            if (cudaMemcpyAsync(slot(1, w, s), dev + pb[p], (size_t)(pb[p + 1] - pb[p]), cudaMemcpyDeviceToHost, ctx->s_copy[1]) != cudaSuccess) failed.store(1);
            cudaEventRecord(event(1, w, s), ctx->s_copy[1]);
            if (pend >= 0) finish();                          // the previous piece comes home while this one crosses PCIe
            pend = p;
            pend_slot = s;
        }
        finish();
    }
    void start_in() {
        for (int w = 0; w < hs->nthr; ++w) thr.emplace_back([this, w] { in_worker(w); });
    }
    void start_out() {
        for (int w = 0; w < hs->nthr; ++w) thr.emplace_back([this, w] { out_worker(w); });
    }
    void wait_in(int64_t bytes) {
        for (int p = 0; p < npieces() && pb[p] < bytes; ++p)
            while (!issued[p].load(std::memory_order_acquire)) {
                if (failed.load()) return;
                std::this_thread::yield();
            }
    }
    void release_out(int64_t bytes) { out_ready.store(bytes, std::memory_order_release); }
    // joins the workers; abort = true: give up on what has not been released
    int finish(bool abort) {
        if (abort) stop.store(1);
        for (auto &t : thr) t.join();
        thr.clear();
        return failed.load() ? POP_ERR_CUDA : POP_OK;
    }
};

#define POP_HOST_FLAT_NOFIT (1 << 20)
// Contiguous host panel (n x nrhs, ldx == n): the panel crosses PCIe as FLAT blocks whose host addresses are 4 KiB
// aligned, independent of the column boundaries (measured on this platform, tools/pcie: 48.5 GB/s each way for aligned
// 32 MiB blocks with both directions busy, 43.3 GB/s when the host address is only 8-byte aligned, which is what column
// chunks of an odd n give).  A device mirror of the whole panel receives the blocks; as soon as a block has landed, the
// columns that are complete are re-pitched into the (even-pitch) work buffer, solved, re-pitched back, and every flat
// block whose columns are all done goes home.  Copy streams and the compute stream only meet through events.
static int run_host_flat(pop_ctx *ctx, int64_t n, int64_t nrhs, const double *X, double *Y, PanelOp &op, int64_t blk_bytes) {   // Y: results (== X in place)
    const int64_t colb = n * 8, total = colb * nrhs;
    const int64_t ldd = (n + 1) & ~(int64_t)1;
    const int64_t A = 4096;
    const int64_t off0 = (A - (int64_t)((uintptr_t)X % A)) % A;      // first aligned host address, relative to X
    // block boundaries (bytes): full blocks of blk_bytes in the middle, ramped up from / down to an eighth of a block at the
    // ends so that the first columns are solved and the last block is on its way home after an eighth of a block time
    std::vector<int64_t> bnd;
    bnd.push_back(0);
    const int64_t U = std::max<int64_t>(A, (blk_bytes / 8) & ~(A - 1));            // ramp unit, a multiple of 4 KiB
    {
        const int64_t nu = (total - off0 + U - 1) / U;                             // units to cover
        std::vector<int64_t> sizes;                                                // in units
        if (nu >= 1 + 2 + 4 + 8 + 4 + 2 + 1) {
            const int64_t head[3] = {1, 2, 4}, tail[3] = {4, 2, 1};
            int64_t left = nu - 14;
            for (int i = 0; i < 3; ++i) sizes.push_back(head[i]);
            while (left > 0) { sizes.push_back(std::min<int64_t>(8, left)); left -= 8; }
            for (int i = 0; i < 3; ++i) sizes.push_back(tail[i]);
        } else {
            for (int64_t left = nu; left > 0; left -= 8) sizes.push_back(std::min<int64_t>(8, left));
        }
        int64_t b = off0;
        for (size_t i = 0; i + 1 < sizes.size(); ++i) {
            b += sizes[i] * U;
            if (b < total) bnd.push_back(b);
        }
    }
    bnd.push_back(total);
    const int nblk = (int)bnd.size() - 1;
    int64_t maxcols = 0;
    for (int k = 0; k < nblk; ++k) maxcols = std::max(maxcols, (bnd[k + 1] - bnd[k]) / colb + 2);
    maxcols = std::min(maxcols, nrhs);
    char *dflat = nullptr;
    const size_t mirror = ((size_t)total + 255) & ~(size_t)255;
    cudaError_t e = cudaMallocAsync((void **)&dflat, mirror + (size_t)ldd * maxcols * 8, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return POP_HOST_FLAT_NOFIT;   // no room for the mirror: the caller falls back to column chunks
    }
    double *work = (double *)(dflat + mirror);
    cudaEvent_t *h2d = ctx->ev, *cmp = ctx->ev + 8;
    cudaEventRecord(ctx->ev[24], ctx->stream);                        // work queued earlier on the compute stream (and the allocation)
    cudaStreamWaitEvent(ctx->s_copy[0], ctx->ev[24], 0);
    cudaStreamWaitEvent(ctx->s_copy[1], ctx->ev[24], 0);
    const bool trace = getenv("POP_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<int> tkind;
    auto mark = [&](cudaStream_t st, int kind) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
        tkind.push_back(kind);
    };
    mark(ctx->stream, -1);
    int rc = POP_OK;
    // pageable host memory: worker threads carry the panel through pinned pieces of one ramp unit (see StagedXfer)
    std::unique_ptr<StagedXfer> sx;
    if (host_is_pageable(X) || host_is_pageable(Y)) {
        if (HostStage *hs = host_stage_get(ctx, (size_t)U)) {
            std::vector<int64_t> pb;
            pb.push_back(0);
            for (int64_t b = off0 > 0 ? off0 : U; b < total; b += U) pb.push_back(b);
            pb.push_back(total);
            sx.reset(new StagedXfer(ctx, hs, const_cast<double *>(X), dflat, std::move(pb), Y));
            sx->start_in();
            sx->start_out();
        }
    }
    int64_t col_done = 0;      // columns [0, col_done) are solved
    int blk_home = 0;          // flat blocks [0, blk_home) have been sent back
    for (int k = 0; k < nblk && rc == POP_OK; ++k) {
        mark(ctx->s_copy[0], 0);
        if (sx) {
            sx->wait_in(bnd[k + 1]);
            if (sx->failed.load()) { pop_set_error("host panel H2D (staged): copy failed"); rc = POP_ERR_CUDA; break; }
        } else {
            e = cudaMemcpyAsync(dflat + bnd[k], (const char *)X + bnd[k], (size_t)(bnd[k + 1] - bnd[k]), cudaMemcpyHostToDevice, ctx->s_copy[0]);
            if (e != cudaSuccess) { pop_set_error("host panel H2D: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; break; }
        }
        cudaEventRecord(h2d[k & 7], ctx->s_copy[0]);
        mark(ctx->s_copy[0], 1);
        const int64_t c_hi = k == nblk - 1 ? nrhs : bnd[k + 1] / colb;   // columns complete on the device
        if (c_hi > col_done) {
            cudaStreamWaitEvent(ctx->stream, h2d[k & 7], 0);
            mark(ctx->stream, 2);
            for (int64_t c0 = col_done; c0 < c_hi && rc == POP_OK; c0 += maxcols) {
                const int64_t cnt = std::min(maxcols, c_hi - c0);
                double *src = (double *)(dflat + c0 * colb);
                k_repitch<<<592, 256, 0, ctx->stream>>>(src, n, work, ldd, n, cnt);
                ctx->launches++;
                rc = op.run(ctx, work, ldd, cnt);
                if (rc) break;
                k_repitch<<<592, 256, 0, ctx->stream>>>(work, ldd, src, n, n, cnt);
                ctx->launches++;
            }
            if (rc) break;
            col_done = c_hi;
            cudaEventRecord(cmp[k & 7], ctx->stream);
            mark(ctx->stream, 3);
            // flat blocks whose bytes all belong to solved columns
            int last = blk_home;
            while (last < nblk && bnd[last + 1] <= col_done * colb) ++last;
            if (last > blk_home) {
                cudaStreamWaitEvent(ctx->s_copy[1], cmp[k & 7], 0);
                if (sx) sx->release_out(bnd[last]);
                for (int j = blk_home; j < last && !sx; ++j) {
                    mark(ctx->s_copy[1], 4);
                    e = cudaMemcpyAsync((char *)Y + bnd[j], dflat + bnd[j], (size_t)(bnd[j + 1] - bnd[j]), cudaMemcpyDeviceToHost, ctx->s_copy[1]);
                    if (e != cudaSuccess) { pop_set_error("host panel D2H: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; break; }
                    mark(ctx->s_copy[1], 5);
                }
                blk_home = last;
            }
        }
    }
    if (sx) {
        const int rs = sx->finish(rc != POP_OK);
        if (rc == POP_OK && rs != POP_OK) { pop_set_error("host panel (staged): a copy failed"); rc = rs; }
    }
    cudaStreamSynchronize(ctx->s_copy[0]);
    cudaStreamSynchronize(ctx->stream);
    e = cudaStreamSynchronize(ctx->s_copy[1]);
    if (trace) {
        static const char *names[] = {"H2D start", "H2D end", "compute start", "compute end", "D2H start", "D2H end"};
        for (size_t i = 1; i < tev.size(); ++i) {
            float t;
            cudaEventElapsedTime(&t, tev[0], tev[i]);
            fprintf(stderr, "%8.3f ms  %s\n", t, names[tkind[i]]);
        }
        for (auto ev : tev) cudaEventDestroy(ev);
    }
    cudaFreeAsync(dflat, ctx->stream);
    if (rc == POP_OK && e != cudaSuccess) { pop_set_error("host panel: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    return rc;
}

// X: the host panel that goes to the device; Y (leading dimension ldy): where the results go -- the same panel for the in-place
// calls (ldiv!), another one for `Y = F \\ X`: the two directions of the link then work on different host buffers, which
// this platform rewards (27.7 against 31 ms per 1.34 GB each way)
static int run_host_panels(pop_ctx *ctx, int side, int64_t n, int64_t nrhs, const double *X, int64_t ldx, double *Y, int64_t ldy, PanelOp &op) {
    if (nrhs == 0 || n == 0) return POP_OK;
    if (side == POP_LEFT && ldx == n && ldy == n) {
        // contiguous panel: flat, address-aligned blocks through a device mirror (unless it would not fit)
        int64_t blk = (int64_t)32 << 20;
        if (const char *e = getenv("POP_HOST_CHUNK_MB")) blk = std::max<int64_t>(1, atoll(e)) << 20;
        size_t free_b = 0, total_b = 0;
        int use_flat = 1;
        if (const char *e = getenv("POP_HOST_FLAT")) use_flat = atoi(e);
        if (use_flat && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (size_t)n * nrhs * 8 < free_b / 2) {
            const int rc = run_host_flat(ctx, n, nrhs, X, Y, op, blk);
            if (rc != POP_HOST_FLAT_NOFIT) return rc;
        }
        cudaGetLastError();
    }
    // LEFT: panel = n x nrhs, chunks of columns.  RIGHT: panel = nrhs x n, chunks of rows.
    int64_t target = (int64_t)32 << 20;  // bytes per chunk (measured optimum of the H2D / compute / D2H pipeline)
    if (const char *e = getenv("POP_HOST_CHUNK_MB")) target = std::max<int64_t>(1, atoll(e)) << 20;
    int64_t per = std::max<int64_t>(1, target / (n * 8));
    if (side == POP_RIGHT) per = std::max<int64_t>(per, 32);
    per = std::min(per, nrhs);
    int nbuf = 3;
    if (const char *e = getenv("POP_HOST_NBUF")) nbuf = std::min(std::max(atoi(e), 2), 8);
    const int64_t ldd = side == POP_LEFT ? ((n + 1) & ~(int64_t)1) : ((per + 1) & ~(int64_t)1);
    const size_t buf_doubles = side == POP_LEFT ? (size_t)ldd * per : (size_t)ldd * n;
    // a contiguous host panel whose pitch differs from the (even) device pitch crosses PCIe as ONE linear copy per
    // chunk into a staging area and is re-pitched on the device (pitched 2-D DMA is ~25 % slower on this platform)
    const bool linear = side == POP_LEFT && ldx == n && ldy == n && ldd != n;
    const size_t stg_doubles = linear ? (size_t)n * per : 0;
    double *dbuf = nullptr;
    cudaError_t e = cudaMallocAsync(&dbuf, (buf_doubles + stg_doubles) * nbuf * sizeof(double), ctx->stream);
    if (e != cudaSuccess) {
        pop_set_error("host panel: device buffer of %zu bytes: %s", buf_doubles * nbuf * sizeof(double), cudaGetErrorString(e));
        return POP_ERR_ALLOC;
    }
    cudaEvent_t *h2d = ctx->ev, *cmp = ctx->ev + 8, *d2h = ctx->ev + 16;
    // POP_HOST_TRACE=1: timed events around every copy and compute stage, printed after the call (debugging aid)
    const bool trace = getenv("POP_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
    };
    mark(ctx->stream);
    int rc = POP_OK;
    int64_t done = 0;
    int it = 0;
    // make the copy streams wait for work already queued on the compute stream
    cudaEventRecord(ctx->ev[24], ctx->stream);
    cudaStreamWaitEvent(ctx->s_copy[0], ctx->ev[24], 0);
    while (done < nrhs && rc == POP_OK) {
        const int64_t cnt = std::min(per, nrhs - done);
        const int b = it % nbuf;
        double *d = dbuf + (size_t)b * buf_doubles;
        double *stg = dbuf + (size_t)nbuf * buf_doubles + (size_t)b * stg_doubles;
        if (it >= nbuf) cudaStreamWaitEvent(ctx->s_copy[0], d2h[b], 0);
        mark(ctx->s_copy[0]);
        if (linear)
            e = cudaMemcpyAsync(stg, X + done * ldx, (size_t)n * cnt * 8, cudaMemcpyHostToDevice, ctx->s_copy[0]);
        else if (side == POP_LEFT)
            e = cudaMemcpy2DAsync(d, ldd * 8, X + done * ldx, ldx * 8, n * 8, cnt, cudaMemcpyHostToDevice, ctx->s_copy[0]);
        else
            e = cudaMemcpy2DAsync(d, ldd * 8, X + done, ldx * 8, cnt * 8, n, cudaMemcpyHostToDevice, ctx->s_copy[0]);
        if (e != cudaSuccess) { pop_set_error("host panel H2D: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; break; }
        cudaEventRecord(h2d[b], ctx->s_copy[0]);
        mark(ctx->s_copy[0]);
        cudaStreamWaitEvent(ctx->stream, h2d[b], 0);
        mark(ctx->stream);
        if (linear) {
            k_repitch<<<592, 256, 0, ctx->stream>>>(stg, n, d, ldd, n, cnt);
            ctx->launches++;
        }
        rc = op.run(ctx, d, ldd, cnt);
        if (rc) break;
        if (linear) {
            k_repitch<<<592, 256, 0, ctx->stream>>>(d, ldd, stg, n, n, cnt);
            ctx->launches++;
        }
        cudaEventRecord(cmp[b], ctx->stream);
        mark(ctx->stream);
        cudaStreamWaitEvent(ctx->s_copy[1], cmp[b], 0);
        mark(ctx->s_copy[1]);
        if (linear)
            e = cudaMemcpyAsync(Y + done * ldy, stg, (size_t)n * cnt * 8, cudaMemcpyDeviceToHost, ctx->s_copy[1]);
        else if (side == POP_LEFT)
            e = cudaMemcpy2DAsync(Y + done * ldy, ldy * 8, d, ldd * 8, n * 8, cnt, cudaMemcpyDeviceToHost, ctx->s_copy[1]);
        else
            e = cudaMemcpy2DAsync(Y + done, ldy * 8, d, ldd * 8, cnt * 8, n, cudaMemcpyDeviceToHost, ctx->s_copy[1]);
        if (e != cudaSuccess) { pop_set_error("host panel D2H: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; break; }
        cudaEventRecord(d2h[b], ctx->s_copy[1]);
        mark(ctx->s_copy[1]);
        done += cnt;
        ++it;
    }
    cudaStreamSynchronize(ctx->s_copy[0]);
    cudaStreamSynchronize(ctx->stream);
    e = cudaStreamSynchronize(ctx->s_copy[1]);
    if (trace && tev.size() > 1) {
        // per chunk: [H2D start, H2D end, compute start, compute end, D2H start, D2H end] relative to the call's start
        for (size_t i = 1; i + 5 < tev.size() + 1 && i + 5 <= tev.size() - 0; i += 6) {
            float t[6];
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], tev[0], tev[i + k]);
            fprintf(stderr, "chunk %2zu: H2D %7.3f-%7.3f  compute %7.3f-%7.3f  D2H %7.3f-%7.3f ms\n", (i - 1) / 6, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (auto ev : tev) cudaEventDestroy(ev);
    }
    cudaFreeAsync(dbuf, ctx->stream);
    if (rc == POP_OK && e != cudaSuccess) { pop_set_error("host panel: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    return rc;
}

struct SolveOp : PanelOp {
    const pop_arrowhead *U; int side, algo;
    int run(pop_ctx *ctx, double *d, int64_t ldd, int64_t cnt) override { return pop_solve(ctx, U, side, d, ldd, cnt, algo); }
};
struct TrsmOp : PanelOp {
    const pop_arrowhead *A; int side, uplo, trans, unit;
    int run(pop_ctx *ctx, double *d, int64_t ldd, int64_t cnt) override { return pop_trsm(ctx, A, side, uplo, trans, unit, d, ldd, cnt); }
};

static int check_host_panel(const pop_arrowhead *A, int side, const double *X, int64_t ldx, int64_t nrhs, const char *who) {
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "%s: bad side", who);
    POP_REQUIRE(nrhs >= 0, POP_ERR_DIM, "%s: negative nrhs", who);
    POP_REQUIRE(X || nrhs == 0, POP_ERR_ARG, "%s: null matrix", who);
    const int64_t rows = side == POP_LEFT ? A->n() : nrhs;
    POP_REQUIRE(ldx >= std::max<int64_t>(rows, 1), POP_ERR_DIM, "%s: DimensionMismatch: leading dimension %lld < %lld", who, (long long)ldx, (long long)rows);
    return POP_OK;
}

extern "C" int pop_solve_host(pop_ctx *ctx, const pop_arrowhead *U, int side, double *X, int64_t ldx, int64_t nrhs, int algo) {
    return pop_solve_host_to(ctx, U, side, X, ldx, X, ldx, nrhs, algo);
}

extern "C" int pop_solve_host_to(pop_ctx *ctx, const pop_arrowhead *U, int side, const double *X, int64_t ldx, double *Y, int64_t ldy, int64_t nrhs,
                                 int algo) {
    POP_REQUIRE(ctx && U, POP_ERR_ARG, "pop_solve_host: null argument");
    POP_REQUIRE(U->is_factor, POP_ERR_ARG, "pop_solve_host: handle is not a reverse-Cholesky factor");
    int rc = check_host_panel(U, side, X, ldx, nrhs, "pop_solve_host");
    if (rc) return rc;
    rc = check_host_panel(U, side, Y, ldy, nrhs, "pop_solve_host");
    if (rc) return rc;
    SolveOp op; op.U = U; op.side = side; op.algo = algo;
    return run_host_panels(ctx, side, U->n(), nrhs, X, ldx, Y, ldy, op);
}

extern "C" int pop_trsm_host(pop_ctx *ctx, const pop_arrowhead *A, int side, int uplo, int trans, int unit, double *X, int64_t ldx, int64_t nrhs) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_trsm_host: null argument");
    int rc = check_host_panel(A, side, X, ldx, nrhs, "pop_trsm_host");
    if (rc) return rc;
    TrsmOp op; op.A = A; op.side = side; op.uplo = uplo; op.trans = trans; op.unit = unit;
    return run_host_panels(ctx, side, A->n(), nrhs, X, ldx, X, ldx, op);
}

extern "C" int pop_mul_host(pop_ctx *ctx, const pop_arrowhead *A, int side, int sym, int trans, double alpha, const double *X, int64_t ldx,
                            double beta, double *Y, int64_t ldy, int64_t nrhs) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_mul_host: null argument");
    int rc = check_host_panel(A, side, X, ldx, nrhs, "pop_mul_host");
    if (rc) return rc;
    rc = check_host_panel(A, side, Y, ldy, nrhs, "pop_mul_host");
    if (rc || nrhs == 0) return rc;
    const int64_t n = A->n();
    const int64_t rows = side == POP_LEFT ? n : nrhs, cols = side == POP_LEFT ? nrhs : n;
    const int64_t ldd = (rows + 1) & ~(int64_t)1;
    double *dX = nullptr;
    cudaError_t e = cudaMalloc(&dX, (size_t)ldd * cols * 2 * sizeof(double));
    if (e != cudaSuccess) { pop_set_error("pop_mul_host: device buffer: %s", cudaGetErrorString(e)); return POP_ERR_ALLOC; }
    double *dY = dX + (size_t)ldd * cols;
    e = cudaMemcpy2DAsync(dX, ldd * 8, X, ldx * 8, rows * 8, cols, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && beta != 0.0) e = cudaMemcpy2DAsync(dY, ldd * 8, Y, ldy * 8, rows * 8, cols, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        rc = pop_mul(ctx, A, side, sym, trans, alpha, dX, ldd, beta, dY, ldd, nrhs);
        if (rc == POP_OK) e = cudaMemcpy2DAsync(Y, ldy * 8, dY, ldd * 8, rows * 8, cols, cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(dX);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) { pop_set_error("pop_mul_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2)); return POP_ERR_CUDA; }
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// ADI shifts (examples/adi_poisson.jl:11-40): host arithmetic
// ------------------------------------------------------------------------------------
static double mobius(double z, double a, double b, double c, double d, double al) {  // :11-18
    const double t1 = a * (-al * b + b + al * c + c) - 2 * b * c;
    const double t2 = a * (al * (b + c) - b + c) - 2 * al * b * c;
    const double t3 = 2 * a - (al + 1) * b + (al - 1) * c;
    const double t4 = -al * (-2 * a + b + c) - b + c;
    return (t1 * z + t2) / (t3 * z + t4);
}

// Complete elliptic integral K(m) and Jacobi dn(u|m), m = 1 - m1, by the arithmetic-geometric
// mean / descending Landen recurrence (Abramowitz & Stegun 16.4, 17.6); the reference calls
// Elliptic.K and Elliptic.Jacobi.dn (:30-31).
static void agm_setup(double m1, double *a, double *cc, int *N) {
    double an = 1.0, bn = std::sqrt(m1), cn = std::sqrt(1.0 - m1);
    a[0] = an; cc[0] = cn;
    int n = 0;
    while (std::fabs(cn) > 1e-17 * an && n < 60) {
        const double a1 = 0.5 * (an + bn);
        cn = 0.5 * (an - bn);
        bn = std::sqrt(an * bn);
        an = a1;
        ++n;
        a[n] = an; cc[n] = cn;
    }
    *N = n;
}
static double ellip_K_m1(double m1) {
    double a[64], c[64]; int N;
    agm_setup(m1, a, c, &N);
    return M_PI / (2.0 * a[N]);
}
static double jacobi_dn_m1(double u, double m1) {
    double a[64], c[64]; int N;
    agm_setup(m1, a, c, &N);
    double phi = std::ldexp(a[N] * u, N), phi_prev = phi;
    for (int n = N; n >= 1; --n) {
        phi_prev = phi;
        phi = 0.5 * (phi + std::asin(c[n] * std::sin(phi) / a[n]));
    }
    if (N == 0) return 1.0;  // m = 0
    return std::cos(phi) / std::cos(phi_prev - phi);
}

extern "C" int pop_adi_num_iterations(double a, double b, double c, double d, double tol) {
    const double gamma = (c - a) * (d - b) / ((c - b) * (d - a));   // :47
    return (int)std::ceil(std::log(16 * gamma) * std::log(4 / tol) / (M_PI * M_PI));  // :48
}

extern "C" int pop_adi_shifts(int J, double a, double b, double c, double d, double *p, double *q) {
    POP_REQUIRE(J >= 1 && p && q, POP_ERR_ARG, "pop_adi_shifts: bad argument");
    const double gamma = (c - a) * (d - b) / ((c - b) * (d - a));   // :24
    const double alpha = -1 + 2 * gamma + 2 * std::sqrt(gamma * gamma - gamma);   // :25
    POP_REQUIRE(std::isfinite(alpha) && alpha >= 1, POP_ERR_ARG, "pop_adi_shifts: invalid spectral intervals (alpha=%g)", alpha);
    const double m1 = 1.0 / (alpha * alpha);
    for (int i = 0; i < J; ++i) {
        double dn;
        if (alpha < 1e7) {                                          // :29-31
            const double K = ellip_K_m1(m1);
            dn = jacobi_dn_m1((2.0 * i + 1.0) * K / (2.0 * J), m1);
        } else {                                                    // :32-37
            const double K = 2 * std::log(2.0) + std::log(alpha) + (-1 + 2 * std::log(2.0) + std::log(alpha)) / (alpha * alpha) / 4;
            const double u = (i + 0.5) * K / J;
            const double sech = 1.0 / std::cosh(u);
            dn = sech + m1 / 4 * (std::sinh(u) * std::cosh(u) + u) * std::tanh(u) * sech;
        }
        p[i] = mobius(-alpha * dn, a, b, c, d, alpha);              // :39
        q[i] = mobius(alpha * dn, a, b, c, d, alpha);
    }
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// ADI Poisson driver (examples/adi_poisson.jl:42-59), one GPU
// ------------------------------------------------------------------------------------
__global__ void k_scale_copy(const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, double s) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) dst[c * ldd + r] = s * src[c * lds + r];
}

static int scale_copy(pop_ctx *ctx, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t rows, int64_t cols, double s) {
    dim3 grid((unsigned)((rows + 255) / 256), (unsigned)std::min<int64_t>(cols, 32768));
    k_scale_copy<<<grid, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, cols, s);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

static void free_handles(pop_ctx *ctx, std::vector<pop_arrowhead *> &v) {
    for (auto *h : v) pop_arrowhead_free(ctx, h);
    v.clear();
}

extern "C" int pop_adi_poisson(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, const double *dF, int64_t ldf, double *dX,
                               int64_t ldx, double a, double b, double c, double d, double tol, int J) {
    POP_REQUIRE(ctx && K && M && dF && dX, POP_ERR_ARG, "pop_adi_poisson: null argument");
    POP_REQUIRE(K->d.m == M->d.m && K->d.nh == M->d.nh && K->d.q == M->d.q, POP_ERR_DIM, "pop_adi_poisson: K and M differ in size");
    const int64_t n = K->n();
    POP_REQUIRE(ldf >= n && ldx >= n, POP_ERR_DIM, "pop_adi_poisson: leading dimensions too small");
    // The transpose-free driver (row sweeps with lanes over the rows, pop_row.cu / pop_row2.cu) on a decomposition of one
    // rank: 70 ms instead of 88 ms per solve at n = 8191.  Needs the shared D block and aligned panels; the variant with
    // explicit transposes below serves every other case (POP_ADI_TRANSPOSE=1 forces it).
    {
        const char *e = getenv("POP_ADI_TRANSPOSE");
        const bool aligned = !(ldf & 1) && !((uintptr_t)dF & 15);
        if (!(e && atoi(e)) && aligned && K->mD == 1 && M->mD == 1 && K->d.q <= 128 && K->d.m >= 1 && K->d.m <= 1024 && abs(K->d.nh - K->d.m) <= 1 && K->d.uD <= 2 && M->d.uD <= 2) {
            pop_dd *dd = nullptr;
            int rc = pop_dd_create(ctx, 0, 1, K->d.m, K->d.nh, K->d.q, &dd);
            if (rc == POP_OK) {
                rc = pop_adi_poisson_dd(ctx, dd, K, M, dF, ldf, dX, ldx, a, b, c, d, tol, J);
                pop_dd_destroy(ctx, dd);
                return rc;
            }
            pop_dd_destroy(ctx, dd);   // e.g. out of memory for the work panel: the transposing variant needs less
            cudaGetLastError();
        }
    }
    if (J <= 0) J = pop_adi_num_iterations(a, b, c, d, tol);
    POP_REQUIRE(J >= 1 && J <= 2000, POP_ERR_ARG, "pop_adi_poisson: iteration count %d out of range", J);
    std::vector<double> p(J), q(J);
    int rc = pop_adi_shifts(J, a, b, c, d, p.data(), q.data());
    if (rc) return rc;
    // factors of S1_j = K + M/p_j, S2_j = K - M/q_j  and products T1_j = M/p_j - K, T2_j = -M/q_j - K
    std::vector<double> al(2 * J), sg(2 * J), ta(2 * J), tb(2 * J);
    for (int j = 0; j < J; ++j) {
        al[2 * j] = 1.0; sg[2 * j] = 1.0 / p[j];
        al[2 * j + 1] = 1.0; sg[2 * j + 1] = -1.0 / q[j];
        ta[2 * j] = -1.0; tb[2 * j] = 1.0 / p[j];
        ta[2 * j + 1] = -1.0; tb[2 * j + 1] = -1.0 / q[j];
    }
    std::vector<pop_arrowhead *> fac(2 * J, nullptr), top(2 * J, nullptr);
    rc = pop_revchol_shifted_batch(ctx, K, M, al.data(), sg.data(), 2 * J, fac.data(), nullptr);
    if (rc) { free_handles(ctx, fac); return rc; }
    {
        pop_arrowhead Ku = *K, Mu = *M;   // upper data (Symmetric views)
        Ku.d.nC = 0; Mu.d.nC = 0;
        Ku.p.D = K->p.D + (size_t)K->d.lD * K->d.q * K->mD; Ku.d.lD = 0;
        Mu.p.D = M->p.D + (size_t)M->d.lD * M->d.q * M->mD; Mu.d.lD = 0;
        rc = pop_launch_axpby_batch(ctx, &Ku, &Mu, ta.data(), tb.data(), 2 * J, top.data());
        if (rc) { free_handles(ctx, fac); free_handles(ctx, top); return rc; }
    }
    const int64_t ld = (n + 1) & ~(int64_t)1;
    double *buf = nullptr;
    cudaError_t e = cudaMallocAsync(&buf, (size_t)ld * n * 3 * sizeof(double), ctx->stream);
    if (e != cudaSuccess) {
        free_handles(ctx, fac); free_handles(ctx, top);
        pop_set_error("pop_adi_poisson: %zu bytes of workspace: %s", (size_t)ld * n * 3 * 8, cudaGetErrorString(e));
        return POP_ERR_ALLOC;
    }
    double *W = buf, *V = buf + (size_t)ld * n, *Ft = buf + (size_t)2 * ld * n;
    do {
        if ((rc = pop_transpose(ctx, dF, ldf, n, n, Ft, ld))) break;
        // X0 = 0  =>  R1 = -F/p_1 ; W holds R1^T
        if ((rc = scale_copy(ctx, Ft, ld, W, ld, n, n, -1.0 / p[0]))) break;
        for (int j = 0; j < J && rc == POP_OK; ++j) {
            // W <- T2_j (S1_j^{-1} W) - F^T/q_j      (= R2^T)
            if ((rc = pop_solve_mul_axpy(ctx, fac[2 * j], top[2 * j + 1], -1.0 / q[j], Ft, ld, W, ld, n))) break;
            if ((rc = pop_transpose(ctx, W, ld, n, n, V, ld))) break;
            if (j + 1 < J) {
                // V <- T1_{j+1} (S2_j^{-1} V) - F/p_{j+1}   (= R1 of the next iteration)
                if ((rc = pop_solve_mul_axpy(ctx, fac[2 * j + 1], top[2 * j + 2], -1.0 / p[j + 1], dF, ldf, V, ld, n))) break;
                if ((rc = pop_transpose(ctx, V, ld, n, n, W, ld))) break;
            } else {
                if ((rc = pop_solve(ctx, fac[2 * j + 1], POP_LEFT, V, ld, n, POP_SOLVE_AUTO))) break;
            }
        }
        if (rc) break;
        e = cudaMemcpy2DAsync(dX, ldx * 8, V, ld * 8, n * 8, n, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { pop_set_error("pop_adi_poisson: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    } while (0);
    cudaFreeAsync(buf, ctx->stream);
    free_handles(ctx, fac);
    free_handles(ctx, top);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

// Contiguous pageable host range <-> device range through the pinned pieces (StagedXfer), synchronous for the caller:
// to the device: returns with the copies queued and ctx->stream waiting for them; from the device: returns with the data home.
static int staged_copy(pop_ctx *ctx, bool to_device, void *host, void *dev, int64_t bytes) {
    const int64_t U = (int64_t)4 << 20;
    HostStage *hs = host_stage_get(ctx, (size_t)U);
    POP_REQUIRE(hs, POP_ERR_ALLOC, "staged copy: no pinned staging memory");
    std::vector<int64_t> pb;
    const int64_t off0 = (4096 - (int64_t)((uintptr_t)host % 4096)) % 4096;
    pb.push_back(0);
    for (int64_t b = off0 > 0 ? off0 : U; b < bytes; b += U) pb.push_back(b);
    pb.push_back(bytes);
    StagedXfer sx(ctx, hs, host, dev, std::move(pb));
    if (to_device) {
        POP_CUDA(cudaEventRecord(ctx->ev[24], ctx->stream));
        POP_CUDA(cudaStreamWaitEvent(ctx->s_copy[0], ctx->ev[24], 0));      // the destination is free
        sx.start_in();
        sx.wait_in(bytes);
        const int rc = sx.finish(false);
        POP_REQUIRE(rc == POP_OK, POP_ERR_CUDA, "staged copy to the device failed");
        POP_CUDA(cudaEventRecord(ctx->ev[25], ctx->s_copy[0]));
        POP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev[25], 0));
    } else {
        POP_CUDA(cudaEventRecord(ctx->ev[24], ctx->stream));
        POP_CUDA(cudaStreamWaitEvent(ctx->s_copy[1], ctx->ev[24], 0));      // the source is complete
        sx.start_out();
        sx.release_out(bytes);
        const int rc = sx.finish(false);
        POP_REQUIRE(rc == POP_OK, POP_ERR_CUDA, "staged copy from the device failed");
        POP_CUDA(cudaStreamSynchronize(ctx->s_copy[1]));
    }
    return POP_OK;
}

// Host-buffer variant: F (n x n, column-major, leading dimension ldf) in host memory in, X out.  One H2D copy, the whole
// iteration on the device, one D2H copy: 16 B per entry over PCIe for 2J sweeps of 24 B per entry over HBM.
extern "C" int pop_adi_poisson_host(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, const double *F, int64_t ldf, double *X, int64_t ldx,
                                    double a, double b, double c, double d, double tol, int J) {
    POP_REQUIRE(ctx && K && M && F && X, POP_ERR_ARG, "pop_adi_poisson_host: null argument");
    const int64_t n = K->n();
    POP_REQUIRE(ldf >= n && ldx >= n, POP_ERR_DIM, "pop_adi_poisson_host: leading dimensions too small");
    const int64_t ld = (n + 1) & ~(int64_t)1;
    // Contiguous host matrices (leading dimension n) cross PCIe as ONE linear copy each way and are re-pitched on the device
    // (pitched 2-D DMA with an odd row length runs far below the link rate on this platform); pageable ones (a Julia
    // Array) go through the pinned staging pieces on the way.
    const bool flatF = ldf == n && n > 1, flatX = ldx == n && n > 1;
    const bool stageF = flatF && host_is_pageable(F), stageX = flatX && host_is_pageable(X);
    double *buf = nullptr;
    // from the stream-ordered pool (kept between calls: mapping 1.6 GB of fresh device memory costs tens of milliseconds)
    cudaError_t e = cudaMallocAsync((void **)&buf, (size_t)ld * n * ((flatF || flatX) ? 3 : 2) * sizeof(double), ctx->stream);
    if (e != cudaSuccess) { cudaGetLastError(); pop_set_error("pop_adi_poisson_host: %zu bytes of device memory: %s", (size_t)ld * n * 16, cudaGetErrorString(e)); return POP_ERR_ALLOC; }
    double *dF = buf, *dX = buf + (size_t)ld * n, *flat = buf + (size_t)2 * ld * n;
    int rc = POP_OK;
    if (flatF) {
        if (stageF) rc = staged_copy(ctx, true, const_cast<double *>(F), flat, n * n * 8);
        else e = cudaMemcpyAsync(flat, F, (size_t)n * n * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (rc == POP_OK && e == cudaSuccess) {
            k_repitch<<<592, 256, 0, ctx->stream>>>(flat, n, dF, ld, n, n);
            ctx->launches++;
        }
    } else {
        e = cudaMemcpy2DAsync(dF, ld * 8, F, ldf * 8, n * 8, n, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess && rc == POP_OK) {
        rc = pop_adi_poisson(ctx, K, M, dF, ld, dX, ld, a, b, c, d, tol, J);
        if (rc == POP_OK) {
            if (flatX) {
                k_repitch<<<592, 256, 0, ctx->stream>>>(dX, ld, flat, n, n, n);
                ctx->launches++;
                if (stageX) rc = staged_copy(ctx, false, X, flat, n * n * 8);
                else e = cudaMemcpyAsync(X, flat, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
            } else {
                e = cudaMemcpy2DAsync(X, ldx * 8, dX, ld * 8, n * 8, n, cudaMemcpyDeviceToHost, ctx->stream);
            }
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) { pop_set_error("pop_adi_poisson_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2)); return POP_ERR_CUDA; }
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// heat stepping (examples/heat.jl:58-63): A = factorize(M - h*Delta); u <- A \ (M*u)
// ------------------------------------------------------------------------------------
extern "C" int pop_heat_steps(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, double dt, int nsteps, int dims, double *dU,
                              int64_t ldu, int64_t nrhs) {
    return pop_heat_steps_theta(ctx, K, M, dt, 1.0, nsteps, dims, dU, ldu, nrhs);
}

// theta scheme: (M + theta dt K) u_{k+1} = (M - (1-theta) dt K) u_k ; theta = 1: backward Euler (examples/heat.jl:58-63),
// theta = 1/2: the trapezoid rule A \ (B u), A = M - h/2 Delta, B = M + h/2 Delta (examples/heat.jl:26-45)
extern "C" int pop_heat_steps_theta(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, double dt, double theta, int nsteps, int dims,
                                    double *dU, int64_t ldu, int64_t nrhs) {
    POP_REQUIRE(ctx && K && M && dU, POP_ERR_ARG, "pop_heat_steps: null argument");
    POP_REQUIRE(dims == 1 || dims == 2, POP_ERR_ARG, "pop_heat_steps: dims must be 1 or 2");
    POP_REQUIRE(nsteps >= 0 && dt > 0, POP_ERR_ARG, "pop_heat_steps: need nsteps >= 0, dt > 0");
    POP_REQUIRE(theta > 0 && theta <= 1, POP_ERR_ARG, "pop_heat_steps: theta must lie in (0, 1]");
    const int64_t n = K->n();
    if (dims == 2) nrhs = n;
    POP_REQUIRE(ldu >= n && nrhs >= 0, POP_ERR_DIM, "pop_heat_steps: DimensionMismatch");
    if (nsteps == 0 || nrhs == 0) return POP_OK;
    pop_arrowhead *U = nullptr, *Tm = nullptr;
    double one = 1.0, sdt = theta * dt;
    int rc = pop_revchol_shifted_batch(ctx, M, K, &one, &sdt, 1, &U, nullptr);   // S = M + theta*dt*K
    if (rc) { pop_arrowhead_free(ctx, U); return rc; }
    const pop_arrowhead *Top = M;                                                // right-hand-side operator M - (1-theta)*dt*K
    if (theta < 1.0) {
        rc = pop_axpby(ctx, 1.0, M, -(1.0 - theta) * dt, K, &Tm);
        if (rc) { pop_arrowhead_free(ctx, U); return rc; }
        Top = Tm;
    }
    const int64_t ld = (n + 1) & ~(int64_t)1;
    double *buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)ld * nrhs * sizeof(double));
    if (e != cudaSuccess) { pop_arrowhead_free(ctx, U); pop_arrowhead_free(ctx, Tm); pop_set_error("pop_heat_steps: workspace: %s", cudaGetErrorString(e)); return POP_ERR_ALLOC; }
    MulView mv;
    pop_make_mulview(Top, POP_SYM_UPPER, 0, &mv);
    do {
        if (dims == 1) {
            // v = M u;  (nsteps-1) x  v <- M (S^{-1} v)  in one pass each;  u = S^{-1} v
            if ((rc = pop_launch_mul(ctx, mv, 1.0, dU, 1, ldu, 0.0, buf, 1, ld, nrhs))) break;
            for (int s = 0; s + 1 < nsteps && rc == POP_OK; ++s) rc = pop_solve_mul_axpy(ctx, U, Top, 0.0, buf, ld, buf, ld, nrhs);
            if (rc) break;
            if ((rc = pop_solve(ctx, U, POP_LEFT, buf, ld, nrhs, POP_SOLVE_AUTO))) break;
            e = cudaMemcpy2DAsync(dU, ldu * 8, buf, ld * 8, n * 8, nrhs, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) { pop_set_error("pop_heat_steps: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
        } else {
            // nsteps of U <- S^{-1} M U M S^{-1} = G^nsteps U (G^T)^nsteps with G = S^{-1} M: the column operator and the
            // row operator commute, so all nsteps column sweeps run first, then ONE transpose, all row sweeps, and
            // one transpose back (2 transposes per call instead of 2 per step).  G^k = S^{-1} (M S^{-1})^{k-1} M: one product,
            // k-1 fused solve+product passes (16 B/entry each), one solve.
            double *cur = dU; int64_t ldc = ldu;
            double *oth = buf; int64_t ldo = ld;
            for (int pass = 0; pass < 2 && rc == POP_OK; ++pass) {
                if ((rc = pop_launch_mul(ctx, mv, 1.0, cur, 1, ldc, 0.0, oth, 1, ldo, n))) break;
                for (int s = 0; s + 1 < nsteps && rc == POP_OK; ++s) rc = pop_solve_mul_axpy(ctx, U, Top, 0.0, oth, ldo, oth, ldo, n);
                if (rc) break;
                if ((rc = pop_solve(ctx, U, POP_LEFT, oth, ldo, n, POP_SOLVE_AUTO))) break;
                if ((rc = pop_transpose(ctx, oth, ldo, n, n, cur, ldc))) break;
            }
        }
    } while (0);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    pop_arrowhead_free(ctx, U);
    pop_arrowhead_free(ctx, Tm);
    return rc;
}
