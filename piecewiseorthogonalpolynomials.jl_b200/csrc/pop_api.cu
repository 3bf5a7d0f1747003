// Context, packed-arrowhead container, operator views, assembly and algebra.
// Reference interfaces replaced: see include/pop_arrowhead.h.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#include <limits.h>

#include "pop_internal.h"

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void pop_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *pop_last_error(void) { return g_err; }
extern "C" int pop_abi_version(void) { return POP_ABI_VERSION; }

// ------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------
static int ctx_init(pop_ctx *ctx, int device, void *stream);
extern "C" int pop_create(pop_ctx **out, int device, void *stream) {
    POP_REQUIRE(out != nullptr, POP_ERR_ARG, "pop_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        pop_set_error("pop_create: no CUDA device (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return POP_ERR_NOGPU;
    }
    POP_REQUIRE(device >= 0 && device < ndev, POP_ERR_ARG, "pop_create: device %d out of range [0,%d)", device, ndev);
    POP_CUDA(cudaSetDevice(device));
    pop_ctx *ctx = new pop_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    const int rc = ctx_init(ctx, device, stream);
    if (rc != POP_OK) {   // nothing half-built survives a failed set-up
        pop_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return POP_OK;
}

static int ctx_init(pop_ctx *ctx, int device, void *stream) {
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        POP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    POP_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    POP_CUDA(cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    {   // keep freed stream-ordered allocations in the pool (handles and driver workspaces are recycled)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    POP_CUDA(cudaMalloc(&ctx->d_info, 4096 * sizeof(int)));
    POP_CUDA(cudaMallocHost(&ctx->h_info, 4096 * sizeof(int)));
    for (int i = 0; i < 2; ++i) POP_CUDA(cudaStreamCreateWithFlags(&ctx->s_copy[i], cudaStreamNonBlocking));
    for (int i = 0; i < 32; ++i) POP_CUDA(cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    {
        // stream-ordered allocations (work panels, the host-panel mirror) stay in the pool between calls instead of going
        // back to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return POP_OK;
}

extern "C" int pop_destroy(pop_ctx *ctx) {
    if (!ctx) return POP_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream || !ctx->own_stream) cudaStreamSynchronize(ctx->stream);
    pop_host_stage_free(ctx);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_info) cudaFree(ctx->d_info);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    for (int i = 0; i < 2; ++i)
        if (ctx->s_copy[i]) cudaStreamDestroy(ctx->s_copy[i]);
    for (int i = 0; i < 32; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return POP_OK;
}

// ---- device-resident panels for hosts without a CUDA binding of their own (the Julia shim's DevicePanel) -----------
extern "C" int pop_panel_alloc(pop_ctx *ctx, int64_t rows, int64_t cols, double **dptr, int64_t *ld) {
    POP_REQUIRE(ctx && dptr && ld, POP_ERR_ARG, "pop_panel_alloc: null argument");
    POP_REQUIRE(rows >= 0 && cols >= 0, POP_ERR_DIM, "pop_panel_alloc: negative size");
    *ld = std::max<int64_t>((rows + 1) & ~(int64_t)1, 2);   // even pitch: every column is 16-byte aligned (the fused kernels' TMA path)
    *dptr = nullptr;
    cudaError_t e = cudaMalloc((void **)dptr, (size_t)*ld * (size_t)std::max<int64_t>(cols, 1) * sizeof(double));
    if (e != cudaSuccess) {
        pop_set_error("pop_panel_alloc: %lld x %lld doubles: %s", (long long)rows, (long long)cols, cudaGetErrorString(e));
        return POP_ERR_ALLOC;
    }
    POP_CUDA(cudaMemsetAsync(*dptr, 0, (size_t)*ld * (size_t)std::max<int64_t>(cols, 1) * sizeof(double), ctx->stream));
    return POP_OK;
}

extern "C" int pop_panel_free(pop_ctx *ctx, double *dptr) {
    POP_REQUIRE(ctx, POP_ERR_ARG, "pop_panel_free: null ctx");
    if (!dptr) return POP_OK;
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    POP_CUDA(cudaFree(dptr));
    return POP_OK;
}

extern "C" int pop_panel_upload(pop_ctx *ctx, const double *host, int64_t ldh, double *dptr, int64_t ldd, int64_t rows, int64_t cols) {
    POP_REQUIRE(ctx && (host || rows * cols == 0) && (dptr || rows * cols == 0), POP_ERR_ARG, "pop_panel_upload: null argument");
    POP_REQUIRE(rows >= 0 && cols >= 0 && ldh >= std::max<int64_t>(rows, 1) && ldd >= std::max<int64_t>(rows, 1), POP_ERR_DIM, "pop_panel_upload: DimensionMismatch");
    if (rows == 0 || cols == 0) return POP_OK;
    POP_CUDA(cudaMemcpy2DAsync(dptr, ldd * 8, host, ldh * 8, rows * 8, cols, cudaMemcpyHostToDevice, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));   // the host array may be freed by its owner as soon as the call returns
    return POP_OK;
}

extern "C" int pop_panel_download(pop_ctx *ctx, const double *dptr, int64_t ldd, double *host, int64_t ldh, int64_t rows, int64_t cols) {
    POP_REQUIRE(ctx && (host || rows * cols == 0) && (dptr || rows * cols == 0), POP_ERR_ARG, "pop_panel_download: null argument");
    POP_REQUIRE(rows >= 0 && cols >= 0 && ldh >= std::max<int64_t>(rows, 1) && ldd >= std::max<int64_t>(rows, 1), POP_ERR_DIM, "pop_panel_download: DimensionMismatch");
    if (rows == 0 || cols == 0) return POP_OK;
    POP_CUDA(cudaMemcpy2DAsync(host, ldh * 8, dptr, ldd * 8, rows * 8, cols, cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    return POP_OK;
}

extern "C" int pop_set_stream(pop_ctx *ctx, void *stream) {
    POP_REQUIRE(ctx, POP_ERR_ARG, "pop_set_stream: null ctx");
    if ((cudaStream_t)stream == ctx->stream && !ctx->own_stream) return POP_OK;
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    // the handle is used as given; NULL is the CUDA (legacy) default stream, which is also what
    // torch.cuda.current_stream().cuda_stream reports for torch's default stream
    ctx->stream = (cudaStream_t)stream;
    return POP_OK;
}

extern "C" int pop_sync(pop_ctx *ctx) {
    POP_REQUIRE(ctx, POP_ERR_ARG, "pop_sync: null ctx");
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    return POP_OK;
}

extern "C" int64_t pop_launch_count(const pop_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pop_ctx_scratch(pop_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->d_scratch_bytes) {
        POP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scratch) POP_CUDA(cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
        ctx->d_scratch_bytes = 0;
        size_t want = std::max(bytes, (size_t)1 << 20);
        cudaError_t e = cudaMalloc(&ctx->d_scratch, want);
        if (e != cudaSuccess) {
            pop_set_error("scratch allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            return POP_ERR_ALLOC;
        }
        ctx->d_scratch_bytes = want;
    }
    *out = ctx->d_scratch;
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------
static inline size_t even(size_t x) { return (x + 1) & ~(size_t)1; }

ArrowSizes pop_sizes(const pop_desc &d, int mD) {
    ArrowSizes s;
    s.Ad = even(d.nh);
    s.Au = even(d.nh > 0 ? d.nh - 1 : 0);
    s.Al = s.Au;
    s.B = even((size_t)d.nB * 3 * d.m);
    s.C = even((size_t)d.nC * 3 * d.m);
    s.D = even((size_t)(d.lD + d.uD + 1) * d.q * mD);
    s.rdA = even(d.nh);
    s.rdD = even((size_t)d.q * mD);
    s.total = s.Ad + s.Au + s.Al + s.B + s.C + s.D + s.rdA + s.rdD;
    return s;
}

static void carve(pop_arrowhead *h, double *base) {
    ArrowSizes s = pop_sizes(h->d, h->mD);
    double *p = base;
    h->p.Ad = p; p += s.Ad;
    h->p.Au = p; p += s.Au;
    h->p.Al = p; p += s.Al;
    h->p.B = p; p += s.B;
    h->p.C = p; p += s.C;
    h->p.D = p; p += s.D;
    h->p.rdA = p; p += s.rdA;
    h->p.rdD = p; p += s.rdD;
}

int pop_alloc_handles(pop_ctx *ctx, const pop_desc &d, int mD, int count, pop_arrowhead **out) {
    ArrowSizes s = pop_sizes(d, mD);
    size_t bytes = std::max<size_t>(s.total, 2) * sizeof(double) * (size_t)count;
    pop_slab *slab = new pop_slab();
    // stream-ordered pool allocation: the ADI drivers create and drop 4J small handles per solve
    cudaError_t e = cudaMallocAsync(&slab->dptr, bytes, ctx->stream);
    if (e != cudaSuccess) {
        delete slab;
        pop_set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return POP_ERR_ALLOC;
    }
    e = cudaMemsetAsync(slab->dptr, 0, bytes, ctx->stream);
    if (e != cudaSuccess) {
        cudaFreeAsync(slab->dptr, ctx->stream);
        delete slab;
        pop_set_error("memset failed: %s", cudaGetErrorString(e));
        return POP_ERR_CUDA;
    }
    slab->refs = count;
    for (int i = 0; i < count; ++i) {
        pop_arrowhead *h = new pop_arrowhead();
        h->d = d;
        h->d.uniform = (mD == 1) ? 1 : 0;
        h->mD = mD;
        h->slab = slab;
        h->has_recip = false;
        h->is_factor = false;
        h->bmask = 0;
        h->band1_zero = false;
        carve(h, (double *)slab->dptr + (size_t)i * std::max<size_t>(s.total, 2));
        out[i] = h;
    }
    return POP_OK;
}

static int validate_desc(const pop_desc *d) {
    POP_REQUIRE(d, POP_ERR_ARG, "null descriptor");
    POP_REQUIRE(d->m >= 1 && d->nh >= 1 && d->q >= 0, POP_ERR_DIM, "descriptor: need m>=1, nh>=1, q>=0 (m=%d nh=%d q=%d)", d->m, d->nh, d->q);
    POP_REQUIRE(d->nB >= 0 && d->nC >= 0, POP_ERR_DIM, "descriptor: negative block count");
    POP_REQUIRE(d->lD >= 0 && d->uD >= 0 && d->lD <= POP_MAXD && d->uD <= POP_MAXD, POP_ERR_DIM,
                "descriptor: D bandwidths (%d,%d) outside [0,%d]", d->lD, d->uD, POP_MAXD);
    // hats touched by element k are k-1,k,k+1: the hat count must be within one of m
    POP_REQUIRE(std::abs(d->nh - d->m) <= 1, POP_ERR_STRUCT,
                "descriptor: nh=%d incompatible with m=%d (blocks of B/C must have bandwidths within (1,1), src/arrowhead.jl:100-111)", d->nh, d->m);
    return POP_OK;
}

extern "C" int pop_arrowhead_upload(pop_ctx *ctx, const pop_desc *desc, const double *Ad, const double *Au,
                                    const double *Al, const double *B, const double *C, const double *D,
                                    pop_arrowhead **out) {
    POP_REQUIRE(ctx && out, POP_ERR_ARG, "pop_arrowhead_upload: null ctx/out");
    *out = nullptr;
    int rc = validate_desc(desc);
    if (rc) return rc;
    POP_REQUIRE(Ad != nullptr, POP_ERR_ARG, "pop_arrowhead_upload: Ad is required");
    POP_REQUIRE(D != nullptr || desc->q == 0, POP_ERR_ARG, "pop_arrowhead_upload: D is required");
    POP_REQUIRE((B != nullptr) || desc->nB == 0, POP_ERR_ARG, "pop_arrowhead_upload: nB>0 but B is null");
    POP_REQUIRE((C != nullptr) || desc->nC == 0, POP_ERR_ARG, "pop_arrowhead_upload: nC>0 but C is null");
    const int m = desc->m, nh = desc->nh, q = desc->q, nbands = desc->lD + desc->uD + 1;
    // slots that address hats outside [0,nh) must be empty
    for (int pass = 0; pass < 2; ++pass) {
        const double *S = pass ? C : B;
        int nb = pass ? desc->nC : desc->nB;
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < m; ++k) {
                    int i = k + t - 1;
                    if ((i < 0 || i >= nh) && S[((size_t)b * 3 + t) * m + k] != 0.0) {
                        pop_set_error("pop_arrowhead_upload: %s[%d][%d][%d] couples to non-existent hat %d", pass ? "C" : "B", b, t, k, i);
                        return POP_ERR_STRUCT;
                    }
                }
    }
    // dedupe D when every element carries the same block (Fill(D, m))
    int mD_in = desc->uniform ? 1 : m;
    int mD = mD_in;
    if (mD_in == m && m > 1) {
        bool same = true;
        for (size_t r = 0; same && r < (size_t)nbands * q; ++r)
            for (int k = 1; k < m; ++k)
                if (D[r * m + k] != D[r * m]) { same = false; break; }
        if (same) mD = 1;
    }
    // D entries outside the block must be zero
    for (int d = -desc->lD; d <= desc->uD; ++d)
        for (int j = 0; j < q; ++j)
            if (j + d < 0 || j + d >= q)
                for (int k = 0; k < mD_in; ++k)
                    if (D[((size_t)(desc->lD + d) * q + j) * mD_in + k] != 0.0) {
                        pop_set_error("pop_arrowhead_upload: D band %d row %d lies outside the block and must be 0", d, j);
                        return POP_ERR_STRUCT;
                    }
    pop_arrowhead *h = nullptr;
    rc = pop_alloc_handles(ctx, *desc, mD, 1, &h);
    if (rc) return rc;
    ArrowSizes s = pop_sizes(h->d, mD);
    std::vector<double> stage(s.total, 0.0);
    double *p = stage.data();
    std::copy(Ad, Ad + nh, p); p += s.Ad;
    if (Au && nh > 1) std::copy(Au, Au + nh - 1, p);
    p += s.Au;
    if (Al && nh > 1) std::copy(Al, Al + nh - 1, p);
    p += s.Al;
    if (B) std::copy(B, B + (size_t)desc->nB * 3 * m, p);
    p += s.B;
    if (C) std::copy(C, C + (size_t)desc->nC * 3 * m, p);
    p += s.C;
    if (mD == mD_in) {
        if (D) std::copy(D, D + (size_t)nbands * q * mD, p);
    } else {
        for (size_t r = 0; r < (size_t)nbands * q; ++r) p[r] = D[r * m];
    }
    for (int b = 0; b < desc->nB; ++b)
        for (int t = 0; t < 3; ++t)
            for (int k = 0; k < m; ++k)
                if (B[((size_t)b * 3 + t) * m + k] != 0.0) h->bmask |= 1 << t;
    h->band1_zero = true;
    for (int dd = -1; dd <= 1; dd += 2)
        if (dd >= -desc->lD && dd <= desc->uD)
            for (size_t r = 0; r < (size_t)q * mD_in; ++r)
                if (D[(size_t)(desc->lD + dd) * q * mD_in + r] != 0.0) h->band1_zero = false;
    cudaError_t e = cudaMemcpyAsync(h->slab->dptr, stage.data(), s.total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        pop_arrowhead_free(ctx, h);
        pop_set_error("upload failed: %s", cudaGetErrorString(e));
        return POP_ERR_CUDA;
    }
    *out = h;
    return POP_OK;
}

extern "C" int pop_arrowhead_desc(const pop_arrowhead *A, pop_desc *desc) {
    POP_REQUIRE(A && desc, POP_ERR_ARG, "pop_arrowhead_desc: null argument");
    *desc = A->d;
    desc->uniform = A->mD == 1 ? 1 : 0;
    return POP_OK;
}

extern "C" int pop_arrowhead_download(pop_ctx *ctx, const pop_arrowhead *A, int expand_uniform, double *Ad, double *Au,
                                      double *Al, double *B, double *C, double *D) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_arrowhead_download: null argument");
    ArrowSizes s = pop_sizes(A->d, A->mD);
    std::vector<double> stage(s.total);
    POP_CUDA(cudaMemcpyAsync(stage.data(), A->slab ? (void *)A->p.Ad : nullptr, s.total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    const pop_desc &d = A->d;
    const double *p = stage.data();
    if (Ad) std::copy(p, p + d.nh, Ad);
    p += s.Ad;
    if (Au && d.nh > 1) std::copy(p, p + d.nh - 1, Au);
    p += s.Au;
    if (Al && d.nh > 1) std::copy(p, p + d.nh - 1, Al);
    p += s.Al;
    if (B) std::copy(p, p + (size_t)d.nB * 3 * d.m, B);
    p += s.B;
    if (C) std::copy(p, p + (size_t)d.nC * 3 * d.m, C);
    p += s.C;
    if (D) {
        size_t rows = (size_t)(d.lD + d.uD + 1) * d.q;
        if (A->mD == 1 && expand_uniform) {
            for (size_t r = 0; r < rows; ++r)
                for (int k = 0; k < d.m; ++k) D[r * d.m + k] = p[r];
        } else {
            std::copy(p, p + rows * A->mD, D);
        }
    }
    return POP_OK;
}

extern "C" int pop_arrowhead_free(pop_ctx *ctx, pop_arrowhead *A) {
    if (!A) return POP_OK;
    if (A->slab && --A->slab->refs == 0) {
        if (ctx) {
            cudaFreeAsync(A->slab->dptr, ctx->stream);   // ordered after the kernels that still use it
        } else {
            cudaDeviceSynchronize();
            cudaFree(A->slab->dptr);
        }
        delete A->slab;
    }
    delete A;
    return POP_OK;
}

extern "C" int pop_arrowhead_copy(pop_ctx *ctx, const pop_arrowhead *A, pop_arrowhead **out) {
    POP_REQUIRE(ctx && A && out, POP_ERR_ARG, "pop_arrowhead_copy: null argument");
    pop_arrowhead *h = nullptr;
    int rc = pop_alloc_handles(ctx, A->d, A->mD, 1, &h);
    if (rc) return rc;
    ArrowSizes s = pop_sizes(A->d, A->mD);
    POP_CUDA(cudaMemcpyAsync(h->p.Ad, A->p.Ad, s.total * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    h->has_recip = A->has_recip;
    h->is_factor = A->is_factor;
    h->bmask = A->bmask;
    h->band1_zero = A->band1_zero;
    *out = h;
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// operator views
// ------------------------------------------------------------------------------------
int pop_make_triview(const pop_arrowhead *A, int uplo, int trans, int unit, TriView *tv) {
    const pop_desc &d = A->d;
    const size_t plane = (size_t)d.q * A->mD;
    tv->m = d.m; tv->nh = d.nh; tv->q = d.q; tv->mD = A->mD;
    tv->unit = unit;
    tv->rdA = A->p.rdA;
    tv->rdD = A->p.rdD;
    for (int i = 0; i < POP_MAXD; ++i) tv->W[i] = nullptr;
    if (uplo == POP_UPPER) {
        tv->offA = A->p.Au;
        tv->S = A->p.B;
        tv->nb = std::min(d.nB, d.q);
        tv->nd = d.uD;
        for (int dd = 1; dd <= d.uD; ++dd) tv->W[dd - 1] = A->p.D + (size_t)(d.lD + dd) * plane;
    } else {
        tv->offA = A->p.Al;
        tv->S = A->p.C;
        tv->nb = std::min(d.nC, d.q);
        tv->nd = d.lD;
        // L[i+dd, i] = D[lD-dd][i+dd]
        for (int dd = 1; dd <= d.lD; ++dd) tv->W[dd - 1] = A->p.D + (size_t)(d.lD - dd) * plane + (size_t)dd * A->mD;
    }
    (void)trans;
    return POP_OK;
}

int pop_make_mulview(const pop_arrowhead *A, int sym, int trans, MulView *mv) {
    const pop_desc &d = A->d;
    const ptrdiff_t plane = (ptrdiff_t)d.q * A->mD;
    const int mD = A->mD;
    mv->m = d.m; mv->nh = d.nh; mv->q = d.q; mv->mD = mD;
    mv->Ad = A->p.Ad;
    for (int i = 0; i < 2 * POP_MAXD + 1; ++i) mv->Dp[i] = nullptr;
    if (sym == POP_SYM_UPPER) {
        mv->Aup = A->p.Au; mv->Alo = A->p.Au;
        mv->Srow = A->p.B; mv->Scol = A->p.B;
        mv->nb_row = mv->nb_col = std::min(d.nB, d.q);
        mv->lD = mv->uD = d.uD;
        for (int dd = 0; dd <= d.uD; ++dd) {
            mv->Dp[POP_MAXD + dd] = A->p.D + (ptrdiff_t)(d.lD + dd) * plane;
            mv->Dp[POP_MAXD - dd] = A->p.D + (ptrdiff_t)(d.lD + dd) * plane - (ptrdiff_t)dd * mD;
        }
    } else if (!trans) {
        mv->Aup = A->p.Au; mv->Alo = A->p.Al;
        mv->Srow = A->p.B; mv->Scol = A->p.C;
        mv->nb_row = std::min(d.nB, d.q); mv->nb_col = std::min(d.nC, d.q);
        mv->lD = d.lD; mv->uD = d.uD;
        for (int dd = -d.lD; dd <= d.uD; ++dd) mv->Dp[POP_MAXD + dd] = A->p.D + (ptrdiff_t)(d.lD + dd) * plane;
    } else {
        mv->Aup = A->p.Al; mv->Alo = A->p.Au;
        mv->Srow = A->p.C; mv->Scol = A->p.B;
        mv->nb_row = std::min(d.nC, d.q); mv->nb_col = std::min(d.nB, d.q);
        mv->lD = d.uD; mv->uD = d.lD;
        // (A^T)[j, j+dd] = A[j+dd, j] = D[lD-dd][j+dd]
        for (int dd = -d.uD; dd <= d.lD; ++dd)
            mv->Dp[POP_MAXD + dd] = A->p.D + (ptrdiff_t)(d.lD - dd) * plane + (ptrdiff_t)dd * mD;
    }
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// small kernels: reciprocals, assembly, linear combinations
// ------------------------------------------------------------------------------------
#define LAUNCH_COUNT(ctx) ((ctx)->launches++)

__global__ void k_recip(ArrowPtrs p, int nh, int lD, size_t plane) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)nh) p.rdA[i] = 1.0 / p.Ad[i];
    if (i < plane) p.rdD[i] = 1.0 / p.D[(size_t)lD * plane + i];
}

// The upper data of A already hold a reverse-Cholesky factor U (S = U U'), e.g. one the reference computed on the host:
// `F.U` of a ReverseCholesky object.  Marks the handle as a factor for pop_solve / pop_solve_mul_axpy (reciprocal pivots
// are computed once).  Non-positive or non-finite pivots are rejected like a failed factorisation (1-based index).
__global__ void k_check_pivots(ArrowPtrs p, int nh, size_t lD_off, size_t plane, int *info) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)nh && !(p.Ad[i] > 0.0) ) atomicMin(info, (int)i + 1);
    if (i < plane && !(p.D[lD_off + i] > 0.0)) atomicMin(info, nh + (int)i + 1);
}
extern "C" int pop_arrowhead_as_factor(pop_ctx *ctx, pop_arrowhead *A) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_arrowhead_as_factor: null argument");
    const size_t plane = (size_t)A->d.q * A->mD;
    const size_t n = std::max<size_t>(plane, (size_t)A->d.nh);
    int big = INT_MAX;
    POP_CUDA(cudaMemcpyAsync(ctx->d_info, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        k_check_pivots<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(A->p, A->d.nh, (size_t)A->d.lD * plane, plane, ctx->d_info);
        LAUNCH_COUNT(ctx);
    }
    POP_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_info[0] != INT_MAX) {
        pop_set_error("pop_arrowhead_as_factor: pivot %d of the factor is not positive", ctx->h_info[0]);
        return ctx->h_info[0];
    }
    A->has_recip = false;
    int rc = pop_ensure_recip(ctx, A);
    if (rc) return rc;
    A->is_factor = true;
    return POP_OK;
}

int pop_ensure_recip(pop_ctx *ctx, pop_arrowhead *A) {
    if (A->has_recip) return POP_OK;
    size_t plane = (size_t)A->d.q * A->mD;
    size_t n = std::max<size_t>(plane, A->d.nh);
    k_recip<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(A->p, A->d.nh, A->d.lD, plane);
    LAUNCH_COUNT(ctx);
    POP_CUDA(cudaGetLastError());
    A->has_recip = true;
    return POP_OK;
}

// Closed forms on a uniform mesh of step h (upper data only).
//   mass: src/continuouspolynomial.jl:274-290 (C1), src/dirichlet.jl:180-196 (Dirichlet)
//   lap : src/continuouspolynomial.jl:348-357 (C1), src/dirichlet.jl:169-178 (Dirichlet)
__global__ void k_assemble(ArrowPtrs M, ArrowPtrs L, int basis, int m, int nh, int q, double h) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double si = 1.0 / h;
    if (i < nh) {
        bool end = (basis == POP_BASIS_C1) && (i == 0 || i == nh - 1);
        if (M.Ad) M.Ad[i] = end ? h / 3 : 2 * h / 3;
        if (L.Ad) L.Ad[i] = end ? -si : -2 * si;
        if (i < nh - 1) {
            if (M.Au) M.Au[i] = h / 6;
            if (L.Au) L.Au[i] = si;
        }
    }
    if (i < m && M.B) {
        // left hat of element i: C1 -> hat i (slot 1); Dirichlet -> hat i-1 (slot 0)
        int tl = basis == POP_BASIS_C1 ? 1 : 0;
        int il = i + tl - 1, ir = il + 1;
        if (il >= 0 && il < nh) {
            M.B[(0 * 3 + tl) * (size_t)m + i] = h / 6;
            M.B[(1 * 3 + tl) * (size_t)m + i] = -h / 30;
        }
        if (ir >= 0 && ir < nh) {
            M.B[(0 * 3 + tl + 1) * (size_t)m + i] = h / 6;
            M.B[(1 * 3 + tl + 1) * (size_t)m + i] = h / 30;
        }
    }
    if (i < q) {
        double j1 = i + 1.0;
        if (M.D) {
            M.D[i] = (h / 2) * 4.0 / ((2 * j1 - 1) * (2 * j1 + 1) * (2 * j1 + 3));
            M.D[q + i] = 0.0;
            M.D[2 * (size_t)q + i] = (i + 2 < q) ? (h / 2) * (-2.0) / ((2 * j1 + 1) * (2 * j1 + 3) * (2 * j1 + 5)) : 0.0;
        }
        if (L.D) L.D[i] = -4.0 / (h * (2 * j1 + 1));
    }
}

extern "C" int pop_assemble(pop_ctx *ctx, int basis, int m, int N, double h, pop_arrowhead **mass, pop_arrowhead **lap) {
    POP_REQUIRE(ctx, POP_ERR_ARG, "pop_assemble: null ctx");
    POP_REQUIRE(basis == POP_BASIS_C1 || basis == POP_BASIS_DIRICHLET, POP_ERR_ARG, "pop_assemble: unknown basis %d", basis);
    POP_REQUIRE(N >= 1 && h > 0, POP_ERR_ARG, "pop_assemble: need N>=1, h>0");
    int nh = basis == POP_BASIS_C1 ? m + 1 : m - 1;
    POP_REQUIRE(m >= 1 && nh >= 1, POP_ERR_DIM, "pop_assemble: mesh of %d elements has no hat functions for this basis", m);
    int q = N - 1;
    pop_arrowhead *hm = nullptr, *hl = nullptr;
    ArrowPtrs pm = {}, pl = {};
    if (mass) {
        pop_desc d = {m, nh, q, 2, 0, 0, 2, 1};
        int rc = pop_alloc_handles(ctx, d, 1, 1, &hm);
        if (rc) return rc;
        pm = hm->p;
    }
    if (lap) {
        pop_desc d = {m, nh, q, 0, 0, 0, 0, 1};
        int rc = pop_alloc_handles(ctx, d, 1, 1, &hl);
        if (rc) { pop_arrowhead_free(ctx, hm); return rc; }
        pl = hl->p;
    }
    int nthr = std::max(std::max(nh, m), q);
    k_assemble<<<(nthr + 255) / 256, 256, 0, ctx->stream>>>(pm, pl, basis, m, nh, q, h);
    LAUNCH_COUNT(ctx);
    POP_CUDA(cudaGetLastError());
    if (mass) { hm->bmask = basis == POP_BASIS_C1 ? 6 : 3; hm->band1_zero = true; *mass = hm; }
    if (lap) { hl->band1_zero = true; *lap = hl; }
    return POP_OK;
}

// Non-uniform mesh: element k has length hk[k]; one D block per element (mD = m), element index fastest.
// The element integrals are those of the uniform closed forms with h -> h_k, i.e. what the reference's generic
// L' * grammatrix(P) * L (src/continuouspolynomial.jl:259-265,293-297) and -(diff C)' G (diff C)
// (src/continuouspolynomial.jl:318-325) evaluate to.
// ak (may be null = 1): diffusion coefficient of element k, the piecewise-constant case of `-(D*C)' * (a .* (D*C))`
// (examples/heat.jl:77-79): every element integral of the weak Laplacian is scaled by a_k.
__global__ void k_assemble_mesh(ArrowPtrs M, ArrowPtrs L, int basis, int m, int nh, int q, const double *__restrict__ hk, const double *__restrict__ ak) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nh) {
        // hat i sits on node i (C1) or node i+1 (Dirichlet): elements node-1 and node
        const int node = basis == POP_BASIS_C1 ? i : i + 1;
        const double hl = node >= 1 ? hk[node - 1] : 0.0, hr = node < m ? hk[node] : 0.0;
        if (M.Ad) M.Ad[i] = (hl + hr) / 3;
        const double al = (ak && node >= 1) ? ak[node - 1] : 1.0, ar = (ak && node < m) ? ak[node] : 1.0;
        if (L.Ad) L.Ad[i] = -((node >= 1 ? al / hl : 0.0) + (node < m ? ar / hr : 0.0));
        if (i < nh - 1) {   // hats i and i+1 share element `node`
            if (M.Au) M.Au[i] = hr / 6;
            if (L.Au) L.Au[i] = ar / hr;
        }
    }
    if (i < m && M.B) {
        const double h = hk[i];
        const int tl = basis == POP_BASIS_C1 ? 1 : 0;
        const int il = i + tl - 1, ir = il + 1;
        if (il >= 0 && il < nh) {
            M.B[(0 * 3 + tl) * (size_t)m + i] = h / 6;
            M.B[(1 * 3 + tl) * (size_t)m + i] = -h / 30;
        }
        if (ir >= 0 && ir < nh) {
            M.B[(0 * 3 + tl + 1) * (size_t)m + i] = h / 6;
            M.B[(1 * 3 + tl + 1) * (size_t)m + i] = h / 30;
        }
    }
    // D[band][j][k]: one thread per (j, k)
    const size_t plane = (size_t)q * m;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e / m), k = (int)(e % m);
        const double h = hk[k], j1 = j + 1.0;
        if (M.D) {
            M.D[e] = (h / 2) * 4.0 / ((2 * j1 - 1) * (2 * j1 + 1) * (2 * j1 + 3));
            M.D[plane + e] = 0.0;
            M.D[2 * plane + e] = (j + 2 < q) ? (h / 2) * (-2.0) / ((2 * j1 + 1) * (2 * j1 + 3) * (2 * j1 + 5)) : 0.0;
        }
        if (L.D) L.D[e] = -4.0 * (ak ? ak[k] : 1.0) / (h * (2 * j1 + 1));
    }
}

extern "C" int pop_assemble_mesh(pop_ctx *ctx, int basis, int m, int N, const double *h_elems, pop_arrowhead **mass, pop_arrowhead **lap) {
    return pop_assemble_mesh_coef(ctx, basis, m, N, h_elems, nullptr, mass, lap);
}

extern "C" int pop_assemble_mesh_coef(pop_ctx *ctx, int basis, int m, int N, const double *h_elems, const double *a_elems, pop_arrowhead **mass,
                                      pop_arrowhead **lap) {
    POP_REQUIRE(ctx && h_elems, POP_ERR_ARG, "pop_assemble_mesh: null argument");
    POP_REQUIRE(basis == POP_BASIS_C1 || basis == POP_BASIS_DIRICHLET, POP_ERR_ARG, "pop_assemble_mesh: unknown basis %d", basis);
    POP_REQUIRE(N >= 1, POP_ERR_ARG, "pop_assemble_mesh: need N>=1");
    const int nh = basis == POP_BASIS_C1 ? m + 1 : m - 1;
    POP_REQUIRE(m >= 1 && nh >= 1, POP_ERR_DIM, "pop_assemble_mesh: mesh of %d elements has no hat functions for this basis", m);
    for (int k = 0; k < m; ++k) POP_REQUIRE(h_elems[k] > 0, POP_ERR_ARG, "pop_assemble_mesh: element %d has non-positive length (points must be increasing)", k);
    const int q = N - 1;
    double *dh = nullptr, *da = nullptr;
    int rc = pop_ctx_scratch(ctx, (size_t)2 * m * sizeof(double), (void **)&dh);
    if (rc) return rc;
    POP_CUDA(cudaMemcpyAsync(dh, h_elems, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (a_elems) {
        for (int k = 0; k < m; ++k) POP_REQUIRE(a_elems[k] > 0, POP_ERR_ARG, "pop_assemble_mesh: element %d has a non-positive diffusion coefficient", k);
        da = dh + m;
        POP_CUDA(cudaMemcpyAsync(da, a_elems, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    pop_arrowhead *hm = nullptr, *hl = nullptr;
    ArrowPtrs pm = {}, pl = {};
    if (mass) {
        pop_desc d = {m, nh, q, 2, 0, 0, 2, 0};
        rc = pop_alloc_handles(ctx, d, m, 1, &hm);
        if (rc) return rc;
        pm = hm->p;
    }
    if (lap) {
        pop_desc d = {m, nh, q, 0, 0, 0, 0, 0};
        rc = pop_alloc_handles(ctx, d, m, 1, &hl);
        if (rc) { pop_arrowhead_free(ctx, hm); return rc; }
        pl = hl->p;
    }
    const size_t work = std::max<size_t>((size_t)q * m, (size_t)std::max(nh, m));
    const int blocks = (int)std::min<size_t>((work + 255) / 256, 4096);
    k_assemble_mesh<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(pm, pl, basis, m, nh, q, dh, da);
    LAUNCH_COUNT(ctx);
    POP_CUDA(cudaGetLastError());
    POP_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller's h_elems and the scratch copy may be reused at once
    if (mass) { hm->bmask = basis == POP_BASIS_C1 ? 6 : 3; hm->band1_zero = true; *mass = hm; }
    if (lap) { hl->band1_zero = true; *lap = hl; }
    return POP_OK;
}

struct BatchPtrs {  // device-side array element for batched outputs
    ArrowPtrs p;
};

// out[e] = alpha_s * (e < nx ? x[e] : 0) + beta_s * (e < ny ? y[e] : 0)   (ragged prefix fields)
__global__ void k_axpby_prefix(const BatchPtrs *outs, int field, const double *x, size_t nx, const double *y, size_t ny,
                               size_t nout, const double *alpha, const double *beta) {
    int s = blockIdx.y;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nout) return;
    const ArrowPtrs &o = outs[s].p;
    double *dst = field == 0 ? o.Ad : field == 1 ? o.Au : field == 2 ? o.Al : field == 3 ? o.B : o.C;
    double a = alpha[s], b = beta[s];
    double v = 0.0;
    if (e < nx) v += a * x[e];
    if (e < ny) v += b * y[e];
    dst[e] = v;
}

__global__ void k_axpby_D(const BatchPtrs *outs, const double *xD, int lDx, int uDx, int mDx, const double *yD, int lDy,
                          int uDy, int mDy, int lDo, int uDo, int mDo, int q, const double *alpha, const double *beta) {
    int s = blockIdx.y;
    size_t plane = (size_t)q * mDo;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= plane * (size_t)(lDo + uDo + 1)) return;
    int band = (int)(e / plane) - lDo;
    size_t r = e % plane;
    int j = (int)(r / mDo), k = (int)(r % mDo);
    double v = 0.0;
    if (band >= -lDx && band <= uDx) v += alpha[s] * xD[((size_t)(lDx + band) * q + j) * mDx + (mDx == 1 ? 0 : k)];
    if (band >= -lDy && band <= uDy) v += beta[s] * yD[((size_t)(lDy + band) * q + j) * mDy + (mDy == 1 ? 0 : k)];
    outs[s].p.D[e] = v;
}

int pop_launch_axpby_batch(pop_ctx *ctx, const pop_arrowhead *X, const pop_arrowhead *Y, const double *h_alpha,
                           const double *h_beta, int count, pop_arrowhead **out) {
    POP_REQUIRE(X->d.m == Y->d.m && X->d.nh == Y->d.nh && X->d.q == Y->d.q, POP_ERR_DIM,
                "arrowhead +/-: sizes differ (m %d/%d, nh %d/%d, q %d/%d)", X->d.m, Y->d.m, X->d.nh, Y->d.nh, X->d.q, Y->d.q);
    POP_REQUIRE(count >= 1 && count <= 65535, POP_ERR_ARG, "batch count %d out of range", count);
    pop_desc d = X->d;
    d.nB = std::max(X->d.nB, Y->d.nB);
    d.nC = std::max(X->d.nC, Y->d.nC);
    d.lD = std::max(X->d.lD, Y->d.lD);
    d.uD = std::max(X->d.uD, Y->d.uD);
    int mD = std::max(X->mD, Y->mD);
    d.uniform = mD == 1;
    int rc = pop_alloc_handles(ctx, d, mD, count, out);
    if (rc) return rc;
    for (int i = 0; i < count; ++i) {
        out[i]->band1_zero = X->band1_zero && Y->band1_zero;
        out[i]->bmask = X->bmask | Y->bmask;
    }
    // stage [alpha | beta | BatchPtrs] in scratch
    size_t bytes = (size_t)count * (2 * sizeof(double) + sizeof(BatchPtrs));
    std::vector<char> host(bytes);
    double *ha = (double *)host.data();
    double *hb = ha + count;
    BatchPtrs *hp = (BatchPtrs *)(hb + count);
    for (int i = 0; i < count; ++i) {
        ha[i] = h_alpha[i];
        hb[i] = h_beta[i];
        hp[i].p = out[i]->p;
    }
    void *dscr = nullptr;
    rc = pop_ctx_scratch(ctx, bytes, &dscr);
    if (rc) return rc;
    POP_CUDA(cudaMemcpyAsync(dscr, host.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));  // host staging vector goes out of scope
    const double *da = (const double *)dscr, *db = da + count;
    const BatchPtrs *dp = (const BatchPtrs *)(db + count);
    const int m = d.m, nh = d.nh;
    struct F { int id; const double *x; size_t nx; const double *y; size_t ny; size_t nout; };
    F fields[5] = {
        {0, X->p.Ad, (size_t)nh, Y->p.Ad, (size_t)nh, (size_t)nh},
        {1, X->p.Au, (size_t)nh - 1, Y->p.Au, (size_t)nh - 1, (size_t)nh - 1},
        {2, X->p.Al, (size_t)nh - 1, Y->p.Al, (size_t)nh - 1, (size_t)nh - 1},
        {3, X->p.B, (size_t)X->d.nB * 3 * m, Y->p.B, (size_t)Y->d.nB * 3 * m, (size_t)d.nB * 3 * m},
        {4, X->p.C, (size_t)X->d.nC * 3 * m, Y->p.C, (size_t)Y->d.nC * 3 * m, (size_t)d.nC * 3 * m},
    };
    for (const F &f : fields) {
        if (f.nout == 0) continue;
        dim3 grid((unsigned)((f.nout + 255) / 256), count);
        k_axpby_prefix<<<grid, 256, 0, ctx->stream>>>(dp, f.id, f.x, f.nx, f.y, f.ny, f.nout, da, db);
        LAUNCH_COUNT(ctx);
    }
    size_t nD = (size_t)(d.lD + d.uD + 1) * d.q * mD;
    if (nD) {
        dim3 grid((unsigned)((nD + 255) / 256), count);
        k_axpby_D<<<grid, 256, 0, ctx->stream>>>(dp, X->p.D, X->d.lD, X->d.uD, X->mD, Y->p.D, Y->d.lD, Y->d.uD, Y->mD, d.lD,
                                                d.uD, mD, d.q, da, db);
        LAUNCH_COUNT(ctx);
    }
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

extern "C" int pop_axpby(pop_ctx *ctx, double alpha, const pop_arrowhead *X, double beta, const pop_arrowhead *Y,
                         pop_arrowhead **out) {
    POP_REQUIRE(ctx && X && Y && out, POP_ERR_ARG, "pop_axpby: null argument");
    return pop_launch_axpby_batch(ctx, X, Y, &alpha, &beta, 1, out);
}

// ------------------------------------------------------------------------------------
// factorisation entry points (kernels in pop_factor.cu)
// ------------------------------------------------------------------------------------
static int check_factor_structure(const pop_arrowhead *S) {
    const pop_desc &d = S->d;
    if (std::min(d.nB, d.q) > 0) {
        // the reference supports B bandwidths (1,0) with nh == m+1 (src/arrowhead.jl:299,305)
        // or (0,1) with nh == m-1 (:348,356)
        POP_REQUIRE(d.nh == d.m + 1 || d.nh == d.m - 1, POP_ERR_STRUCT,
                    "reversecholesky: need nh == m+1 (lower B) or nh == m-1 (upper B), got nh=%d m=%d", d.nh, d.m);
        POP_REQUIRE(!((S->bmask & 1) && (S->bmask & 4)), POP_ERR_STRUCT,
                    "reversecholesky: B blocks must have bandwidths (1,0) or (0,1) (src/arrowhead.jl:299,348)");
    }
    return POP_OK;
}

extern "C" int pop_revchol_inplace(pop_ctx *ctx, pop_arrowhead *S) {
    POP_REQUIRE(ctx && S, POP_ERR_ARG, "pop_revchol_inplace: null argument");
    int rc = check_factor_structure(S);
    if (rc) return rc;
    int32_t info = 0;
    rc = pop_launch_factor(ctx, &S, 1, &info);
    if (rc) return rc;
    if (info > 0) pop_set_error("reversecholesky: matrix is not positive definite (pivot %d)", info);
    return info;
}

extern "C" int pop_revchol(pop_ctx *ctx, const pop_arrowhead *S, pop_arrowhead **U) {
    POP_REQUIRE(ctx && S && U, POP_ERR_ARG, "pop_revchol: null argument");
    *U = nullptr;
    int rc = check_factor_structure(S);
    if (rc) return rc;
    // reversecholcopy (src/arrowhead.jl:266-272): upper data only -> lD = 0, no C
    double one = 1.0, zero = 0.0;
    pop_arrowhead tmp = *S;
    tmp.d.nC = 0;
    pop_arrowhead *h = nullptr;
    // copy the upper bands through the batched combiner: 1*S_upper + 0*S_upper
    pop_arrowhead up = tmp;
    up.d.lD = 0;
    up.p.D = S->p.D + (size_t)S->d.lD * S->d.q * S->mD;
    rc = pop_launch_axpby_batch(ctx, &up, &up, &one, &zero, 1, &h);
    if (rc) return rc;
    int32_t info = 0;
    rc = pop_launch_factor(ctx, &h, 1, &info);
    if (rc) { pop_arrowhead_free(ctx, h); return rc; }
    if (info > 0) {
        pop_arrowhead_free(ctx, h);
        pop_set_error("reversecholesky: matrix is not positive definite (pivot %d)", info);
        return info;
    }
    *U = h;
    return POP_OK;
}

extern "C" int pop_revchol_shifted_batch(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, const double *alpha,
                                         const double *sigma, int nshift, pop_arrowhead **U, int32_t *info) {
    POP_REQUIRE(ctx && K && M && alpha && sigma && U, POP_ERR_ARG, "pop_revchol_shifted_batch: null argument");
    POP_REQUIRE(nshift >= 1, POP_ERR_ARG, "pop_revchol_shifted_batch: nshift must be >= 1");
    pop_arrowhead Ku = *K, Mu = *M;  // upper-data views
    Ku.d.nC = 0; Mu.d.nC = 0;
    Ku.p.D = K->p.D + (size_t)K->d.lD * K->d.q * K->mD; Ku.d.lD = 0;
    Mu.p.D = M->p.D + (size_t)M->d.lD * M->d.q * M->mD; Mu.d.lD = 0;
    int done = 0, first_bad = 0;
    while (done < nshift) {
        int cnt = std::min(nshift - done, 2048);
        int rc = pop_launch_axpby_batch(ctx, &Ku, &Mu, alpha + done, sigma + done, cnt, U + done);
        if (rc) return rc;
        rc = check_factor_structure(U[done]);
        if (rc) return rc;
        std::vector<int32_t> inf(cnt, 0);
        rc = pop_launch_factor(ctx, U + done, cnt, inf.data());
        if (rc) return rc;
        for (int i = 0; i < cnt; ++i) {
            if (info) info[done + i] = inf[i];
            if (inf[i] > 0 && !first_bad) first_bad = inf[i];
        }
        done += cnt;
    }
    if (first_bad) pop_set_error("reversecholesky (batched): a shifted matrix is not positive definite (pivot %d)", first_bad);
    return first_bad;
}
