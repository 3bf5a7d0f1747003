// Host side of the one-pass row half step (kernel: pop_row1.cuh): plan (chain class, rows per block, cluster
// size, persistent grid) and launch.  Used by the element-decomposed ADI driver (pop_row.cu) when ONE GPU holds
// the whole matrix; with several GPUs the hat exchange crosses NVLink and the two-pass kernels of pop_row.cu run.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "pop_row.h"
#include "pop_row1.cuh"

static r1_kernel_t r1_kernel(int jmax, int rb, int npar, bool has1) {
    switch (jmax) {
        case 16: return r1_kernel_r16(rb, npar, has1);
        case 32: return r1_kernel_r32(rb, npar, has1);
        case 48: return r1_kernel_r48(rb, npar, has1);
        case 64: return r1_kernel_r64(rb, npar, has1);
    }
    return nullptr;
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int row1_launch(pop_ctx *ctx, const RowOp &op, const double *F, double *R);

int pop_row1_half_step(pop_ctx *ctx, const RowOp &op, const double *F, double *R) {
    // POP_ROW1: 0 = never, 1 (default) = plain solves only, 2 = also the fused solve + product half step.  Measured on
    // B200 (profiles/r1_row1.txt): the plain row solve gains 1.4x over the two-pass kernels, the half step with the
    // product does not (the register-resident block leaves one CTA per SM and the phases of a block run back to back).
    const int mode = env_int("POP_ROW1", 1);
    if (op.world != 1 || mode == 0 || (op.nbT >= 0 && mode < 2)) return POP_ROW1_UNSUPPORTED;
    const int rc = row1_launch(ctx, op, F, R);
    if (rc == POP_ROW1_UNSUPPORTED && env_int("POP_ROW1_STRICT", 0) != 0) {
        // tests force a plan (POP_ROW1_CS / POP_ROW1_RB): falling back silently would test the wrong kernel
        pop_set_error("pop_row1: the forced plan does not apply to m=%d q=%d", op.m, op.q);
        return POP_ERR_DIM;
    }
    return rc;
}

static int row1_launch(pop_ctx *ctx, const RowOp &op, const double *F, double *R) {
    if (op.q < 1 || op.q > 64 || op.nb > R1_NBMAX || op.nbT > R1_NBMAX) return POP_ROW1_UNSUPPORTED;
    const int rows = op.q <= 16 ? 16 : op.q <= 32 ? 32 : op.q <= 48 ? 48 : 64;
    // two threads per (element, row) when the even / odd degree chains decouple (POP_ROW1_NPAR=1 forces one)
    // (measured slower than one thread per chain at the ADI shape, so opt-in)
    const int npar = (op.has1 == 0 && op.q >= 2 && env_int("POP_ROW1_NPAR", 1) == 2) ? 2 : 1;
    const int tmax = r1_tmax(rows / npar);
    const int want_cs = env_int("POP_ROW1_CS", 0), want_rb = env_int("POP_ROW1_RB", 0);
    int rb = 0, cs = 0;
    for (int cand_rb = 16; cand_rb >= 8 && !rb; cand_rb >>= 1) {
        if (want_rb && cand_rb != want_rb) continue;
        for (int c = 1; c <= 8; c <<= 1) {
            if (want_cs && c != want_cs) continue;
            if (op.m % c || op.m / c < 1) continue;
            if ((op.m / c) * cand_rb * npar > tmax) continue;
            rb = cand_rb;
            cs = c;
            break;
        }
    }
    if (!rb) return POP_ROW1_UNSUPPORTED;
    // the kernel addresses the degrees of an element with 32-bit offsets
    if ((int64_t)rows * op.m * op.ld >= ((int64_t)1 << 31) || (int64_t)rows * op.m * std::max<int64_t>(op.ldf, 1) >= ((int64_t)1 << 31)) return POP_ROW1_UNSUPPORTED;
    r1_kernel_t kern = r1_kernel(rows, rb, npar, op.has1 != 0);
    if (!kern) return POP_ROW1_UNSUPPORTED;
    const int mlc = op.m / cs;
    const int nthr = (npar * mlc * rb + 31) & ~31;
    const int hcap = mlc + 2;
    const bool useF = op.nbT >= 0 && F != nullptr && op.gamma != 0.0;
    bool stage_f = useF && env_int("POP_ROW1_STAGE_F", 1) != 0;
    R1Smem L = r1_smem(rows, rb, mlc, hcap, nthr, stage_f, npar);
    if (stage_f && (size_t)L.total * 8 + R1_BAR_BYTES > (size_t)ctx->max_smem_optin) {
        stage_f = false;
        L = r1_smem(rows, rb, mlc, hcap, nthr, false, npar);
    }
    const size_t smem = (size_t)L.total * 8 + R1_BAR_BYTES;
    if (smem > (size_t)ctx->max_smem_optin) return POP_ROW1_UNSUPPORTED;
    POP_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    R1Params P;
    P.rd = op.rd; P.w1 = op.w1; P.w2 = op.w2; P.S = op.S; P.rdA = op.rdA; P.offA = op.offA; P.nb = op.nb;
    P.t0 = op.t0; P.t1 = op.t1; P.t2 = op.t2; P.Ts = op.Ts; P.TAd = op.TAd; P.TAu = op.TAu; P.nbT = op.nbT;
    P.gamma = op.gamma;
    P.F = useF ? F : nullptr;
    P.ldf = op.ldf;
    P.X = R; P.ld = op.ld; P.nrows = op.nrows;
    P.m = op.m; P.nh = op.nh; P.q = op.q;
    P.cs = cs; P.mlc = mlc; P.hcap = hcap;
    P.nblk = (int)((op.nrows + rb - 1) / rb);
    P.stage_f = stage_f ? 1 : 0;
    P.contig = env_int("POP_ROW1_CONTIG", 1);

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3((unsigned)nthr);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent clusters: as many as are co-resident (queried once per shape)
    static std::mutex plan_mu;
    static int cached_key[7] = {0, 0, 0, 0, 0, 0, -1}, cached_ncl = 0;
    const int key[7] = {rows, rb, npar * 2 + (op.has1 != 0), cs, nthr, (int)smem, ctx->device};
    std::lock_guard<std::mutex> lk(plan_mu);
    int ncl = cached_ncl;
    if (!std::equal(key, key + 7, cached_key) || ncl <= 0) {
        cfg.gridDim = dim3((unsigned)(cs * ctx->sm_count));
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, (const void *)kern, &cfg) != cudaSuccess || nc <= 0) {
            cudaGetLastError();
            return POP_ROW1_UNSUPPORTED;
        }
        ncl = nc;
        std::copy(key, key + 7, cached_key);
        cached_ncl = ncl;
    }
    const int cap = env_int("POP_ROW1_NCL", 0);
    if (cap > 0) ncl = std::min(ncl, cap);
    ncl = std::min(ncl, P.nblk);
    cfg.gridDim = dim3((unsigned)(ncl * cs));
    if (env_int("POP_ROW1_VERBOSE", 0))
        fprintf(stderr, "pop_row1: class %d (%d per thread), %d rows/block, cluster %d x %d threads, %zu B smem, %d clusters, %d row blocks, stage_f %d\n", rows, rows / npar, rb, cs, nthr,
                smem, ncl, P.nblk, P.stage_f);
    POP_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    ctx->launches++;
    return POP_OK;
}

// X / F for a row panel X (nrhs x n, leading dimension ldx) on one GPU: rows are the systems (pop_solve, side RIGHT)
int pop_row1_solve_right(pop_ctx *ctx, const pop_arrowhead *U, double *dX, int64_t ldx, int64_t nrhs) {
    const pop_desc &u = U->d;
    if (!U->is_factor || U->mD != 1 || u.uD > 2 || abs(u.nh - u.m) > 1) return POP_ROW1_UNSUPPORTED;
    RowOp op;
    memset(&op, 0, sizeof(op));
    const size_t plane = (size_t)u.q;
    op.rd = U->p.rdD;
    op.w1 = u.uD >= 1 ? U->p.D + (size_t)(u.lD + 1) * plane : nullptr;
    op.w2 = u.uD >= 2 ? U->p.D + (size_t)(u.lD + 2) * plane : nullptr;
    op.S = U->p.B;
    op.rdA = U->p.rdA;
    op.offA = U->p.Au;
    op.nb = std::min(u.nB, u.q);
    op.nbT = -1;
    op.has1 = (u.uD >= 1 && !U->band1_zero) ? 1 : 0;
    op.m = u.m; op.nh = u.nh; op.q = u.q; op.k0 = 0; op.ml = u.m; op.hA = 0; op.nhl = u.nh; op.hB = u.nh;
    op.rank = 0; op.world = 1;
    op.nrows = nrhs; op.ld = ldx; op.ldf = 0;
    const int rc = pop_row2_half_step(ctx, op, nullptr, dX, nullptr);
    if (rc != POP_ROW1_UNSUPPORTED) return rc;
    return pop_row1_half_step(ctx, op, nullptr, dX);
}
