// Triangular solves with packed arrowheads: the "staged" (multi-kernel, global-memory)
// path.  General in every respect the reference is: any D bandwidth <= POP_MAXD,
// per-element or shared D blocks, unit / non-unit, left (columns) or right (rows).
//   U \ b : src/arrowhead.jl:425-457      L \ b : src/arrowhead.jl:458-486
// The single-pass on-chip kernel for the ReverseCholesky solve lives in pop_fused.cu.
//
// Data access: a vector entry v of batch member c sits at X[c*sc + v*sv].
//   LEFT  (columns of an n x nrhs matrix): sv = 1,  sc = ld  -> lanes run over elements k
//   RIGHT (rows of an nrhs x n matrix)   : sv = ld, sc = 1   -> lanes run over the batch
// so every warp access is a run of consecutive doubles in both cases.
#include "pop_internal.h"
#include "pop_row.h"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------
// bubble sweeps: per (element, batch member) banded substitution along the degree index
// ------------------------------------------------------------------------------------
template <bool BACKWARD, int ND, bool LANE_BATCH>
__global__ void __launch_bounds__(256) k_bubble_sweep(TriView tv, double *__restrict__ X, int64_t sv, int64_t sc, int64_t nbatch) {
    const int64_t nL = LANE_BATCH ? nbatch : tv.m;
    const int64_t nO = LANE_BATCH ? tv.m : nbatch;
    const int64_t l = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (l >= nL) return;
    const int m = tv.m, q = tv.q, mD = tv.mD, nd = tv.nd;
    for (int64_t o = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; o < nO; o += (int64_t)gridDim.y * blockDim.y) {
        const int k = (int)(LANE_BATCH ? o : l);
        const int64_t c = LANE_BATCH ? l : o;
        const int kD = mD == 1 ? 0 : k;
        double *x = X + c * sc + ((int64_t)tv.nh + k) * sv;   // entry j at x[j*m*sv]
        const int64_t js = (int64_t)m * sv;
        double w[ND > 0 ? ND : 1];
#pragma unroll
        for (int d = 0; d < ND; ++d) w[d] = 0.0;
        for (int jj = 0; jj < q; ++jj) {
            const int j = BACKWARD ? q - 1 - jj : jj;
            double v = x[j * js];
#pragma unroll
            for (int d = 1; d <= ND; ++d) {
                if (d <= nd && jj >= d) {
                    // BACKWARD: couples j and j+d -> W_d(j);  FORWARD: couples j-d and j -> W_d(j-d)
                    const int i = BACKWARD ? j : j - d;
                    v -= tv.W[d - 1][(size_t)i * mD + kD] * w[d - 1];
                }
            }
            if (!tv.unit) v *= tv.rdD[(size_t)j * mD + kD];
            x[j * js] = v;
#pragma unroll
            for (int d = ND - 1; d >= 1; --d) w[d] = w[d - 1];
            if (ND > 0) w[0] = v;
        }
    }
}

// ------------------------------------------------------------------------------------
// hat phase, LEFT: one warp per column.
//   gather : H[i] -= sum_b sum_t S[b][t][k] * Z[b][k],   k = i-t+1        (:449-451)
//   scan   : bidiagonal substitution as a segmented affine scan over the 32 lanes  (:453 / :472)
//   scatter: Z[b][k] -= sum_t S[b][t][k] * H[k+t-1]                          (:475-477)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void hat_gather(const TriView &tv, double *H, const double *Z, int lane) {
    for (int i = lane; i < tv.nh; i += 32) {
        double v = H[i];
        for (int b = 0; b < tv.nb; ++b)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int k = i - t + 1;
                if (k >= 0 && k < tv.m) v -= tv.S[((size_t)b * 3 + t) * tv.m + k] * Z[(size_t)b * tv.m + k];
            }
        H[i] = v;
    }
}

__device__ __forceinline__ void hat_scatter(const TriView &tv, const double *H, double *Z, int lane) {
    for (int k = lane; k < tv.m; k += 32)
        for (int b = 0; b < tv.nb; ++b) {
            double v = Z[(size_t)b * tv.m + k];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int i = k + t - 1;
                if (i >= 0 && i < tv.nh) v -= tv.S[((size_t)b * 3 + t) * tv.m + k] * H[i];
            }
            Z[(size_t)b * tv.m + k] = v;
        }
}

// h[i] = (H[i] - offA[i] h[i+1]) * rd[i], i = nh-1..0  (BACKWARD)
// h[i] = (H[i] - offA[i-1] h[i-1]) * rd[i], i = 0..nh-1 (FORWARD)
// Lane l owns the contiguous segment [l*s, (l+1)*s).  Pass 1 reduces the segment to an
// affine map of the incoming value, a 5-step shuffle scan composes the maps, pass 2 is the
// reference recurrence itself with the exact incoming value.
template <bool BACKWARD>
__device__ __forceinline__ void hat_scan(const TriView &tv, double *H, int lane) {
    const int nh = tv.nh;
    const int s = (nh + 31) / 32;
    const int lo = min(lane * s, nh), hi = min(lo + s, nh);
    double t = 0.0, P = 1.0;
    for (int ii = lo; ii < hi; ++ii) {
        const int i = BACKWARD ? (hi - 1 - (ii - lo)) : ii;
        const bool has = BACKWARD ? (i < nh - 1) : (i > 0);
        const double off = has ? tv.offA[BACKWARD ? i : i - 1] : 0.0;
        const double rd = tv.unit ? 1.0 : tv.rdA[i];
        t = (H[i] - off * t) * rd;
        P = -off * rd * P;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t2 = BACKWARD ? __shfl_down_sync(FULL, t, o) : __shfl_up_sync(FULL, t, o);
        const double P2 = BACKWARD ? __shfl_down_sync(FULL, P, o) : __shfl_up_sync(FULL, P, o);
        const bool ok = BACKWARD ? (lane + o < 32) : (lane - o >= 0);
        if (ok) {
            t = t + P * t2;
            P = P * P2;
        }
    }
    double hin = BACKWARD ? __shfl_down_sync(FULL, t, 1) : __shfl_up_sync(FULL, t, 1);
    if (BACKWARD ? (lane == 31) : (lane == 0)) hin = 0.0;
    for (int ii = lo; ii < hi; ++ii) {
        const int i = BACKWARD ? (hi - 1 - (ii - lo)) : ii;
        const bool has = BACKWARD ? (i < nh - 1) : (i > 0);
        const double off = has ? tv.offA[BACKWARD ? i : i - 1] : 0.0;
        const double rd = tv.unit ? 1.0 : tv.rdA[i];
        hin = (H[i] - off * hin) * rd;
        H[i] = hin;
    }
}

// MODE 0: backward part (after the bubble backward sweep), 1: forward part (before the bubble
// forward sweep), 2: both (ReverseCholesky solve; `up` and `lo` describe U and U^T)
template <int MODE>
__global__ void __launch_bounds__(128) k_hat_left(TriView up, TriView lo, double *__restrict__ X, int64_t ld, int64_t nbatch) {
    const int lane = threadIdx.x;
    const int64_t c = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (c >= nbatch) return;
    double *H = X + c * ld;
    double *Z = H + up.nh;
    if (MODE == 0 || MODE == 2) {
        hat_gather(up, H, Z, lane);
        __syncwarp();
        hat_scan<true>(up, H, lane);
        __syncwarp();
    }
    if (MODE == 1 || MODE == 2) {
        hat_scan<false>(lo, H, lane);
        __syncwarp();
        hat_scatter(lo, H, Z, lane);
    }
}

// hat phase, RIGHT: one thread per batch member (row), everything sequential in the hat
// index, every access coalesced across the batch.
template <int MODE>
__global__ void __launch_bounds__(128) k_hat_right(TriView up, TriView lo, double *__restrict__ X, int64_t sv, int64_t nbatch) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nbatch) return;
    double *x = X + c;
    const int nh = up.nh, m = up.m;
    if (MODE == 0 || MODE == 2) {
        double hn = 0.0;
        for (int i = nh - 1; i >= 0; --i) {
            double v = x[(int64_t)i * sv];
            for (int b = 0; b < up.nb; ++b)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int k = i - t + 1;
                    if (k >= 0 && k < m) v -= up.S[((size_t)b * 3 + t) * m + k] * x[((int64_t)nh + (int64_t)b * m + k) * sv];
                }
            if (i < nh - 1) v -= up.offA[i] * hn;
            if (!up.unit) v *= up.rdA[i];
            x[(int64_t)i * sv] = v;
            hn = v;
        }
    }
    if (MODE == 1 || MODE == 2) {
        double hp = 0.0;
        for (int i = 0; i < nh; ++i) {
            double v = x[(int64_t)i * sv];
            if (i > 0) v -= lo.offA[i - 1] * hp;
            if (!lo.unit) v *= lo.rdA[i];
            x[(int64_t)i * sv] = v;
            hp = v;
        }
        for (int k = 0; k < m; ++k)
            for (int b = 0; b < lo.nb; ++b) {
                double v = x[((int64_t)nh + (int64_t)b * m + k) * sv];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = k + t - 1;
                    if (i >= 0 && i < nh) v -= lo.S[((size_t)b * 3 + t) * m + k] * x[(int64_t)i * sv];
                }
                x[((int64_t)nh + (int64_t)b * m + k) * sv] = v;
            }
    }
}

// ------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------
template <bool BACKWARD, bool LANE_BATCH>
static void launch_sweep(pop_ctx *ctx, const TriView &tv, double *X, int64_t sv, int64_t sc, int64_t nbatch) {
    if (tv.q == 0 || nbatch == 0) return;
    const int64_t nL = LANE_BATCH ? nbatch : tv.m;
    const int64_t nO = LANE_BATCH ? tv.m : nbatch;
    dim3 block(32, 8);
    dim3 grid((unsigned)((nL + 31) / 32), (unsigned)std::min<int64_t>((nO + 7) / 8, 32768));
    if (tv.nd == 0)
        k_bubble_sweep<BACKWARD, 0, LANE_BATCH><<<grid, block, 0, ctx->stream>>>(tv, X, sv, sc, nbatch);
    else if (tv.nd == 1)
        k_bubble_sweep<BACKWARD, 1, LANE_BATCH><<<grid, block, 0, ctx->stream>>>(tv, X, sv, sc, nbatch);
    else if (tv.nd == 2)
        k_bubble_sweep<BACKWARD, 2, LANE_BATCH><<<grid, block, 0, ctx->stream>>>(tv, X, sv, sc, nbatch);
    else
        k_bubble_sweep<BACKWARD, POP_MAXD, LANE_BATCH><<<grid, block, 0, ctx->stream>>>(tv, X, sv, sc, nbatch);
    ctx->launches++;
}

template <int MODE>
static void launch_hat(pop_ctx *ctx, const TriView &up, const TriView &lo, double *X, int64_t sv, int64_t sc, int64_t nbatch) {
    if (nbatch == 0) return;
    if (sv == 1) {
        dim3 block(32, 4);
        k_hat_left<MODE><<<(unsigned)((nbatch + 3) / 4), block, 0, ctx->stream>>>(up, lo, X, sc, nbatch);
    } else {
        k_hat_right<MODE><<<(unsigned)((nbatch + 127) / 128), 128, 0, ctx->stream>>>(up, lo, X, sv, nbatch);
    }
    ctx->launches++;
}

int pop_launch_trsm_staged(pop_ctx *ctx, const TriView &tv, bool backward, double *X, int64_t sv, int64_t sc, int64_t nbatch) {
    const bool lane_batch = sv != 1;
    POP_REQUIRE(sv == 1 || sc == 1, POP_ERR_ARG, "trsm: either the vector or the batch stride must be 1");
    if (backward) {
        if (lane_batch) launch_sweep<true, true>(ctx, tv, X, sv, sc, nbatch); else launch_sweep<true, false>(ctx, tv, X, sv, sc, nbatch);
        launch_hat<0>(ctx, tv, tv, X, sv, sc, nbatch);
    } else {
        launch_hat<1>(ctx, tv, tv, X, sv, sc, nbatch);
        if (lane_batch) launch_sweep<false, true>(ctx, tv, X, sv, sc, nbatch); else launch_sweep<false, false>(ctx, tv, X, sv, sc, nbatch);
    }
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

int pop_launch_solve_staged(pop_ctx *ctx, const TriView &up, const TriView &lo, double *X, int64_t sv, int64_t sc, int64_t nbatch) {
    const bool lane_batch = sv != 1;
    POP_REQUIRE(sv == 1 || sc == 1, POP_ERR_ARG, "solve: either the vector or the batch stride must be 1");
    if (lane_batch) launch_sweep<true, true>(ctx, up, X, sv, sc, nbatch); else launch_sweep<true, false>(ctx, up, X, sv, sc, nbatch);
    launch_hat<2>(ctx, up, lo, X, sv, sc, nbatch);
    if (lane_batch) launch_sweep<false, true>(ctx, lo, X, sv, sc, nbatch); else launch_sweep<false, false>(ctx, lo, X, sv, sc, nbatch);
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

// ------------------------------------------------------------------------------------
// ABI entry points
// ------------------------------------------------------------------------------------
static int check_panel(const pop_arrowhead *A, int side, const void *X, int64_t ldx, int64_t nrhs, const char *who) {
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "%s: bad side %d", who, side);
    POP_REQUIRE(nrhs >= 0, POP_ERR_DIM, "%s: negative nrhs", who);
    POP_REQUIRE(X != nullptr || nrhs == 0, POP_ERR_ARG, "%s: null matrix pointer", who);
    const int64_t rows = side == POP_LEFT ? A->n() : nrhs;
    POP_REQUIRE(ldx >= std::max<int64_t>(rows, 1), POP_ERR_DIM, "%s: DimensionMismatch: leading dimension %lld < %lld rows", who, (long long)ldx, (long long)rows);
    return POP_OK;
}

extern "C" int pop_trsm(pop_ctx *ctx, const pop_arrowhead *A, int side, int uplo, int trans, int unit, double *dX, int64_t ldx, int64_t nrhs) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_trsm: null argument");
    POP_REQUIRE(uplo == POP_UPPER || uplo == POP_LOWER, POP_ERR_ARG, "pop_trsm: bad uplo %d", uplo);
    int rc = check_panel(A, side, dX, ldx, nrhs, "pop_trsm");
    if (rc) return rc;
    if (nrhs == 0) return POP_OK;
    if (!unit) {
        rc = pop_ensure_recip(ctx, const_cast<pop_arrowhead *>(A));
        if (rc) return rc;
    }
    TriView tv;
    pop_make_triview(A, uplo, trans, unit, &tv);
    // X * op(T)^{-1} acts on rows with op(T)^T:   (X T^{-1})^T = T^{-T} X^T
    const bool eff_trans = (trans != 0) != (side == POP_RIGHT);
    const bool backward = (uplo == POP_UPPER) != eff_trans;
    const int64_t sv = side == POP_LEFT ? 1 : ldx, sc = side == POP_LEFT ? ldx : 1;
    return pop_launch_trsm_staged(ctx, tv, backward, dX, sv, sc, nrhs);
}

extern "C" int pop_solve(pop_ctx *ctx, const pop_arrowhead *U, int side, double *dX, int64_t ldx, int64_t nrhs, int algo) {
    POP_REQUIRE(ctx && U, POP_ERR_ARG, "pop_solve: null argument");
    POP_REQUIRE(U->is_factor, POP_ERR_ARG, "pop_solve: handle is not a reverse-Cholesky factor");
    int rc = check_panel(U, side, dX, ldx, nrhs, "pop_solve");
    if (rc) return rc;
    if (nrhs == 0) return POP_OK;
    if (side == POP_LEFT && algo != POP_SOLVE_STAGED) {
        if (algo != POP_SOLVE_FUSED_SMEM && pop_fused2_supported(ctx, U, nullptr, ldx, dX))
            return pop_launch_solve_fused2(ctx, U, nullptr, 0.0, nullptr, 0, dX, ldx, nrhs);
        if (pop_fused_supported(ctx, U, nullptr)) return pop_launch_solve_fused(ctx, U, nullptr, 0.0, nullptr, 0, dX, ldx, nrhs);
        POP_REQUIRE(algo != POP_SOLVE_FUSED && algo != POP_SOLVE_FUSED_SMEM, POP_ERR_DIM, "pop_solve: fused kernel does not support this shape (n=%lld)", (long long)U->n());
    }
    if (side == POP_RIGHT && (algo == POP_SOLVE_AUTO || algo == POP_SOLVE_FUSED)) {
        // X / F: the one-pass row kernel keeps a block of rows on chip for both sweeps (pop_row1.cuh)
        rc = pop_row1_solve_right(ctx, U, dX, ldx, nrhs);
        if (rc != POP_ROW1_UNSUPPORTED) return rc;
    }
    POP_REQUIRE(!(side == POP_RIGHT && (algo == POP_SOLVE_FUSED || algo == POP_SOLVE_FUSED_SMEM)), POP_ERR_ARG,
                "pop_solve: no single-pass kernel for this row panel (shared D block with <= 2 bands, q <= 64 needed)");
    TriView up, lo;
    pop_make_triview(U, POP_UPPER, 0, 0, &up);
    pop_make_triview(U, POP_UPPER, 1, 0, &lo);
    const int64_t sv = side == POP_LEFT ? 1 : ldx, sc = side == POP_LEFT ? ldx : 1;
    // S^{-1} is symmetric: X S^{-1} = (S^{-1} X^T)^T, same two sweeps on rows
    return pop_launch_solve_staged(ctx, up, lo, dX, sv, sc, nrhs);
}

extern "C" int pop_solve_shift_batch(pop_ctx *ctx, pop_arrowhead *const *U, int nshift, double *dX, int64_t ldx, int64_t stride_shift,
                                     int64_t nrhs, int algo) {
    POP_REQUIRE(ctx && U && nshift >= 0, POP_ERR_ARG, "pop_solve_shift_batch: bad argument");
    for (int i = 0; i < nshift; ++i) {
        int rc = pop_solve(ctx, U[i], POP_LEFT, dX + (int64_t)i * stride_shift, ldx, nrhs, algo);
        if (rc) return rc;
    }
    return POP_OK;
}

extern "C" int pop_solve_mul_axpy(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *Tm, double gamma, const double *dF, int64_t ldf,
                                  double *dR, int64_t ldr, int64_t nrhs) {
    POP_REQUIRE(ctx && U, POP_ERR_ARG, "pop_solve_mul_axpy: null argument");
    POP_REQUIRE(U->is_factor, POP_ERR_ARG, "pop_solve_mul_axpy: U is not a reverse-Cholesky factor");
    int rc = check_panel(U, POP_LEFT, dR, ldr, nrhs, "pop_solve_mul_axpy");
    if (rc) return rc;
    if (Tm) {
        POP_REQUIRE(Tm->d.m == U->d.m && Tm->d.nh == U->d.nh && Tm->d.q == U->d.q, POP_ERR_DIM, "pop_solve_mul_axpy: operator sizes differ");
        POP_REQUIRE(gamma == 0.0 || (dF != nullptr && ldf >= U->n()), POP_ERR_DIM, "pop_solve_mul_axpy: F panel missing or too small");
    }
    if (nrhs == 0) return POP_OK;
    if ((!Tm || gamma == 0.0 || (!(ldf & 1) && !((uintptr_t)dF & 15))) && pop_fused2_supported(ctx, U, Tm, ldr, dR))
        return pop_launch_solve_fused2(ctx, U, Tm, gamma, dF, ldf, dR, ldr, nrhs);
    // the shared-memory fused kernel moves columns by TMA bulk copies: 16-byte aligned column starts for R and F, or the
    // staged path below (never a half-updated panel and an error)
    const bool aligned = !(ldr & 1) && !((uintptr_t)dR & 15) && (!Tm || gamma == 0.0 || (!(ldf & 1) && !((uintptr_t)dF & 15)));
    if (aligned && pop_fused_supported(ctx, U, Tm)) return pop_launch_solve_fused(ctx, U, Tm, gamma, dF, ldf, dR, ldr, nrhs);
    // staged fallback: solve, then R <- T*X + gamma*F through a scratch copy
    TriView up, lo;
    pop_make_triview(U, POP_UPPER, 0, 0, &up);
    pop_make_triview(U, POP_UPPER, 1, 0, &lo);
    rc = pop_launch_solve_staged(ctx, up, lo, dR, 1, ldr, nrhs);
    if (rc || !Tm) return rc;
    const int64_t n = U->n();
    void *scr = nullptr;
    // the scratch also backs small batch tables; grab a private buffer
    cudaError_t e = cudaMallocAsync(&scr, (size_t)n * nrhs * sizeof(double), ctx->stream);
    if (e != cudaSuccess) { pop_set_error("pop_solve_mul_axpy: scratch of %lld bytes: %s", (long long)(n * nrhs * 8), cudaGetErrorString(e)); return POP_ERR_ALLOC; }
    POP_CUDA(cudaMemcpy2DAsync(scr, n * sizeof(double), dR, ldr * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToDevice, ctx->stream));
    if (gamma != 0.0)
        POP_CUDA(cudaMemcpy2DAsync(dR, ldr * sizeof(double), dF, ldf * sizeof(double), n * sizeof(double), nrhs, cudaMemcpyDeviceToDevice, ctx->stream));
    MulView mv;
    pop_make_mulview(Tm, POP_SYM_UPPER, 0, &mv);
    rc = pop_launch_mul(ctx, mv, 1.0, (const double *)scr, 1, n, gamma, dR, 1, ldr, nrhs);
    cudaFreeAsync(scr, ctx->stream);
    return rc;
}
