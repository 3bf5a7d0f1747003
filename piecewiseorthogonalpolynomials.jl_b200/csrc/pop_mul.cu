// Products with packed arrowheads and the transposition kernel.
//   y <- alpha*A*x + beta*y           src/arrowhead.jl:191-212
//   Y <- alpha*A*X + beta*Y           src/arrowhead.jl:214-235
//   Y <- alpha*X*A + beta*Y           src/arrowhead.jl:237-258
// One thread per output entry in gather form; lanes run over the contiguous index (vector
// index for LEFT, batch index for RIGHT) so all accesses coalesce.  Each X entry is re-read
// by the (lD+uD+1) rows that touch it; those re-reads are 2m doubles apart and hit L1/L2.
#include "pop_internal.h"

template <bool LANE_BATCH>
__global__ void __launch_bounds__(256) k_mul(MulView mv, double alpha, const double *__restrict__ X, int64_t svx, int64_t scx, double beta,
                                             double *__restrict__ Y, int64_t svy, int64_t scy, int64_t n, int64_t nbatch) {
    const int64_t nL = LANE_BATCH ? nbatch : n;
    const int64_t nO = LANE_BATCH ? n : nbatch;
    const int64_t l = (int64_t)blockIdx.x * 32 + threadIdx.x;
    if (l >= nL) return;
    const int m = mv.m, nh = mv.nh, q = mv.q, mD = mv.mD;
    for (int64_t o = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; o < nO; o += (int64_t)gridDim.y * blockDim.y) {
        const int64_t r = LANE_BATCH ? o : l;
        const int64_t c = LANE_BATCH ? l : o;
        const double *x = X + c * scx;
        double acc;
        if (r < nh) {
            const int i = (int)r;
            acc = mv.Ad[i] * x[(int64_t)i * svx];
            if (i + 1 < nh) acc += mv.Aup[i] * x[(int64_t)(i + 1) * svx];
            if (i > 0) acc += mv.Alo[i - 1] * x[(int64_t)(i - 1) * svx];
            for (int b = 0; b < mv.nb_row; ++b)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int k = i - t + 1;
                    if (k >= 0 && k < m) acc += mv.Srow[((size_t)b * 3 + t) * m + k] * x[((int64_t)nh + (int64_t)b * m + k) * svx];
                }
        } else {
            const int64_t idx = r - nh;
            const int j = (int)(idx / m), k = (int)(idx % m);
            const int kD = mD == 1 ? 0 : k;
            acc = 0.0;
            for (int d = -mv.lD; d <= mv.uD; ++d) {
                const int jj = j + d;
                if (jj >= 0 && jj < q) acc += mv.Dp[d + POP_MAXD][(ptrdiff_t)j * mD + kD] * x[((int64_t)nh + (int64_t)jj * m + k) * svx];
            }
            if (j < mv.nb_col) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = k + t - 1;
                    if (i >= 0 && i < nh) acc += mv.Scol[((size_t)j * 3 + t) * m + k] * x[(int64_t)i * svx];
                }
            }
        }
        double *y = Y + c * scy + r * svy;
        *y = beta == 0.0 ? alpha * acc : alpha * acc + beta * *y;
    }
}

int pop_launch_mul(pop_ctx *ctx, const MulView &mv, double alpha, const double *X, int64_t svx, int64_t scx, double beta, double *Y,
                   int64_t svy, int64_t scy, int64_t nbatch) {
    const int64_t n = (int64_t)mv.nh + (int64_t)mv.m * mv.q;
    if (nbatch == 0 || n == 0) return POP_OK;
    const bool lane_batch = svx != 1;
    const int64_t nL = lane_batch ? nbatch : n, nO = lane_batch ? n : nbatch;
    dim3 block(32, 8);
    dim3 grid((unsigned)((nL + 31) / 32), (unsigned)std::min<int64_t>((nO + 7) / 8, 32768));
    if (lane_batch)
        k_mul<true><<<grid, block, 0, ctx->stream>>>(mv, alpha, X, svx, scx, beta, Y, svy, scy, n, nbatch);
    else
        k_mul<false><<<grid, block, 0, ctx->stream>>>(mv, alpha, X, svx, scx, beta, Y, svy, scy, n, nbatch);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

extern "C" int pop_mul(pop_ctx *ctx, const pop_arrowhead *A, int side, int sym, int trans, double alpha, const double *dX, int64_t ldx,
                       double beta, double *dY, int64_t ldy, int64_t nrhs) {
    POP_REQUIRE(ctx && A, POP_ERR_ARG, "pop_mul: null argument");
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "pop_mul: bad side %d", side);
    POP_REQUIRE(nrhs >= 0, POP_ERR_DIM, "pop_mul: negative nrhs");
    if (nrhs == 0) return POP_OK;
    POP_REQUIRE(dX && dY, POP_ERR_ARG, "pop_mul: null matrix pointer");
    POP_REQUIRE(dX != dY, POP_ERR_ARG, "pop_mul: X and Y must not alias");
    const int64_t rows = side == POP_LEFT ? A->n() : nrhs;
    POP_REQUIRE(ldx >= rows && ldy >= rows, POP_ERR_DIM, "pop_mul: DimensionMismatch: leading dimensions (%lld,%lld) < %lld rows", (long long)ldx,
                (long long)ldy, (long long)rows);
    MulView mv;
    // X*A acts on rows with A^T
    const int eff_trans = (trans != 0) != (side == POP_RIGHT);
    pop_make_mulview(A, sym, eff_trans, &mv);
    if (side == POP_LEFT) return pop_launch_mul(ctx, mv, alpha, dX, 1, ldx, beta, dY, 1, ldy, nrhs);
    return pop_launch_mul(ctx, mv, alpha, dX, ldx, 1, beta, dY, ldy, 1, nrhs);
}

// ------------------------------------------------------------------------------------
// tiled transpose through shared memory (32x33 tiles), dst = src^T
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transpose(const double *__restrict__ src, int64_t lds, int64_t rows, int64_t cols, double *__restrict__ dst,
                                                   int64_t ldd) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += 8) {
        const int64_t r = r0 + threadIdx.x, c = c0 + cc;
        if (r < rows && c < cols) tile[cc][threadIdx.x] = src[c * lds + r];
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const int64_t c = c0 + threadIdx.x, r = r0 + rr;
        if (r < rows && c < cols) dst[r * ldd + c] = tile[threadIdx.x][rr];
    }
}

extern "C" int pop_transpose(pop_ctx *ctx, const double *dsrc, int64_t lds, int64_t rows, int64_t cols, double *ddst, int64_t ldd) {
    POP_REQUIRE(ctx && dsrc && ddst, POP_ERR_ARG, "pop_transpose: null argument");
    POP_REQUIRE(rows >= 0 && cols >= 0 && lds >= rows && ldd >= cols, POP_ERR_DIM, "pop_transpose: bad dimensions");
    if (rows == 0 || cols == 0) return POP_OK;
    POP_REQUIRE((cols + 31) / 32 <= 65535, POP_ERR_DIM, "pop_transpose: too many columns");
    dim3 block(32, 8), grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    k_transpose<<<grid, block, 0, ctx->stream>>>(dsrc, lds, rows, cols, ddst, ldd);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}
