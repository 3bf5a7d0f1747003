// Right-hand-side pipeline of examples/adi_poisson.jl (the callers' side of the hot path):
//   * pop_legendre_analysis_2d: `analysis_2D` (examples/adi_poisson.jl:60-76): per mesh cell (i,j) the p x p block of
//     function values on the tensor grid becomes the block of tensor Legendre coefficients  C = T V T^T  (T = the p x p
//     analysis matrix of `plan_grid_transform(legendre(..), (p,p))`), stored interlaced like every block vector here:
//     F[i + n a, j + n b].  One CTA per cell, both small products on FP64 FMAs out of shared memory (FP64 has no
//     tcgen05 kind; at p = 64 the kernel is balanced between the FMA pipe and HBM).
//   * pop_weakform_apply: the arrowhead-shaped rectangular operator (Q'P)[KR,KR] = (P\Q)' (P'P)
//     (examples/adi_poisson.jl:105; P\C: src/continuouspolynomial.jl:226-235, C\Q: src/dirichlet.jl:35-42, P'P:
//     src/continuouspolynomial.jl:259-272) applied from the left (columns) or from the right (rows): a two/four-point
//     stencil per output entry, HBM-bound at 16 B per entry.
#include <algorithm>

#include "pop_internal.h"

// ---------------------------------------------------------------------------------------------------
// weak form:  y = (C'P) x ,  x in P ordering (degree a of element k at a*m + k), y in C1 / Dirichlet ordering
//   hat on node nd:  1/2 [ g(0,nd-1) x(0,nd-1) + g(1,nd-1) x(1,nd-1) ] + 1/2 [ g(0,nd) x(0,nd) - g(1,nd) x(1,nd) ]
//   bubble (j,k), j = 1..q (W_{j+1} = (P_{j-1} - P_{j+1})/(2j+1)):  v_j [ g(j-1,k) x(j-1,k) - g(j+1,k) x(j+1,k) ]
//   g(a,k) = h_k/(2a+1) = <P_a,P_a> on element k,  v_j = 1/(2j+1)
// ---------------------------------------------------------------------------------------------------
struct WeakOp {
    int basis, m, N, nh, q;
    double h;             // uniform step (hk == nullptr)
    const double *hk;     // [m] element lengths or null
};

__device__ __forceinline__ double wf_h(const WeakOp &w, int k) { return w.hk ? w.hk[k] : w.h; }

// sv: stride between consecutive entries of a vector, sc: stride between vectors (LEFT: 1, ld; RIGHT: ld, 1)
// lanes run over `fast` (LEFT: the output row; RIGHT: the batch index), so every access is a run of consecutive doubles
template <bool LEFT>
__global__ void __launch_bounds__(256) k_weakform(WeakOp w, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int64_t cnt) {
    const int64_t n = (int64_t)w.nh + (int64_t)w.m * w.q;
    const int m = w.m, N = w.N;
    // LEFT: x = output row (fast), y = vector ; RIGHT: x = vector (fast, a row of the panels), y = output column
    const int64_t fx = (int64_t)blockIdx.x * 32 + threadIdx.x;
    for (int64_t sy = (int64_t)blockIdx.y * 8 + threadIdx.y; sy < (LEFT ? cnt : n); sy += (int64_t)gridDim.y * 8) {
        const int64_t r = LEFT ? fx : sy, c = LEFT ? sy : fx;
        if (r >= n || c >= cnt) continue;
        const double *x = LEFT ? X + c * ldx : X + c;
        const int64_t xs = LEFT ? 1 : ldx;
        double acc = 0.0;
        if (r < w.nh) {
            const int nd = w.basis == POP_BASIS_C1 ? (int)r : (int)r + 1;
            if (nd >= 1) {
                const int k = nd - 1;
                const double h = wf_h(w, k);
                acc += 0.5 * h * x[(int64_t)k * xs];
                if (N >= 2) acc += 0.5 * (h / 3) * x[((int64_t)m + k) * xs];
            }
            if (nd < m) {
                const int k = nd;
                const double h = wf_h(w, k);
                acc += 0.5 * h * x[(int64_t)k * xs];
                if (N >= 2) acc -= 0.5 * (h / 3) * x[((int64_t)m + k) * xs];
            }
        } else {
            const int64_t b = r - w.nh;
            const int j = (int)(b / m) + 1, k = (int)(b % m);
            const double h = wf_h(w, k), v = 1.0 / (2 * j + 1);
            acc = (h / (2 * j - 1)) * x[((int64_t)(j - 1) * m + k) * xs];
            if (j + 1 <= N - 1) acc -= (h / (2 * j + 3)) * x[((int64_t)(j + 1) * m + k) * xs];
            acc *= v;
        }
        if (LEFT) Y[c * ldy + r] = acc;
        else Y[r * ldy + c] = acc;
    }
}

extern "C" int pop_weakform_apply(pop_ctx *ctx, int basis, int m, int N, double h, const double *h_elems, int side, const double *dX, int64_t ldx,
                                  double *dY, int64_t ldy, int64_t cnt) {
    POP_REQUIRE(ctx && dX && dY, POP_ERR_ARG, "pop_weakform_apply: null argument");
    POP_REQUIRE(basis == POP_BASIS_C1 || basis == POP_BASIS_DIRICHLET, POP_ERR_ARG, "pop_weakform_apply: unknown basis %d", basis);
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "pop_weakform_apply: bad side");
    const int nh = basis == POP_BASIS_C1 ? m + 1 : m - 1;
    POP_REQUIRE(m >= 1 && nh >= 1 && N >= 1, POP_ERR_DIM, "pop_weakform_apply: need m >= 1 elements with hats and N >= 1 blocks");
    POP_REQUIRE(h_elems || h > 0, POP_ERR_ARG, "pop_weakform_apply: need a positive step or element lengths");
    const int64_t n = (int64_t)nh + (int64_t)m * (N - 1), np = (int64_t)m * N;
    POP_REQUIRE(cnt >= 0, POP_ERR_DIM, "pop_weakform_apply: negative count");
    if (side == POP_LEFT) POP_REQUIRE(ldx >= np && ldy >= n, POP_ERR_DIM, "pop_weakform_apply: DimensionMismatch: X needs %lld rows, Y %lld", (long long)np, (long long)n);
    else POP_REQUIRE(ldx >= std::max<int64_t>(cnt, 1) && ldy >= std::max<int64_t>(cnt, 1), POP_ERR_DIM, "pop_weakform_apply: leading dimensions smaller than the row count");
    if (cnt == 0) return POP_OK;
    WeakOp w = {basis, m, N, nh, N - 1, h, nullptr};
    if (h_elems) {
        for (int k = 0; k < m; ++k) POP_REQUIRE(h_elems[k] > 0, POP_ERR_ARG, "pop_weakform_apply: element %d has non-positive length", k);
        double *dh = nullptr;
        int rc = pop_ctx_scratch(ctx, (size_t)m * sizeof(double), (void **)&dh);
        if (rc) return rc;
        POP_CUDA(cudaMemcpyAsync(dh, h_elems, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        w.hk = dh;
    }
    const dim3 blk(32, 8);
    if (side == POP_LEFT) {
        const dim3 grd((unsigned)((n + 31) / 32), (unsigned)std::min<int64_t>((cnt + 7) / 8, 16384));
        k_weakform<true><<<grd, blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    } else {
        const dim3 grd((unsigned)((cnt + 31) / 32), (unsigned)std::min<int64_t>((n + 7) / 8, 16384));
        k_weakform<false><<<grd, blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    }
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    if (h_elems) POP_CUDA(cudaStreamSynchronize(ctx->stream));   // the scratch copy of the element lengths may be reused at once
    return POP_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-cell tensor Legendre analysis  C = T V T^T
// ---------------------------------------------------------------------------------------------------
#define AN_TILE 4
// Shared memory (pp = p rounded up to 4, rows padded by 4 doubles against bank conflicts of the transposing stores):
//   Tt[s][a] = T[a][s] ; Vt[s][t] = V[s][t] of the cell (t fastest) ; Wt[t][a] = (T V)[a][t]
// 256 threads, each a 4x4 tile of a 64x64 product; two passes of p rank-1 updates from 128-bit shared-memory loads.
__global__ void __launch_bounds__(256, 2) k_analysis_2d(int n, int p, const double *__restrict__ T, const double *__restrict__ V, int64_t ldv,
                                                        double *__restrict__ F, int64_t ldf) {
    extern __shared__ __align__(16) double an_sm[];
    const int pp = (p + 3) & ~3, ldp = pp + 4;
    double *Tt = an_sm, *Vt = Tt + (size_t)pp * ldp, *Wt = Vt + (size_t)pp * ldp;
    const int tid = threadIdx.x;
    const int ci = blockIdx.x, cj = blockIdx.y;
    // T is p x p column-major (T[a + p s]); transposed on the way in so that 4 consecutive a share a 32 B run
    for (int e = tid; e < pp * pp; e += 256) {
        const int a = e % pp, s = e / pp;
        Tt[s * ldp + a] = (a < p && s < p) ? T[a + (size_t)p * s] : 0.0;
    }
    // values of the cell: column t of the cell is p consecutive doubles (node s fastest)
    const double *Vc = V + (int64_t)ci * p + (int64_t)cj * p * ldv;
    for (int e = tid; e < pp * pp; e += 256) {
        const int s = e % pp, t = e / pp;
        Vt[s * ldp + t] = (s < p && t < p) ? Vc[s + (int64_t)t * ldv] : 0.0;
    }
    __syncthreads();
    const int nt = pp / AN_TILE;            // tiles per dimension (<= 16)
    const int ta = tid % 16, tb = tid / 16;
    const bool on = ta < nt && tb < nt;
    const int a0 = ta * AN_TILE, t0 = tb * AN_TILE;
    double acc[AN_TILE][AN_TILE];
#pragma unroll
    for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
        for (int y = 0; y < AN_TILE; ++y) acc[x][y] = 0.0;
    if (on) {
        // W[a][t] = sum_s T[a][s] V[s][t]
        for (int s = 0; s < p; ++s) {
            const double2 ta01 = *reinterpret_cast<const double2 *>(Tt + s * ldp + a0), ta23 = *reinterpret_cast<const double2 *>(Tt + s * ldp + a0 + 2);
            const double2 vb01 = *reinterpret_cast<const double2 *>(Vt + s * ldp + t0), vb23 = *reinterpret_cast<const double2 *>(Vt + s * ldp + t0 + 2);
            const double av[4] = {ta01.x, ta01.y, ta23.x, ta23.y}, bv[4] = {vb01.x, vb01.y, vb23.x, vb23.y};
#pragma unroll
            for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
                for (int y = 0; y < AN_TILE; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
        }
#pragma unroll
        for (int y = 0; y < AN_TILE; ++y) {
            *reinterpret_cast<double2 *>(Wt + (t0 + y) * ldp + a0) = make_double2(acc[0][y], acc[1][y]);
            *reinterpret_cast<double2 *>(Wt + (t0 + y) * ldp + a0 + 2) = make_double2(acc[2][y], acc[3][y]);
        }
    }
    __syncthreads();
    if (on) {
        // C[a][b] = sum_t W[a][t] T[b][t]
        const int b0 = t0;
#pragma unroll
        for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
            for (int y = 0; y < AN_TILE; ++y) acc[x][y] = 0.0;
        for (int t = 0; t < p; ++t) {
            const double2 wa01 = *reinterpret_cast<const double2 *>(Wt + t * ldp + a0), wa23 = *reinterpret_cast<const double2 *>(Wt + t * ldp + a0 + 2);
            const double2 tb01 = *reinterpret_cast<const double2 *>(Tt + t * ldp + b0), tb23 = *reinterpret_cast<const double2 *>(Tt + t * ldp + b0 + 2);
            const double av[4] = {wa01.x, wa01.y, wa23.x, wa23.y}, bv[4] = {tb01.x, tb01.y, tb23.x, tb23.y};
#pragma unroll
            for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
                for (int y = 0; y < AN_TILE; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
        }
        // interlaced store: degree a of cell ci at row ci + n a (neighbouring cells in x, i.e. neighbouring CTAs, complete the sector)
#pragma unroll
        for (int y = 0; y < AN_TILE; ++y)
#pragma unroll
            for (int x = 0; x < AN_TILE; ++x) {
                const int a = a0 + x, b = b0 + y;
                if (a < p && b < p) F[(int64_t)ci + (int64_t)n * a + ((int64_t)cj + (int64_t)n * b) * ldf] = acc[x][y];
            }
    }
}

extern "C" int pop_legendre_analysis_2d(pop_ctx *ctx, int n, int p, const double *T_host, const double *dV, int64_t ldv, double *dF, int64_t ldf) {
    POP_REQUIRE(ctx && T_host && dV && dF, POP_ERR_ARG, "pop_legendre_analysis_2d: null argument");
    POP_REQUIRE(n >= 1 && p >= 1, POP_ERR_ARG, "pop_legendre_analysis_2d: need n >= 1 cells per dimension and p >= 1 nodes per cell");
    POP_REQUIRE(p <= 64, POP_ERR_DIM, "pop_legendre_analysis_2d: p = %d nodes per cell: the per-cell kernel holds p <= 64", p);
    const int64_t np = (int64_t)n * p;
    POP_REQUIRE(ldv >= np && ldf >= np, POP_ERR_DIM, "pop_legendre_analysis_2d: DimensionMismatch: panels need %lld rows", (long long)np);
    POP_REQUIRE(n <= 65535, POP_ERR_DIM, "pop_legendre_analysis_2d: at most 65535 cells per dimension");
    double *dT = nullptr;
    int rc = pop_ctx_scratch(ctx, (size_t)p * p * sizeof(double), (void **)&dT);
    if (rc) return rc;
    POP_CUDA(cudaMemcpyAsync(dT, T_host, (size_t)p * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int pp = (p + 3) & ~3;
    const size_t smem = (size_t)3 * pp * (pp + 4) * sizeof(double);
    POP_CUDA(cudaFuncSetAttribute(k_analysis_2d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_analysis_2d<<<dim3((unsigned)n, (unsigned)n), 256, smem, ctx->stream>>>(n, p, dT, dV, ldv, dF, ldf);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    POP_CUDA(cudaStreamSynchronize(ctx->stream));   // T_host and the scratch copy may be reused at once
    return POP_OK;
}
