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

// One thread per (element k, vector c) walks the degrees of its element with a window of three inputs, so every input is
// read once and every output written once (16 B per entry).  LEFT: lanes over k (entries of a column are contiguous in
// k), RIGHT: lanes over c (the vectors are rows of the panels, contiguous in c).  The hat on the left node of element k
// (and, for the last element, on its right node) is produced by the same thread; it re-reads two inputs of element k-1.
template <bool LEFT>
__global__ void __launch_bounds__(256) k_weakform(WeakOp w, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int64_t cnt) {
    const int m = w.m, N = w.N, q = w.q;
    const int64_t ix = (int64_t)blockIdx.x * 32 + threadIdx.x, iy = (int64_t)blockIdx.y * 8 + threadIdx.y;
    const int64_t kk = LEFT ? ix : iy, c = LEFT ? iy : ix;
    if (kk >= m || c >= cnt) return;
    const int k = (int)kk;
    const double *x = LEFT ? X + c * ldx : X + c;
    double *y = LEFT ? Y + c * ldy : Y + c;
    const int64_t xs = LEFT ? 1 : ldx, ys = LEFT ? 1 : ldy;
    const double h = wf_h(w, k);
    const double x0 = x[(int64_t)k * xs];
    const double x1 = N >= 2 ? x[((int64_t)m + k) * xs] : 0.0;
    // hats: node nd carries hat nd (C1) or nd-1 (Dirichlet, interior nodes only)
    const int hoff = w.basis == POP_BASIS_C1 ? 0 : -1;
    {
        const int nd = k, hi = nd + hoff;
        if (hi >= 0 && hi < w.nh) {
            double acc = 0.5 * h * x0 - 0.5 * (h / 3) * x1;               // left node of element k: (P_0 - P_1)/2
            if (k >= 1) {
                const double hl = wf_h(w, k - 1);
                acc += 0.5 * hl * x[(int64_t)(k - 1) * xs];               // right node of element k-1: (P_0 + P_1)/2
                if (N >= 2) acc += 0.5 * (hl / 3) * x[((int64_t)m + k - 1) * xs];
            }
            y[(int64_t)hi * ys] = acc;
        }
    }
    if (k == m - 1) {
        const int hi = m + hoff;
        if (hi >= 0 && hi < w.nh) y[(int64_t)hi * ys] = 0.5 * h * x0 + 0.5 * (h / 3) * x1;
    }
    // bubbles: y(j,k) = v_j [ g(j-1,k) x(j-1,k) - g(j+1,k) x(j+1,k) ],  g(a,k) = h/(2a+1),  v_j = 1/(2j+1)
    double xa = x0, xb = x1;
    const double *xp = x + ((int64_t)2 * m + k) * xs;
    double *yp = y + ((int64_t)w.nh + k) * ys;
    const int64_t xstep = (int64_t)m * xs, ystep = (int64_t)m * ys;
    int j = 1;
    for (; j + 4 <= q && j + 4 <= N - 2; j += 4) {       // four degrees per trip: the loads are issued together
        double xn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xn[u] = xp[u * xstep];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jj = j + u;
            const double v = 1.0 / (2 * jj + 1);
            yp[u * ystep] = v * ((h / (2 * jj - 1)) * xa - (h / (2 * jj + 3)) * xn[u]);
            xa = xb;
            xb = xn[u];
        }
        xp += 4 * xstep;
        yp += 4 * ystep;
    }
    for (; j <= q; ++j) {
        const double xc = j + 1 <= N - 1 ? *xp : 0.0;
        const double v = 1.0 / (2 * j + 1);
        *yp = v * ((h / (2 * j - 1)) * xa - (h / (2 * j + 3)) * xc);
        xa = xb;
        xb = xc;
        xp += xstep;
        yp += ystep;
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
        POP_REQUIRE((cnt + 7) / 8 <= 65535, POP_ERR_DIM, "pop_weakform_apply: at most 524280 columns per call");
        const dim3 grd((unsigned)((m + 31) / 32), (unsigned)((cnt + 7) / 8));
        k_weakform<true><<<grd, blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    } else {
        POP_REQUIRE((m + 7) / 8 <= 65535, POP_ERR_DIM, "pop_weakform_apply: at most 524280 elements");
        const dim3 grd((unsigned)((cnt + 31) / 32), (unsigned)((m + 7) / 8));
        k_weakform<false><<<grd, blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    }
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    if (h_elems) POP_CUDA(cudaStreamSynchronize(ctx->stream));   // the scratch copy of the element lengths may be reused at once
    return POP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Legendre coefficients of a continuous piecewise polynomial -> its C1 / Dirichlet coefficients: the `R \ dat`
// + hat / bubble interlacing of ContinuousPolynomialTransform (src/continuouspolynomial.jl:84-146) and
// adaptivetransform_ldiv (:186-213); DirichletPolynomial drops the end hats (src/dirichlet.jl:44-53,65-69).
// Per element, from the top degree down (W_{j+1} = (P_{j-1} - P_{j+1})/(2j+1) is the only bubble with degree j+1):
//   b_j = (2a-1) ( b_{j+2}/(2a+3) - y_a ),  a = j+1 = N-1 .. 2 ;   t0 = y_0 - b_1/3,  t1 = y_1 - b_2/5
//   hat on the left node = t0 - t1, on the right node = t0 + t1
// Node 0 takes the left value of element 0, node k+1 the right value of element k (:103-111); the left values go to
// `ul` so that the caller can check continuity like the reference does (:105-107).
// ---------------------------------------------------------------------------------------------------
template <bool LEFT>
__global__ void __launch_bounds__(256) k_legendre_to_c1(WeakOp w, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int64_t cnt,
                                                        double *__restrict__ ul) {
    const int m = w.m, N = w.N, q = w.q;   // q = N - 2 bubbles per element
    const int64_t ix = (int64_t)blockIdx.x * 32 + threadIdx.x, iy = (int64_t)blockIdx.y * 8 + threadIdx.y;
    const int64_t kk = LEFT ? ix : iy, c = LEFT ? iy : ix;
    if (kk >= m || c >= cnt) return;
    const int k = (int)kk;
    const double *x = LEFT ? X + c * ldx : X + c;
    double *y = LEFT ? Y + c * ldy : Y + c;
    const int64_t xs = LEFT ? 1 : ldx, ys = LEFT ? 1 : ldy;
    const int64_t xstep = (int64_t)m * xs, ystep = (int64_t)m * ys;
    double b1 = 0.0, b2 = 0.0;                   // b_{j+1}, b_{j+2}
    const double *xp = x + ((int64_t)(N - 1) * m + k) * xs;
    double *yp = y + ((int64_t)w.nh + (int64_t)(q - 1) * m + k) * ys;
    int a = N - 1;
    for (; a - 3 >= 2; a -= 4) {                 // four degrees per trip: the loads are issued together
        double yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) yv[u] = xp[-u * xstep];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int aa = a - u;
            const double bj = (2 * aa - 1) * (b2 / (2 * aa + 3) - yv[u]);
            yp[-u * ystep] = bj;
            b2 = b1;
            b1 = bj;
        }
        xp -= 4 * xstep;
        yp -= 4 * ystep;
    }
    for (; a >= 2; --a) {
        const double bj = (2 * a - 1) * (b2 / (2 * a + 3) - *xp);
        *yp = bj;
        b2 = b1;
        b1 = bj;
        xp -= xstep;
        yp -= ystep;
    }
    // now b1 = b_1 (if q >= 1), b2 = b_2 (if q >= 2)
    const double t0 = x[(int64_t)k * xs] - b1 / 3, t1 = x[((int64_t)m + k) * xs] - b2 / 5;
    const double uL = t0 - t1, uR = t0 + t1;
    const int hoff = w.basis == POP_BASIS_C1 ? 0 : -1;
    if (k == 0 && hoff == 0) y[0] = uL;
    const int hi = k + 1 + hoff;
    if (hi >= 0 && hi < w.nh) y[(int64_t)hi * ys] = uR;
    if (ul) ul[c * (int64_t)(m + 1) + k] = uL;
    if (ul && k == m - 1) ul[c * (int64_t)(m + 1) + m] = uR;     // the right end (a Dirichlet basis expects zero there)
}

// largest |right value of element k-1 - left value of element k| (and, for the Dirichlet basis, |end values|)
__global__ void k_c1_jump(WeakOp w, const double *__restrict__ Y, int64_t ldy, int64_t cnt, bool left, const double *__restrict__ ul,
                          unsigned long long *__restrict__ out) {
    const int m = w.m;
    double worst = 0.0;
    const int64_t total = cnt * (int64_t)(m + 1);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e / (m + 1);
        const int nd = (int)(e % (m + 1));
        const double v = ul[e];
        const int hi = nd + (w.basis == POP_BASIS_C1 ? 0 : -1);
        double d;
        if (hi >= 0 && hi < w.nh && nd >= 1 && nd < m) d = fabs((left ? Y[c * ldy + hi] : Y[(int64_t)hi * ldy + c]) - v);
        else d = (w.basis == POP_BASIS_C1) ? 0.0 : fabs(v);
        worst = fmax(worst, d);
    }
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like integers
}

extern "C" int pop_legendre_to_c1(pop_ctx *ctx, int basis, int m, int N, int side, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t cnt,
                                  double *max_jump) {
    POP_REQUIRE(ctx && dX && dY, POP_ERR_ARG, "pop_legendre_to_c1: null argument");
    POP_REQUIRE(basis == POP_BASIS_C1 || basis == POP_BASIS_DIRICHLET, POP_ERR_ARG, "pop_legendre_to_c1: unknown basis %d", basis);
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "pop_legendre_to_c1: bad side");
    const int nh = basis == POP_BASIS_C1 ? m + 1 : m - 1;
    POP_REQUIRE(m >= 1 && nh >= 1 && N >= 2, POP_ERR_DIM, "pop_legendre_to_c1: need m >= 1 elements with hats and N >= 2 Legendre degrees per element");
    const int64_t n = (int64_t)nh + (int64_t)m * (N - 2), np = (int64_t)m * N;
    POP_REQUIRE(cnt >= 0, POP_ERR_DIM, "pop_legendre_to_c1: negative count");
    if (side == POP_LEFT) POP_REQUIRE(ldx >= np && ldy >= n, POP_ERR_DIM, "pop_legendre_to_c1: DimensionMismatch: X needs %lld rows, Y %lld", (long long)np, (long long)n);
    else POP_REQUIRE(ldx >= std::max<int64_t>(cnt, 1) && ldy >= std::max<int64_t>(cnt, 1), POP_ERR_DIM, "pop_legendre_to_c1: leading dimensions smaller than the row count");
    if (max_jump) *max_jump = 0.0;
    if (cnt == 0) return POP_OK;
    WeakOp w = {basis, m, N, nh, N - 2, 1.0, nullptr};
    double *ul = nullptr;
    unsigned long long *dmax = nullptr;
    if (max_jump) {
        cudaError_t e = cudaMallocAsync((void **)&ul, ((size_t)cnt * (m + 1) + 1) * sizeof(double), ctx->stream);
        if (e != cudaSuccess) { pop_set_error("pop_legendre_to_c1: scratch: %s", cudaGetErrorString(e)); return POP_ERR_ALLOC; }
        dmax = (unsigned long long *)(ul + (size_t)cnt * (m + 1));
        POP_CUDA(cudaMemsetAsync(dmax, 0, sizeof(double), ctx->stream));
    }
    const dim3 blk(32, 8);
    if (side == POP_LEFT) {
        POP_REQUIRE((cnt + 7) / 8 <= 65535, POP_ERR_DIM, "pop_legendre_to_c1: at most 524280 columns per call");
        k_legendre_to_c1<true><<<dim3((unsigned)((m + 31) / 32), (unsigned)((cnt + 7) / 8)), blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt, ul);
    } else {
        POP_REQUIRE((m + 7) / 8 <= 65535, POP_ERR_DIM, "pop_legendre_to_c1: at most 524280 elements");
        k_legendre_to_c1<false><<<dim3((unsigned)((cnt + 31) / 32), (unsigned)((m + 7) / 8)), blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt, ul);
    }
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    if (max_jump) {
        const int64_t total = cnt * (int64_t)(m + 1);
        k_c1_jump<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1184), 256, 0, ctx->stream>>>(w, dY, ldy, cnt, side == POP_LEFT, ul, dmax);
        ctx->launches++;
        POP_CUDA(cudaGetLastError());
        POP_CUDA(cudaMemcpyAsync(max_jump, dmax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        POP_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFreeAsync(ul, ctx->stream);
    }
    return POP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Coefficients in ContinuousPolynomial{1} / DirichletPolynomial (blocks 1..N-1) -> Legendre coefficients per element
// (N degrees, P ordering): the `Pl.R \ dat` of the INVERSE transform `Pl \ cfs` (src/continuouspolynomial.jl:129-146,
// :171-184; src/dirichlet.jl:97-110,143-163), i.e. the truncated conversion (P \ C)[Block.(1:N), Block.(1:N-1)] of
// src/continuouspolynomial.jl:226-235:  hat on the left node = (P_0 - P_1)/2, on the right node = (P_0 + P_1)/2,
// bubble j = W_{j+1} = (P_{j-1} - P_{j+1}) / (2j+1).  One thread per (element, vector), every entry read / written once.
// ---------------------------------------------------------------------------------------------------
template <bool LEFT>
__global__ void __launch_bounds__(256) k_c1_to_legendre(WeakOp w, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy, int64_t cnt) {
    const int m = w.m, N = w.N, q = w.q;   // q = N - 2 bubbles per element
    const int64_t ix = (int64_t)blockIdx.x * 32 + threadIdx.x, iy = (int64_t)blockIdx.y * 8 + threadIdx.y;
    const int64_t kk = LEFT ? ix : iy, c = LEFT ? iy : ix;
    if (kk >= m || c >= cnt) return;
    const int k = (int)kk;
    const double *x = LEFT ? X + c * ldx : X + c;
    double *y = LEFT ? Y + c * ldy : Y + c;
    const int64_t xs = LEFT ? 1 : ldx, ys = LEFT ? 1 : ldy;
    const int64_t xstep = (int64_t)m * xs, ystep = (int64_t)m * ys;
    const int hoff = w.basis == POP_BASIS_C1 ? 0 : -1;
    const int il = k + hoff, ir = il + 1;                       // hats on the left / right node of element k
    const double hl = (il >= 0 && il < w.nh) ? x[(int64_t)il * xs] : 0.0, hr = (ir >= 0 && ir < w.nh) ? x[(int64_t)ir * xs] : 0.0;
    const double *xp = x + ((int64_t)w.nh + k) * xs;            // bubble 1 of element k
    double *yp = y + (int64_t)k * ys;                           // degree 0 of element k
    // degree a receives + b_{a+1} / (2a+3) - b_{a-1} / (2a-1); b_j = 0 outside 1..q
    double bm = 0.0, b0 = 0.0, bp = q >= 1 ? xp[0] : 0.0;       // b_{a-1}, b_a, b_{a+1} at a = 0
    for (int a = 0; a < N; ++a) {
        double v = bp / (2 * a + 3);
        if (a >= 2) v -= bm / (2 * a - 1);
        if (a == 0) v += 0.5 * (hl + hr);
        if (a == 1) v += 0.5 * (hr - hl);
        *yp = v;
        yp += ystep;
        bm = b0;
        b0 = bp;
        bp = (a + 2 <= q) ? xp[(int64_t)(a + 1) * xstep] : 0.0;   // b_{a+2}
    }
}

extern "C" int pop_c1_to_legendre(pop_ctx *ctx, int basis, int m, int N, int side, const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t cnt) {
    POP_REQUIRE(ctx && dX && dY, POP_ERR_ARG, "pop_c1_to_legendre: null argument");
    POP_REQUIRE(basis == POP_BASIS_C1 || basis == POP_BASIS_DIRICHLET, POP_ERR_ARG, "pop_c1_to_legendre: unknown basis %d", basis);
    POP_REQUIRE(side == POP_LEFT || side == POP_RIGHT, POP_ERR_ARG, "pop_c1_to_legendre: bad side");
    const int nh = basis == POP_BASIS_C1 ? m + 1 : m - 1;
    POP_REQUIRE(m >= 1 && nh >= 1 && N >= 2, POP_ERR_DIM, "pop_c1_to_legendre: need m >= 1 elements with hats and N >= 2 Legendre degrees per element");
    const int64_t n = (int64_t)nh + (int64_t)m * (N - 2), np = (int64_t)m * N;
    POP_REQUIRE(cnt >= 0, POP_ERR_DIM, "pop_c1_to_legendre: negative count");
    if (side == POP_LEFT) POP_REQUIRE(ldx >= n && ldy >= np, POP_ERR_DIM, "pop_c1_to_legendre: DimensionMismatch: X needs %lld rows, Y %lld", (long long)n, (long long)np);
    else POP_REQUIRE(ldx >= std::max<int64_t>(cnt, 1) && ldy >= std::max<int64_t>(cnt, 1), POP_ERR_DIM, "pop_c1_to_legendre: leading dimensions smaller than the row count");
    if (cnt == 0) return POP_OK;
    WeakOp w = {basis, m, N, nh, N - 2, 0.0, nullptr};
    const dim3 blk(32, 8);
    if (side == POP_LEFT) {
        POP_REQUIRE((cnt + 7) / 8 <= 65535, POP_ERR_DIM, "pop_c1_to_legendre: at most 524280 columns per call");
        k_c1_to_legendre<true><<<dim3((unsigned)((m + 31) / 32), (unsigned)((cnt + 7) / 8)), blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    } else {
        POP_REQUIRE((m + 7) / 8 <= 65535, POP_ERR_DIM, "pop_c1_to_legendre: at most 524280 elements");
        k_c1_to_legendre<false><<<dim3((unsigned)((cnt + 31) / 32), (unsigned)((m + 7) / 8)), blk, 0, ctx->stream>>>(w, dX, ldx, dY, ldy, cnt);
    }
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-cell tensor Legendre analysis  C = T V T^T
// ---------------------------------------------------------------------------------------------------
#define AN_TILE 4
// Shared memory (pp = p rounded up to 4, rows padded by 4 doubles against bank conflicts of the transposing stores):
//   Tt[s][a] = T[a][s] ; Vt[s][t] = V[s][t] of the cell (t fastest) ; Wt[t][a] = (T V)[a][t]
// 256 threads, each a 4x4 tile of a 64x64 product: rows {2 ta, 2 ta + 1, 32 + 2 ta, 33 + 2 ta} (so the 16 lanes that
// differ in ta read 16 consecutive 16-byte words: conflict-free 128-bit loads), columns 4 tb .. 4 tb + 3 (two values
// per warp: broadcast).  Two passes of p rank-1 updates.  Persistent CTAs loop over the cells (x fastest, so the CTAs
// that complete one 32 B sector of the interlaced output run at the same time); T is staged once per CTA.
// SYNTH: the inverse direction (`legendretransform \ dat`): the cell's coefficients are gathered from the interlaced
// layout, T is the synthesis matrix (values = T * coefficients), the values go to the cell's block of V.
template <bool SYNTH>
__global__ void __launch_bounds__(256, 2) k_analysis_2d(int n, int p, const double *__restrict__ T, double *__restrict__ V, int64_t ldv,
                                                        double *__restrict__ F, int64_t ldf) {
    extern __shared__ __align__(16) double an_sm[];
    const int pp = (p + 3) & ~3, ldp = 68;   // p <= 64
    double *Tt = an_sm, *Vt = Tt + (size_t)64 * ldp, *Wt = Vt + (size_t)64 * ldp;
    const int tid = threadIdx.x;
    // T is p x p column-major (T[a + p s]); transposed on the way in; zero padding up to 64 x 64
    for (int e = tid; e < 64 * 64; e += 256) {
        const int a = e & 63, s = e >> 6;
        Tt[s * ldp + a] = (a < p && s < p) ? T[a + (size_t)p * s] : 0.0;
    }
    const int ta = tid & 15, tb = tid >> 4;
    const int a0 = 2 * ta, a1 = 32 + 2 * ta, t0 = tb * AN_TILE;
    const bool on = t0 < pp;
    const int64_t ncell = (int64_t)n * n;
    for (int64_t cell = blockIdx.x; cell < ncell; cell += gridDim.x) {
        const int ci = (int)(cell % n), cj = (int)(cell / n);
        // values of the cell: column t of the cell is p consecutive doubles (node s fastest)
        double *Vc = V + (int64_t)ci * p + (int64_t)cj * p * ldv;
        double *Fc = F + (int64_t)ci + (int64_t)cj * ldf;   // interlaced: degree a of cell ci at row ci + n a
        __syncthreads();   // the previous cell's second pass has read Vt/Wt
        for (int e = tid; e < 64 * pp; e += 256) {
            const int sI = e & 63, t = e >> 6;
            double v = 0.0;
            if (sI < p && t < p) v = SYNTH ? Fc[(int64_t)n * sI + (int64_t)n * t * ldf] : Vc[sI + (int64_t)t * ldv];
            Vt[sI * ldp + t] = v;
        }
        __syncthreads();
        double acc[AN_TILE][AN_TILE];
#pragma unroll
        for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
            for (int y = 0; y < AN_TILE; ++y) acc[x][y] = 0.0;
        if (on) {
            // W[a][t] = sum_s T[a][s] V[s][t]
            for (int sI = 0; sI < p; ++sI) {
                const double2 ta01 = *reinterpret_cast<const double2 *>(Tt + sI * ldp + a0), ta23 = *reinterpret_cast<const double2 *>(Tt + sI * ldp + a1);
                const double2 vb01 = *reinterpret_cast<const double2 *>(Vt + sI * ldp + t0), vb23 = *reinterpret_cast<const double2 *>(Vt + sI * ldp + t0 + 2);
                const double av[4] = {ta01.x, ta01.y, ta23.x, ta23.y}, bv[4] = {vb01.x, vb01.y, vb23.x, vb23.y};
#pragma unroll
                for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
                    for (int y = 0; y < AN_TILE; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
            }
#pragma unroll
            for (int y = 0; y < AN_TILE; ++y) {
                *reinterpret_cast<double2 *>(Wt + (t0 + y) * ldp + a0) = make_double2(acc[0][y], acc[1][y]);
                *reinterpret_cast<double2 *>(Wt + (t0 + y) * ldp + a1) = make_double2(acc[2][y], acc[3][y]);
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
                const double2 wa01 = *reinterpret_cast<const double2 *>(Wt + t * ldp + a0), wa23 = *reinterpret_cast<const double2 *>(Wt + t * ldp + a1);
                const double2 tb01 = *reinterpret_cast<const double2 *>(Tt + t * ldp + b0), tb23 = *reinterpret_cast<const double2 *>(Tt + t * ldp + b0 + 2);
                const double av[4] = {wa01.x, wa01.y, wa23.x, wa23.y}, bv[4] = {tb01.x, tb01.y, tb23.x, tb23.y};
#pragma unroll
                for (int x = 0; x < AN_TILE; ++x)
#pragma unroll
                    for (int y = 0; y < AN_TILE; ++y) acc[x][y] = fma(av[x], bv[y], acc[x][y]);
            }
            // interlaced store: degree a of cell ci at row ci + n a (analysis); node a of the cell (synthesis)
#pragma unroll
            for (int y = 0; y < AN_TILE; ++y)
#pragma unroll
                for (int x = 0; x < AN_TILE; ++x) {
                    const int a = (x < 2 ? a0 : a1 - 2) + x, b = b0 + y;
                    if (a < p && b < p) {
                        if (SYNTH) Vc[a + (int64_t)b * ldv] = acc[x][y];
                        else Fc[(int64_t)n * a + (int64_t)n * b * ldf] = acc[x][y];
                    }
                }
        }
    }
}

// Cells of more than 64 nodes per dimension (examples/adi_poisson.jl:87 uses p = 300 on four cells): the per-cell kernel
// above keeps a whole cell on chip, so these go through two tiled passes with the cell matrix T applied from the left, then from
// the right, and an intermediate in global memory:  W[i p + a, c] = sum_k T[a,k] in[row(i,k), c]  and
// out[row(r), col(j,b)] = sum_k W[r, col(j,k)] T[b,k].  64 x 64 output tiles, 16 x 16 threads with 4 x 4 accumulators each,
// k in chunks of 16.  *_il: the index runs interlaced (cell + n * degree) instead of cell-major (cell * p + node).
template <bool RIGHT>
__global__ void __launch_bounds__(256) k_cell_apply(int n, int p, const double *__restrict__ T, const double *__restrict__ in, int64_t ldi, double *__restrict__ out,
                                                    int64_t ldo, int in_il, int out_il, int out_rows_il) {
    __shared__ double As[16][65], Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, cell = blockIdx.z;
    const int64_t np = (int64_t)n * p;
    const int64_t t0 = (int64_t)blockIdx.x * 64;   // LEFT: first column of the tile; RIGHT: first row
    const int b0 = blockIdx.y * 64;                // first row (LEFT) / column (RIGHT) of T
    auto idx = [&](int il, int c, int k) -> int64_t { return il ? (int64_t)c + (int64_t)n * k : (int64_t)c * p + k; };
    double acc[4][4] = {};
    for (int k0 = 0; k0 < p; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = e & 15, x = e >> 4, k = k0 + kk;
            // As[kk][x]: the operand indexed by the output's first index, Bs[kk][x]: by its second
            double av = 0.0, bv = 0.0;
            if (k < p) {
                if (!RIGHT) {
                    if (b0 + x < p) av = T[(b0 + x) + (int64_t)p * k];
                    if (t0 + x < np) bv = in[idx(in_il, cell, k) + (t0 + x) * ldi];
                } else {
                    if (t0 + x < np) av = in[(t0 + x) + idx(in_il, cell, k) * ldi];
                    if (b0 + x < p) bv = T[(b0 + x) + (int64_t)p * k];
                }
            }
            As[kk][x] = av;
            Bs[kk][x] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = As[kk][ty * 4 + u];
                b[u] = Bs[kk][tx * 4 + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (!RIGHT) {
                const int a = b0 + ty * 4 + u;
                const int64_t c = t0 + tx * 4 + v;
                if (a < p && c < np) out[idx(out_il, cell, a) + c * ldo] = acc[u][v];
            } else {
                const int64_t r = t0 + ty * 4 + u;
                const int b = b0 + tx * 4 + v;
                if (r < np && b < p) {
                    const int64_t ro = out_rows_il ? (r / p) + (int64_t)n * (r % p) : r;
                    out[ro + idx(out_il, cell, b) * ldo] = acc[u][v];
                }
            }
        }
}

static int cell_transform_2d_tiled(pop_ctx *ctx, bool synth, int n, int p, const double *T_host, double *dV, int64_t ldv, double *dF, int64_t ldf) {
    const int64_t np = (int64_t)n * p;
    double *dT = nullptr, *W = nullptr;
    POP_CUDA(cudaMallocAsync((void **)&dT, ((size_t)p * p + (size_t)np * np) * sizeof(double), ctx->stream));
    W = dT + (size_t)p * p;
    cudaError_t e = cudaMemcpyAsync(dT, T_host, (size_t)p * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const dim3 grid((unsigned)((np + 63) / 64), (unsigned)((p + 63) / 64), (unsigned)n);
        if (!synth) {   // values (cell-major) -> W (rows cell-major) -> coefficients (interlaced)
            k_cell_apply<false><<<grid, 256, 0, ctx->stream>>>(n, p, dT, dV, ldv, W, np, 0, 0, 0);
            k_cell_apply<true><<<grid, 256, 0, ctx->stream>>>(n, p, dT, W, np, dF, ldf, 0, 1, 1);
        } else {        // coefficients (interlaced) -> W (rows cell-major, columns still interlaced) -> values (cell-major)
            k_cell_apply<false><<<grid, 256, 0, ctx->stream>>>(n, p, dT, dF, ldf, W, np, 1, 0, 0);
            k_cell_apply<true><<<grid, 256, 0, ctx->stream>>>(n, p, dT, W, np, dV, ldv, 1, 0, 0);
        }
        ctx->launches += 2;
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);   // T_host may be reused at once
    cudaFreeAsync(dT, ctx->stream);
    if (e != cudaSuccess || e2 != cudaSuccess) { pop_set_error("cell transform (tiled): %s", cudaGetErrorString(e != cudaSuccess ? e : e2)); return POP_ERR_CUDA; }
    return POP_OK;
}

static int cell_transform_2d(pop_ctx *ctx, bool synth, int n, int p, const double *T_host, double *dV, int64_t ldv, double *dF, int64_t ldf) {
    const char *who = synth ? "pop_legendre_synthesis_2d" : "pop_legendre_analysis_2d";
    POP_REQUIRE(ctx && T_host && dV && dF, POP_ERR_ARG, "%s: null argument", who);
    POP_REQUIRE(n >= 1 && p >= 1, POP_ERR_ARG, "%s: need n >= 1 cells per dimension and p >= 1 nodes per cell", who);
    POP_REQUIRE(p <= 4096, POP_ERR_DIM, "%s: p = %d nodes per cell is out of range", who, p);
    const int64_t np = (int64_t)n * p;
    POP_REQUIRE(ldv >= np && ldf >= np, POP_ERR_DIM, "%s: DimensionMismatch: panels need %lld rows", who, (long long)np);
    if (p > 64) return cell_transform_2d_tiled(ctx, synth, n, p, T_host, dV, ldv, dF, ldf);   // the per-cell kernel holds p <= 64
    double *dT = nullptr;
    int rc = pop_ctx_scratch(ctx, (size_t)p * p * sizeof(double), (void **)&dT);
    if (rc) return rc;
    POP_CUDA(cudaMemcpyAsync(dT, T_host, (size_t)p * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = (size_t)3 * 64 * 68 * sizeof(double);
    const int64_t ncell = (int64_t)n * n;
    const unsigned grid = (unsigned)std::min<int64_t>(ncell, (int64_t)2 * ctx->sm_count);
    if (synth) {
        POP_CUDA(cudaFuncSetAttribute(k_analysis_2d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_analysis_2d<true><<<grid, 256, smem, ctx->stream>>>(n, p, dT, dV, ldv, dF, ldf);
    } else {
        POP_CUDA(cudaFuncSetAttribute(k_analysis_2d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_analysis_2d<false><<<grid, 256, smem, ctx->stream>>>(n, p, dT, dV, ldv, dF, ldf);
    }
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    POP_CUDA(cudaStreamSynchronize(ctx->stream));   // T_host and the scratch copy may be reused at once
    return POP_OK;
}

extern "C" int pop_legendre_analysis_2d(pop_ctx *ctx, int n, int p, const double *T_host, const double *dV, int64_t ldv, double *dF, int64_t ldf) {
    return cell_transform_2d(ctx, false, n, p, T_host, const_cast<double *>(dV), ldv, dF, ldf);
}

extern "C" int pop_legendre_synthesis_2d(pop_ctx *ctx, int n, int p, const double *S_host, const double *dF, int64_t ldf, double *dV, int64_t ldv) {
    return cell_transform_2d(ctx, true, n, p, S_host, dV, ldv, const_cast<double *>(dF), ldf);
}
