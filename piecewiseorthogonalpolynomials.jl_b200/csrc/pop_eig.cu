// Eigenvalues of Symmetric(A) and of the pencil (Symmetric(A), Symmetric(B)) by Lanczos with full
// re-orthogonalisation on the arrowhead kernels.
//
// Replaces, on the device,
//   eigvals(Symmetric(A)), eigvals(Symmetric(A), Symmetric(B))      /root/reference/src/arrowhead.jl:536-537
//     (the reference converts to BandedMatrix and calls LAPACK), and
//   A = U \ (U \ M)';  eigmin(A), eigmax(A)                         /root/reference/examples/adi_poisson.jl:91-94
//     (dense O(n^3) set-up of the ADI spectral interval).
// With B = U U' (reverse Cholesky, the factor is passed in) the pencil A x = lambda B x has the spectrum of the
// symmetric matrix C = U^{-1} A U^{-T}; one Lanczos step is two triangular arrowhead solves and one arrowhead
// product on a single vector, plus a Gram-Schmidt pass against the basis (kernels below).  k = n steps
// tridiagonalise C completely (a breakdown restarts with a fresh orthogonal vector, so multiple eigenvalues are
// found with their multiplicity); k << n gives the extreme eigenvalues with the classical residual bounds
// |lambda - theta_i| <= beta_k |s_{k,i}|.  The k x k tridiagonal eigenproblem is host arithmetic (implicit QL).
#include <algorithm>
#include <cmath>
#include <vector>

#include "pop_internal.h"

// h[j] = v_j . w  for j < nv (v_j = V + j*ldv); block j reduces over n
__global__ void __launch_bounds__(256) k_lz_dots(const double *__restrict__ V, int64_t ldv, const double *__restrict__ w, int64_t n, double *__restrict__ h) {
    const double *v = V + (int64_t)blockIdx.x * ldv;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) acc = fma(v[i], w[i], acc);
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        acc = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
        if (threadIdx.x == 0) h[blockIdx.x] = acc;
    }
}

// w <- w - sum_j h[j] v_j
__global__ void __launch_bounds__(256) k_lz_project(const double *__restrict__ V, int64_t ldv, int nv, const double *__restrict__ h, double *__restrict__ w, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double acc = w[i];
    for (int j = 0; j < nv; ++j) acc = fma(-h[j], V[(int64_t)j * ldv + i], acc);
    w[i] = acc;
}

// dst <- s * src
__global__ void __launch_bounds__(256) k_lz_scale(const double *__restrict__ src, double s, double *__restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = s * src[i];
}

// Eigenvalues of the symmetric tridiagonal (d, e) by the implicit QL algorithm (EISPACK tql2 restricted to
// the LAST row of the eigenvector matrix, which is all the residual bounds need).  d[k], e[k-1]; on return d holds the
// eigenvalues (unsorted) and z[i] the last component of the normalised eigenvector of d[i].
static int tridiag_ql(std::vector<double> &d, std::vector<double> &e_in, std::vector<double> &z) {
    const int n = (int)d.size();
    std::vector<double> e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
    z.assign(n, 0.0);
    if (n) z[n - 1] = 1.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) return -1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + c * f;
                    z[i] = c * z[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return 0;
}

static inline uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void random_normal(std::vector<double> &v, uint64_t &state) {
    for (size_t i = 0; i < v.size(); i += 2) {
        const double u1 = ((splitmix64(state) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const double u2 = (splitmix64(state) >> 11) * (1.0 / 9007199254740992.0);
        const double r = std::sqrt(-2.0 * std::log(u1));
        v[i] = r * std::cos(2.0 * M_PI * u2);
        if (i + 1 < v.size()) v[i + 1] = r * std::sin(2.0 * M_PI * u2);
    }
}

extern "C" int pop_eigvals_lanczos(pop_ctx *ctx, const pop_arrowhead *A, const pop_arrowhead *UB, int k, uint64_t seed, double *theta, double *resid,
                                   int *k_out) {
    POP_REQUIRE(ctx && A && theta && k_out, POP_ERR_ARG, "pop_eigvals_lanczos: null argument");
    const int64_t n = A->n();
    POP_REQUIRE(n >= 1, POP_ERR_DIM, "pop_eigvals_lanczos: empty matrix");
    if (UB) {
        POP_REQUIRE(UB->is_factor, POP_ERR_ARG, "pop_eigvals_lanczos: B must be passed as its reverse-Cholesky factor");
        POP_REQUIRE(UB->d.m == A->d.m && UB->d.nh == A->d.nh && UB->d.q == A->d.q, POP_ERR_DIM, "pop_eigvals_lanczos: A and B differ in size");
    }
    if (k <= 0 || k > n) k = (int)n;
    int rc;
    if (UB && (rc = pop_ensure_recip(ctx, const_cast<pop_arrowhead *>(UB)))) return rc;
    MulView mv;
    pop_make_mulview(A, POP_SYM_UPPER, 0, &mv);
    TriView up, lo;
    if (UB) {
        pop_make_triview(UB, POP_UPPER, 0, 0, &up);   // U \ x   (backward)
        pop_make_triview(UB, POP_UPPER, 1, 0, &lo);   // U' \ x  (forward)
    }
    const int64_t ldv = (n + 1) & ~(int64_t)1;
    double *buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, ((size_t)ldv * (k + 3) + (size_t)k + 8) * sizeof(double));
    if (e != cudaSuccess) { pop_set_error("pop_eigvals_lanczos: %zu bytes for the Lanczos basis: %s", ((size_t)ldv * (k + 3)) * 8, cudaGetErrorString(e)); return POP_ERR_ALLOC; }
    double *V = buf, *w = buf + (size_t)ldv * (k + 1), *t = w + ldv, *h = t + ldv;
    std::vector<double> alpha, beta, hv(n);
    uint64_t state = seed * 0x2545f4914f6cdd1dull + 0x1234567;
    const unsigned gb = (unsigned)((n + 255) / 256);
    double hh[2];
    auto fresh = [&](int nv) -> int {   // V[nv] <- random unit vector orthogonal to V[0..nv)
        for (int attempt = 0; attempt < 4; ++attempt) {
            random_normal(hv, state);
            POP_CUDA(cudaMemcpyAsync(w, hv.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
            POP_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int pass = 0; pass < 2 && nv > 0; ++pass) {
                k_lz_dots<<<nv, 256, 0, ctx->stream>>>(V, ldv, w, n, h);
                k_lz_project<<<gb, 256, 0, ctx->stream>>>(V, ldv, nv, h, w, n);
                ctx->launches += 2;
            }
            k_lz_dots<<<1, 256, 0, ctx->stream>>>(w, ldv, w, n, h);
            ctx->launches++;
            POP_CUDA(cudaMemcpyAsync(hh, h, 8, cudaMemcpyDeviceToHost, ctx->stream));
            POP_CUDA(cudaStreamSynchronize(ctx->stream));
            if (hh[0] > 1e-20 * n) {
                k_lz_scale<<<gb, 256, 0, ctx->stream>>>(w, 1.0 / std::sqrt(hh[0]), V + (size_t)nv * ldv, n);
                ctx->launches++;
                return POP_OK;
            }
        }
        pop_set_error("pop_eigvals_lanczos: could not extend the basis");
        return POP_ERR_ARG;
    };
    rc = fresh(0);
    double scale = 0.0;
    int steps = 0;
    double last_beta = 0.0;
    for (int j = 0; j < k && rc == POP_OK; ++j) {
        double *vj = V + (size_t)j * ldv;
        // w = C v_j,  C = U^{-1} A U^{-T}  (or A)
        if (UB) {
            e = cudaMemcpyAsync(t, vj, n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) { rc = POP_ERR_CUDA; break; }
            if ((rc = pop_launch_trsm_staged(ctx, lo, false, t, 1, ldv, 1))) break;
            if ((rc = pop_launch_mul(ctx, mv, 1.0, t, 1, ldv, 0.0, w, 1, ldv, 1))) break;
            if ((rc = pop_launch_trsm_staged(ctx, up, true, w, 1, ldv, 1))) break;
        } else {
            if ((rc = pop_launch_mul(ctx, mv, 1.0, vj, 1, ldv, 0.0, w, 1, ldv, 1))) break;
        }
        // alpha_j = v_j . w ; full re-orthogonalisation against v_0..v_j, twice
        k_lz_dots<<<1, 256, 0, ctx->stream>>>(vj, ldv, w, n, h + k);
        ctx->launches++;
        for (int pass = 0; pass < 2; ++pass) {
            k_lz_dots<<<j + 1, 256, 0, ctx->stream>>>(V, ldv, w, n, h);
            k_lz_project<<<gb, 256, 0, ctx->stream>>>(V, ldv, j + 1, h, w, n);
            ctx->launches += 2;
        }
        k_lz_dots<<<1, 256, 0, ctx->stream>>>(w, ldv, w, n, h + k + 1);
        ctx->launches++;
        e = cudaMemcpyAsync(hh, h + k, 16, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { pop_set_error("pop_eigvals_lanczos: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; break; }
        const double a = hh[0], b = std::sqrt(std::max(hh[1], 0.0));
        if (!std::isfinite(a) || !std::isfinite(b)) { pop_set_error("pop_eigvals_lanczos: non-finite recurrence coefficient at step %d", j); rc = POP_ERR_ARG; break; }
        alpha.push_back(a);
        scale = std::max(scale, std::fabs(a) + b);
        steps = j + 1;
        last_beta = b;
        if (j + 1 == k) break;
        if (b <= 1e-13 * scale) {            // invariant subspace: restart in its orthogonal complement
            beta.push_back(0.0);
            if ((rc = fresh(j + 1))) break;
            last_beta = 0.0;
        } else {
            beta.push_back(b);
            k_lz_scale<<<gb, 256, 0, ctx->stream>>>(w, 1.0 / b, V + (size_t)(j + 1) * ldv, n);
            ctx->launches++;
        }
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    if (rc) return rc;
    beta.resize(std::max(steps - 1, 0));
    std::vector<double> d = alpha, z;
    if (tridiag_ql(d, beta, z)) { pop_set_error("pop_eigvals_lanczos: the tridiagonal QL iteration did not converge"); return POP_ERR_ARG; }
    std::vector<int> order(steps);
    for (int i = 0; i < steps; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
    for (int i = 0; i < steps; ++i) {
        theta[i] = d[order[i]];
        if (resid) resid[i] = steps == n ? 0.0 : std::fabs(last_beta * z[order[i]]);
    }
    *k_out = steps;
    return POP_OK;
}
