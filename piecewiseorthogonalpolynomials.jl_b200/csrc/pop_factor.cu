// Reverse Cholesky S = U U^T of packed arrowheads, batched over shifts.
// Replaces MatrixFactorizations.reversecholesky_layout!(::ArrowheadLayouts, A, UpperTriangular)
// (src/arrowhead.jl:276-292), _reverse_chol_lower_B! (:297-344) and _reverse_chol_upper_B! (:346-389).
// Index form: SURVEY.md appendix A.  Everything is in place on the upper data.
#include <climits>
#include <vector>

#include "pop_internal.h"

struct FactorBatch {
    ArrowPtrs p;
};

// Per-element banded reverse Cholesky of D_k (the `Threads.@threads for B in A.D` loop, :277-279).
// thread <-> element k (or the single shared block when mD == 1); blockIdx.y <-> shift.
__global__ void k_factor_D(const FactorBatch *batch, int *info, int *band1, int m, int nh, int q, int uD, int mD) {
    const int s = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= mD) return;
    const ArrowPtrs &p = batch[s].p;
    const size_t plane = (size_t)q * mD;
    double *Dd = p.D;  // band d at Dd + d*plane (lD == 0 for factor handles)
    bool failed = false;   // report only the first bad pivot of this chain (later ones are NaN-poisoned)
    for (int j = q - 1; j >= 0; --j) {
        double sdiag = Dd[(size_t)j * mD + k];
        for (int d = 1; d <= uD; ++d)
            if (j + d < q) {
                double u = Dd[d * plane + (size_t)j * mD + k];
                sdiag -= u * u;
            }
        if (!(sdiag > 0.0)) {
            if (!failed) atomicMin(&info[s], nh + j * m + k + 1);
            failed = true;
            sdiag = nan("");
        }
        const double r = sqrt(sdiag);
        Dd[(size_t)j * mD + k] = r;
        p.rdD[(size_t)j * mD + k] = 1.0 / r;
        for (int d = 1; d <= uD; ++d) {
            const int i = j - d;
            if (i < 0) break;
            double acc = Dd[d * plane + (size_t)i * mD + k];
            const int chi = min(i + uD, q - 1);
            for (int c = j + 1; c <= chi; ++c)
                acc -= Dd[(c - i) * plane + (size_t)i * mD + k] * Dd[(c - j) * plane + (size_t)j * mD + k];
            Dd[d * plane + (size_t)i * mD + k] = acc / r;
        }
    }
    if (uD >= 1) {   // structurally decoupled even/odd chains (mass + Laplacian): remember it
        bool nz = false;
        for (int j = 0; j + 1 < q; ++j) nz |= Dd[plane + (size_t)j * mD + k] != 0.0;
        if (nz) band1[s] = 1;
    }
}

// One block per shift: B back-substitution (:308-323 / :359-374), hat Schur update
// (:326-343 / :378-388) in gather form (no atomics), then the bidiagonal reverse Cholesky of the
// hat block (:289).
__global__ void k_factor_hat(const FactorBatch *batch, int *info, int m, int nh, int q, int nb, int uD, int mD) {
    const int s = blockIdx.x;
    const ArrowPtrs &p = batch[s].p;
    const size_t plane = (size_t)q * mD;
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const int kD = mD == 1 ? 0 : k;
        for (int b = nb - 1; b >= 0; --b) {
            const double inv = 1.0 / p.D[(size_t)b * mD + kD];
            for (int t = 0; t < 3; ++t) {
                double v = p.B[((size_t)b * 3 + t) * m + k];
                for (int bt = b + 1; bt < nb && bt - b <= uD; ++bt)
                    v -= p.B[((size_t)bt * 3 + t) * m + k] * p.D[(size_t)(bt - b) * plane + (size_t)b * mD + kD];
                p.B[((size_t)b * 3 + t) * m + k] = v * inv;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
        double ad = p.Ad[i];
        double au = (i < nh - 1) ? p.Au[i] : 0.0;
        for (int b = 0; b < nb; ++b) {
            for (int t = 0; t < 3; ++t) {
                const int k = i - t + 1;
                if (k < 0 || k >= m) continue;
                const double v = p.B[((size_t)b * 3 + t) * m + k];
                ad -= v * v;
                if (t < 2 && i < nh - 1) au -= v * p.B[((size_t)b * 3 + t + 1) * m + k];
            }
        }
        p.Ad[i] = ad;
        if (i < nh - 1) p.Au[i] = au;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool failed = info[s] != INT_MAX;   // a D block already failed: the reference throws there
        for (int i = nh - 1; i >= 0; --i) {
            double sd = p.Ad[i];
            if (i < nh - 1) {
                const double u = p.Au[i];
                sd -= u * u;
            }
            if (!(sd > 0.0)) {
                if (!failed) atomicMin(&info[s], i + 1);
                failed = true;
                sd = nan("");
            }
            const double r = sqrt(sd);
            p.Ad[i] = r;
            p.rdA[i] = 1.0 / r;
            if (i > 0) p.Au[i - 1] = p.Au[i - 1] / r;
        }
    }
}

__global__ void k_fill_int(int *p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int pop_launch_factor(pop_ctx *ctx, pop_arrowhead **S, int count, int32_t *info_out) {
    POP_REQUIRE(count >= 1 && count <= 4096, POP_ERR_ARG, "factor batch size %d out of range [1,4096]", count);
    const pop_desc d = S[0]->d;
    const int mD = S[0]->mD;
    for (int i = 0; i < count; ++i) {
        POP_REQUIRE(S[i]->d.m == d.m && S[i]->d.nh == d.nh && S[i]->d.q == d.q && S[i]->d.uD == d.uD && S[i]->d.lD == d.lD &&
                        S[i]->d.nB == d.nB && S[i]->mD == mD,
                    POP_ERR_DIM, "factor batch: handles differ in shape");
    }
    std::vector<FactorBatch> host(count);
    for (int i = 0; i < count; ++i) {
        host[i].p = S[i]->p;
        // upper bands start after the lD lower ones
        host[i].p.D = S[i]->p.D + (size_t)d.lD * d.q * mD;
    }
    void *dscr = nullptr;
    int rc = pop_ctx_scratch(ctx, sizeof(FactorBatch) * count, &dscr);
    if (rc) return rc;
    POP_CUDA(cudaMemcpyAsync(dscr, host.data(), sizeof(FactorBatch) * count, cudaMemcpyHostToDevice, ctx->stream));
    POP_REQUIRE(count <= 2048, POP_ERR_ARG, "factor batch size %d out of range", count);
    k_fill_int<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_info, count, INT_MAX);
    k_fill_int<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_info + 2048, count, 0);
    ctx->launches += 2;
    const int nb = d.nB < d.q ? d.nB : d.q;
    if (d.q > 0) {
        dim3 grid((mD + 127) / 128, count);
        k_factor_D<<<grid, 128, 0, ctx->stream>>>((const FactorBatch *)dscr, ctx->d_info, ctx->d_info + 2048, d.m, d.nh, d.q, d.uD, mD);
        ctx->launches++;
    }
    k_factor_hat<<<count, 256, 0, ctx->stream>>>((const FactorBatch *)dscr, ctx->d_info, d.m, d.nh, d.q, nb, d.uD, mD);
    ctx->launches++;
    POP_CUDA(cudaGetLastError());
    POP_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int) * count, cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaMemcpyAsync(ctx->h_info + 2048, ctx->d_info + 2048, sizeof(int) * count, cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; ++i) {
        S[i]->has_recip = true;
        S[i]->is_factor = true;
        S[i]->band1_zero = d.uD >= 1 && ctx->h_info[2048 + i] == 0;
        if (info_out) info_out[i] = ctx->h_info[i] == INT_MAX ? 0 : ctx->h_info[i];
    }
    return POP_OK;
}
