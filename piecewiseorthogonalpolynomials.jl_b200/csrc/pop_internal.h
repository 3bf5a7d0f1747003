// Internal declarations shared by the translation units of libpop_arrowhead.so.
// Not part of the ABI (include/pop_arrowhead.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "pop_arrowhead.h"

#define POP_MAXD 8  // largest supported D bandwidth (either side)

void pop_set_error(const char *fmt, ...);

#define POP_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            pop_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return POP_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define POP_REQUIRE(cond, code, ...)   \
    do {                               \
        if (!(cond)) {                 \
            pop_set_error(__VA_ARGS__); \
            return (code);             \
        }                              \
    } while (0)

struct pop_slab {  // one device allocation shared by the handles carved from it
    void *dptr = nullptr;
    int refs = 0;
};

// Device pointers of one packed arrowhead (layout: include/pop_arrowhead.h).
struct ArrowPtrs {
    double *Ad, *Au, *Al;  // [nh], [nh-1], [nh-1]
    double *B, *C;         // [nB][3][m], [nC][3][m]
    double *D;             // [lD+uD+1][q][mD]
    double *rdA;           // [nh]     1/Ad          (valid when has_recip)
    double *rdD;           // [q][mD]  1/D[lD][j][k] (valid when has_recip)
};

struct pop_arrowhead {
    pop_desc d;   // d.uniform is 1 iff mD == 1
    int mD;       // 1 (shared D block) or m
    ArrowPtrs p;
    pop_slab *slab;
    bool has_recip;   // rdA/rdD are up to date
    bool is_factor;   // produced by reverse Cholesky (upper data hold U)
    int bmask;        // bit t set: slot t of some B block may be non-zero
    bool band1_zero;  // the first super-/sub-diagonal of every D block is identically zero (or absent)
    int64_t n() const { return (int64_t)d.nh + (int64_t)d.m * d.q; }
};

struct pop_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    int max_smem_optin;
    int64_t launches;
    // scratch
    void *d_scratch;
    size_t d_scratch_bytes;
    int *d_info;      // device status words
    int *h_info;      // pinned mirror
    // streams/events for the chunked host variants
    cudaStream_t s_copy[2];
    cudaEvent_t ev[32];   // [0,8) H2D done, [8,16) compute done, [16,24) D2H done, 24, 25: misc
    void *host_stage;     // pinned staging pieces for pageable host panels (pop_host.cu), allocated on first use
};
void pop_host_stage_free(pop_ctx *ctx);

int pop_ctx_scratch(pop_ctx *ctx, size_t bytes, void **out);

// sizes (in doubles) of the fields of a packed arrowhead
struct ArrowSizes {
    size_t Ad, Au, Al, B, C, D, rdA, rdD, total;
};
ArrowSizes pop_sizes(const pop_desc &d, int mD);
// allocate `count` handles with descriptor d from a single slab
int pop_alloc_handles(pop_ctx *ctx, const pop_desc &d, int mD, int count, pop_arrowhead **out);

// ---- operator views handed to the kernels by value ---------------------------------
// Triangular operator: unknowns i and i+d of the per-element chain are coupled by
// W[d-1][i*mD + kD]; hats i,i+1 by offA[i]; bubble (b,k) and hat k+t-1 by S[(b*3+t)*m+k].
struct TriView {
    int m, nh, q, nb, nd, mD;
    int unit;
    const double *rdA, *offA, *S, *rdD;
    const double *W[POP_MAXD];
};
// General operator for products: coefficient of (j, j+d) is Dp[d+POP_MAXD][j*mD+kD].
struct MulView {
    int m, nh, q, nb_row, nb_col, lD, uD, mD;
    const double *Ad, *Aup, *Alo;   // y_i += Ad[i] x_i + Aup[i] x_{i+1} + Alo[i-1] x_{i-1}
    const double *Srow;             // hats <- bubbles   (B of op(A)) [nb_row][3][m]
    const double *Scol;             // bubbles <- hats   (C of op(A)) [nb_col][3][m]
    const double *Dp[2 * POP_MAXD + 1];
};
int pop_make_triview(const pop_arrowhead *A, int uplo, int trans, int unit, TriView *tv);
int pop_make_mulview(const pop_arrowhead *A, int sym, int trans, MulView *mv);
int pop_ensure_recip(pop_ctx *ctx, pop_arrowhead *A);

// ---- launchers (device pointers, async on ctx->stream) -----------------------------
// vector index stride sv, batch stride sc (LEFT: sv=1, sc=ld ; RIGHT: sv=ld, sc=1)
int pop_launch_trsm_staged(pop_ctx *ctx, const TriView &tv, bool backward, double *X, int64_t sv,
                           int64_t sc, int64_t nbatch);
int pop_launch_solve_staged(pop_ctx *ctx, const TriView &up, const TriView &lo, double *X,
                            int64_t sv, int64_t sc, int64_t nbatch);
int pop_launch_mul(pop_ctx *ctx, const MulView &mv, double alpha, const double *X, int64_t svx,
                   int64_t scx, double beta, double *Y, int64_t svy, int64_t scy, int64_t nbatch);
// single-pass on-chip kernel (LEFT only). Returns POP_ERR_DIM if the column does not fit.
bool pop_fused_supported(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T);
int pop_launch_solve_fused(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, double gamma,
                           const double *F, int64_t ldf, double *X, int64_t ldx, int64_t nrhs);
// register-resident single-pass kernel (pop_fused2.cu): LEFT only, shared D block, q <= 128
bool pop_fused2_supported(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, int64_t ldx, const double *X);
int pop_launch_solve_fused2(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *F, int64_t ldf,
                            double *X, int64_t ldx, int64_t nrhs);
int pop_launch_factor(pop_ctx *ctx, pop_arrowhead **S, int count, int32_t *info_out);
int pop_launch_axpby_batch(pop_ctx *ctx, const pop_arrowhead *X, const pop_arrowhead *Y,
                           const double *h_alpha, const double *h_beta, int count, pop_arrowhead **out);
