/* Plain-C client of include/pop_arrowhead.h: no Python, no ctypes, no CUDA headers -- what a maintainer's ccall / cgo /
 * JNI stub sees.  README case of the reference (ContinuousPolynomial{1}(range(-1,1;length=4)), N = 10 blocks, S = -Delta + M,
 * test/test_arrowhead.jl:120-134): assemble -> axpby -> download -> reversecholesky -> F \ b through host buffers and through
 * a device-resident panel, checked against a dense matrix built here from the downloaded packed fields and against the
 * known-answer values of SURVEY.md 8c.  Exit codes: 0 ok, 77 no CUDA device (POP_ERR_NOGPU), 1 failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pop_arrowhead.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != POP_OK) {                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, pop_last_error());             \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(void) {
    if (pop_abi_version() != POP_ABI_VERSION) {
        fprintf(stderr, "ABI version %d, header %d\n", pop_abi_version(), POP_ABI_VERSION);
        return 1;
    }
    pop_ctx *ctx = NULL;
    int rc = pop_create(&ctx, 0, NULL);
    if (rc == POP_ERR_NOGPU) {
        printf("abi_smoke: no GPU: %s\n", pop_last_error());
        return strstr(pop_last_error(), "no CPU fallback") ? 77 : 1;
    }
    if (rc != POP_OK) {
        fprintf(stderr, "pop_create -> %d: %s\n", rc, pop_last_error());
        return 1;
    }
    const int m = 3, N = 10;
    const double h = 2.0 / m;
    pop_arrowhead *M = NULL, *L = NULL, *S = NULL, *U = NULL;
    CHECK(pop_assemble(ctx, POP_BASIS_C1, m, N, h, &M, &L));
    CHECK(pop_axpby(ctx, -1.0, L, 1.0, M, &S)); /* S = -Delta + M (upper data) */
    pop_desc d;
    CHECK(pop_arrowhead_desc(S, &d));
    const int nh = d.nh, q = d.q, n = nh + m * q, nbands = d.lD + d.uD + 1;
    if (d.m != m || nh != m + 1 || q != N - 1 || n != 31) {
        fprintf(stderr, "unexpected descriptor m=%d nh=%d q=%d\n", d.m, nh, q);
        return 1;
    }
    /* download (D expanded to [bands][q][m]) and build the dense symmetric matrix from the upper data */
    double *Ad = calloc(nh, 8), *Au = calloc(nh, 8), *B = calloc((size_t)(d.nB > 0 ? d.nB : 1) * 3 * m, 8), *D = calloc((size_t)nbands * q * m, 8);
    CHECK(pop_arrowhead_download(ctx, S, 1, Ad, Au, NULL, d.nB ? B : NULL, NULL, D));
    double *A = calloc((size_t)n * n, 8);
    for (int i = 0; i < nh; ++i) {
        A[i * n + i] = Ad[i];
        if (i + 1 < nh) A[i * n + i + 1] = A[(i + 1) * n + i] = Au[i];
    }
    for (int b = 0; b < d.nB; ++b)
        for (int t = 0; t < 3; ++t)
            for (int k = 0; k < m; ++k) {
                const int hat = k + t - 1, col = nh + b * m + k;
                const double v = B[((size_t)b * 3 + t) * m + k];
                if (hat < 0 || hat >= nh) {
                    if (v != 0.0) { fprintf(stderr, "non-zero slot outside the hats\n"); return 1; }
                    continue;
                }
                A[hat * n + col] = A[col * n + hat] = v;
            }
    for (int dd = 0; dd <= d.uD; ++dd)
        for (int j = 0; j + dd < q; ++j)
            for (int k = 0; k < m; ++k) {
                const int r = nh + j * m + k, c = nh + (j + dd) * m + k;
                A[r * n + c] = A[c * n + r] = D[((size_t)(d.lD + dd) * q + j) * m + k];
            }
    /* F = reversecholesky(Symmetric(S)); x = F \ b with b = (1:31)/31, host buffers */
    CHECK(pop_revchol(ctx, S, &U));
    double *b = malloc(n * 8), *x = malloc(n * 8), *x2 = malloc(n * 8);
    for (int i = 0; i < n; ++i) x[i] = b[i] = (i + 1) / 31.0;
    CHECK(pop_solve_host(ctx, U, POP_LEFT, x, n, 1, POP_SOLVE_AUTO));
    double rn = 0, bn = 0;
    for (int i = 0; i < n; ++i) {
        double s = -b[i];
        for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
        rn += s * s;
        bn += b[i] * b[i];
    }
    const double res = sqrt(rn / bn);
    /* known answers of the README case (independent dense NumPy) */
    const double gold[4][2] = {{0, 0.11098908536056581}, {3, 0.18024989943421074}, {4, 0.06638120377479541}, {30, 3.1655746467356325}};
    double gerr = 0;
    for (int g = 0; g < 4; ++g) gerr = fmax(gerr, fabs(x[(int)gold[g][0]] - gold[g][1]) / fabs(gold[g][1]));
    /* the same through a device-resident panel (three columns; the staged kernels) */
    double *dX = NULL;
    int64_t ld = 0;
    CHECK(pop_panel_alloc(ctx, n, 3, &dX, &ld));
    double *b3 = malloc((size_t)n * 3 * 8);
    for (int c = 0; c < 3; ++c) memcpy(b3 + (size_t)c * n, b, n * 8);
    CHECK(pop_panel_upload(ctx, b3, n, dX, ld, n, 3));
    CHECK(pop_solve(ctx, U, POP_LEFT, dX, ld, 3, POP_SOLVE_STAGED));
    CHECK(pop_panel_download(ctx, dX, ld, b3, n, n, 3));
    double derr = 0;
    for (int i = 0; i < 3 * n; ++i) derr = fmax(derr, fabs(b3[i] - x[i % n]));
    /* product on the device panel: S x = b */
    double *dY = NULL;
    int64_t ldy = 0;
    CHECK(pop_panel_alloc(ctx, n, 3, &dY, &ldy));
    CHECK(pop_mul(ctx, S, POP_LEFT, POP_SYM_UPPER, 0, 1.0, dX, ld, 0.0, dY, ldy, 3));
    CHECK(pop_panel_download(ctx, dY, ldy, b3, n, n, 3));
    double merr = 0;
    for (int i = 0; i < 3 * n; ++i) merr = fmax(merr, fabs(b3[i] - b[i % n]));
    /* error conventions: DimensionMismatch comes back as a status and a message, never as a crash */
    memcpy(x2, b, n * 8);
    rc = pop_solve_host(ctx, U, POP_LEFT, x2, n - 1, 1, POP_SOLVE_AUTO);
    const int err_ok = rc == POP_ERR_DIM && strlen(pop_last_error()) > 0;
    rc = pop_solve_host(ctx, S, POP_LEFT, x2, n, 1, POP_SOLVE_AUTO); /* S is not a factor */
    const int err_ok2 = rc == POP_ERR_ARG;
    printf("abi_smoke: n=%d residual %.2e, known-answer err %.2e, device panel diff %.2e, product err %.2e, launches %lld, errors %s\n", n, res, gerr, derr,
           merr, (long long)pop_launch_count(ctx), (err_ok && err_ok2) ? "ok" : "WRONG");
    CHECK(pop_panel_free(ctx, dX));
    CHECK(pop_panel_free(ctx, dY));
    CHECK(pop_arrowhead_free(ctx, U));
    CHECK(pop_arrowhead_free(ctx, S));
    CHECK(pop_arrowhead_free(ctx, M));
    CHECK(pop_arrowhead_free(ctx, L));
    CHECK(pop_destroy(ctx));
    free(Ad); free(Au); free(B); free(D); free(A); free(b); free(x); free(x2); free(b3);
    return (res < 1e-12 && gerr < 1e-12 && derr < 1e-13 && merr < 1e-12 && err_ok && err_ok2) ? 0 : 1;
}
