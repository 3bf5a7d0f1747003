/* C restatement of the reference's arrowhead algorithms at full size (CPU oracle, level 2).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and the timed CPU baseline; never by the product.
 * Parity pin: validated against oracle/arrowhead_oracle.py (itself pinned to the reference's
 * golden value, see its header) in tests/test_oracle_c.py.
 *
 * Restates (0-based, packed element-fastest layout of include/pop_arrowhead.h, D per element):
 *   reversecholesky_layout!           /root/reference/src/arrowhead.jl:276-292
 *   _reverse_chol_lower_B!/_upper_B!  :297-344 / :346-389
 *   U \ b, L \ b                      :425-457 / :458-486
 *   mul! (Symmetric view)             :191-235
 * Threading: the reference threads the per-element loops (Threads.@threads, :277,:439) and runs
 * matrix right-hand sides column by column; here columns are distributed over OpenMP threads and
 * the element loop is the contiguous (vectorisable) inner loop -- a favourable CPU port.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DD(d, j, k) D[((size_t)(d) * q + (j)) * m + (k)]
#define BB(b, t, k) B[((size_t)(b) * 3 + (t)) * m + (k)]

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* B <- A' for a column-major rows x cols matrix A (the X' of X / F = (F \ X')', examples/adi_poisson.jl:54). */
void oracle_transpose(const double *A, int64_t lda, int64_t rows, int64_t cols, double *B, int64_t ldb) {
    const int64_t T = 32;
#pragma omp parallel for schedule(static) collapse(2)
    for (int64_t c0 = 0; c0 < cols; c0 += T)
        for (int64_t r0 = 0; r0 < rows; r0 += T) {
            int64_t c1 = c0 + T < cols ? c0 + T : cols, r1 = r0 + T < rows ? r0 + T : rows;
            for (int64_t c = c0; c < c1; ++c)
                for (int64_t r = r0; r < r1; ++r) B[c + r * ldb] = A[r + c * lda];
        }
}

/* In place on upper data: Ad[nh], Au[nh-1], B[nB][3][m], D[uD+1][q][m].  Returns 0 or the 1-based
 * index of a non-positive pivot. */
int oracle_revchol(int m, int nh, int q, int nB, int uD, double *Ad, double *Au, double *B, double *D) {
    int info = 0;
    int nb = nB < q ? nB : q;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < m; ++k) { /* :277-279 */
        for (int j = q - 1; j >= 0; --j) {
            double s = DD(0, j, k);
            for (int d = 1; d <= uD; ++d)
                if (j + d < q) s -= DD(d, j, k) * DD(d, j, k);
            if (!(s > 0)) {
#pragma omp critical
                if (!info) info = nh + j * m + k + 1;
                s = NAN;
            }
            double r = sqrt(s);
            DD(0, j, k) = r;
            for (int d = 1; d <= uD; ++d) {
                int i = j - d;
                if (i < 0) break;
                double acc = DD(d, i, k);
                int chi = i + uD < q - 1 ? i + uD : q - 1;
                for (int c = j + 1; c <= chi; ++c) acc -= DD(c - i, i, k) * DD(c - j, j, k);
                DD(d, i, k) = acc / r;
            }
        }
    }
    if (info) return info;
    for (int b = nb - 1; b >= 0; --b) /* :308-323 / :359-374 */
        for (int k = m - 1; k >= 0; --k) {
            for (int bt = b + 1; bt < nb; ++bt) {
                if (bt - b > uD) continue;
                double Akj = DD(bt - b, b, k);
                if (Akj != 0)
                    for (int t = 0; t < 3; ++t) BB(b, t, k) -= BB(bt, t, k) * Akj;
            }
            double inv = 1.0 / DD(0, b, k);
            for (int t = 0; t < 3; ++t) BB(b, t, k) *= inv;
        }
    for (int b = 0; b < nb; ++b) /* :326-343 / :378-388 */
        for (int k = 0; k < m; ++k)
            for (int t = 0; t < 3; ++t) {
                int i = k + t - 1;
                if (i < 0 || i >= nh) continue;
                Ad[i] -= BB(b, t, k) * BB(b, t, k);
                if (t < 2 && i + 1 < nh) Au[i] -= BB(b, t, k) * BB(b, t + 1, k);
            }
    for (int i = nh - 1; i >= 0; --i) { /* :289 */
        double s = Ad[i] - (i < nh - 1 ? Au[i] * Au[i] : 0.0);
        if (!(s > 0)) return i + 1;
        Ad[i] = sqrt(s);
        if (i > 0) Au[i - 1] /= Ad[i];
    }
    return 0;
}

/* X <- U^{-T} U^{-1} X, X is n x nrhs column-major (ld), U the packed factor. */
void oracle_solve(int m, int nh, int q, int nB, int uD, const double *Ad, const double *Au, const double *B, const double *D,
                  double *X, int64_t ld, int64_t nrhs) {
    int nb = nB < q ? nB : q;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < nrhs; ++c) {
        double *H = X + c * ld, *Z = H + nh;
        /* U \ b : :425-457 */
        for (int j = q - 1; j >= 0; --j) {
            double *zj = Z + (size_t)j * m;
            for (int d = 1; d <= uD; ++d)
                if (j + d < q) {
                    const double *zd = Z + (size_t)(j + d) * m;
                    for (int k = 0; k < m; ++k) zj[k] -= DD(d, j, k) * zd[k];
                }
            for (int k = 0; k < m; ++k) zj[k] /= DD(0, j, k);
        }
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < m; ++k) {
                    int i = k + t - 1;
                    if (i >= 0 && i < nh) H[i] -= BB(b, t, k) * Z[(size_t)b * m + k];
                }
        for (int i = nh - 1; i >= 0; --i) {
            double v = H[i];
            if (i + 1 < nh) v -= Au[i] * H[i + 1];
            H[i] = v / Ad[i];
        }
        /* U' \ b : :458-486 */
        for (int i = 0; i < nh; ++i) {
            double v = H[i];
            if (i > 0) v -= Au[i - 1] * H[i - 1];
            H[i] = v / Ad[i];
        }
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < m; ++k) {
                    int i = k + t - 1;
                    if (i >= 0 && i < nh) Z[(size_t)b * m + k] -= BB(b, t, k) * H[i];
                }
        for (int j = 0; j < q; ++j) {
            double *zj = Z + (size_t)j * m;
            for (int d = 1; d <= uD; ++d)
                if (j - d >= 0) {
                    const double *zd = Z + (size_t)(j - d) * m;
                    for (int k = 0; k < m; ++k) zj[k] -= DD(d, j - d, k) * zd[k];
                }
            for (int k = 0; k < m; ++k) zj[k] /= DD(0, j, k);
        }
    }
}

/* The same solve with the LOOP STRUCTURE of the reference (timing context only, not a checker): a matrix right-hand side is
 * solved column by column (ArrayLayouts has no MatLdivMat method for the arrowhead); inside a column the upper solve threads over the
 * mesh elements (Threads.@threads, src/arrowhead.jl:439) with each element's chain on a stride-m view (b[nh+k : m : end]), and the
 * lower solve runs its element loop serially (:480). */
void oracle_solve_refloop(int m, int nh, int q, int nB, int uD, const double *Ad, const double *Au, const double *B, const double *D,
                          double *X, int64_t ld, int64_t nrhs) {
    int nb = nB < q ? nB : q;
    for (int64_t c = 0; c < nrhs; ++c) {
        double *H = X + c * ld, *Z = H + nh;
#pragma omp parallel for schedule(static)
        for (int k = 0; k < m; ++k) /* :439-441 */
            for (int j = q - 1; j >= 0; --j) {
                double v = Z[(size_t)j * m + k];
                for (int d = 1; d <= uD; ++d)
                    if (j + d < q) v -= DD(d, j, k) * Z[(size_t)(j + d) * m + k];
                Z[(size_t)j * m + k] = v / DD(0, j, k);
            }
        for (int b = 0; b < nb; ++b) /* :449-451 */
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < m; ++k) {
                    int i = k + t - 1;
                    if (i >= 0 && i < nh) H[i] -= BB(b, t, k) * Z[(size_t)b * m + k];
                }
        for (int i = nh - 1; i >= 0; --i) { /* :453 */
            double v = H[i];
            if (i + 1 < nh) v -= Au[i] * H[i + 1];
            H[i] = v / Ad[i];
        }
        for (int i = 0; i < nh; ++i) { /* :472 */
            double v = H[i];
            if (i > 0) v -= Au[i - 1] * H[i - 1];
            H[i] = v / Ad[i];
        }
        for (int b = 0; b < nb; ++b) /* :475-477 */
            for (int t = 0; t < 3; ++t)
                for (int k = 0; k < m; ++k) {
                    int i = k + t - 1;
                    if (i >= 0 && i < nh) Z[(size_t)b * m + k] -= BB(b, t, k) * H[i];
                }
        for (int k = 0; k < m; ++k) /* :480-482, serial in the reference */
            for (int j = 0; j < q; ++j) {
                double v = Z[(size_t)j * m + k];
                for (int d = 1; d <= uD; ++d)
                    if (j - d >= 0) v -= DD(d, j - d, k) * Z[(size_t)(j - d) * m + k];
                Z[(size_t)j * m + k] = v / DD(0, j, k);
            }
    }
}

/* Y <- alpha * Symmetric(P) * X + beta * Y with P given by upper data. */
void oracle_mul_sym(int m, int nh, int q, int nB, int uD, const double *Ad, const double *Au, const double *B, const double *D,
                    double alpha, const double *X, int64_t ldx, double beta, double *Y, int64_t ldy, int64_t nrhs) {
    int nb = nB < q ? nB : q;
    size_t n = (size_t)nh + (size_t)m * q;
#pragma omp parallel
    {
        double *r = (double *)malloc(n * sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int64_t c = 0; c < nrhs; ++c) {
            const double *x = X + c * ldx, *xz = x + nh;
            double *y = Y + c * ldy, *rz = r + nh;
            for (int i = 0; i < nh; ++i) {
                double v = Ad[i] * x[i];
                if (i + 1 < nh) v += Au[i] * x[i + 1];
                if (i > 0) v += Au[i - 1] * x[i - 1];
                r[i] = v;
            }
            for (int j = 0; j < q; ++j)
                for (int k = 0; k < m; ++k) rz[(size_t)j * m + k] = DD(0, j, k) * xz[(size_t)j * m + k];
            for (int d = 1; d <= uD; ++d)
                for (int j = 0; j + d < q; ++j)
                    for (int k = 0; k < m; ++k) {
                        double w = DD(d, j, k);
                        rz[(size_t)j * m + k] += w * xz[(size_t)(j + d) * m + k];
                        rz[(size_t)(j + d) * m + k] += w * xz[(size_t)j * m + k];
                    }
            for (int b = 0; b < nb; ++b)
                for (int t = 0; t < 3; ++t)
                    for (int k = 0; k < m; ++k) {
                        int i = k + t - 1;
                        if (i < 0 || i >= nh) continue;
                        r[i] += BB(b, t, k) * xz[(size_t)b * m + k];
                        rz[(size_t)b * m + k] += BB(b, t, k) * x[i];
                    }
            for (size_t e = 0; e < n; ++e) y[e] = beta == 0.0 ? alpha * r[e] : alpha * r[e] + beta * y[e];
        }
        free(r);
    }
}
