/*
 * pop_arrowhead.h -- C ABI of the B200-native arrowhead hot path.
 *
 * Drop-in boundary for PiecewiseOrthogonalPolynomials.jl's BBBArrowheadMatrix
 * linear algebra (reference: /root/reference/src/arrowhead.jl).  Every entry
 * point cites the reference interface it replaces (file:line relative to the
 * reference root).  Plain pointers and sizes only; no C++/torch types cross it.
 * The Julia-side `ccall` bindings are shown in INTEGRATION.md and julia/.
 *
 * Conventions
 *   - Float64 everywhere.  Vectors / matrices use the reference's own ordering
 *     (src/arrowhead.jl:57-62): nh hats first, then bubble block j (j=0..q-1),
 *     entry k (element, k=0..m-1) at index nh + j*m + k; n = nh + m*q.
 *     Matrices are column-major with leading dimension ld >= rows.
 *   - Return value: 0 ok; >0 = 1-based global index of the first non-positive
 *     pivot (the reference throws PosDefException from the block factor);
 *     <0 = POP_ERR_* with a message available from pop_last_error().
 *     Nothing throws across the boundary.
 *   - A pop_ctx owns one device + one stream; calls on it are serialised on that
 *     stream.  Functions taking device pointers are asynchronous on the stream,
 *     the *_host variants copy in, run, copy out and synchronise.
 *   - The caller owns every host pointer; the library never retains them.
 *
 * Packed arrowhead (all 0-based)
 *   Ad[nh], Au[nh-1], Al[nh-1]      hat block A: diagonal, super-, sub-diagonal
 *   B[nB][3][m]   B[b][t][k] = entry (hat k+t-1 , bubble (b,k)), t=0,1,2
 *   C[nC][3][m]   C[b][t][k] = entry (bubble (b,k), hat k+t-1)
 *                 (hats outside [0,nh) do not exist: those slots must be 0;
 *                  ContinuousPolynomial{1}: nh=m+1, slots t=1,2 used ("lower B",
 *                  src/arrowhead.jl:297); DirichletPolynomial: nh=m-1, slots
 *                  t=0,1 ("upper B", src/arrowhead.jl:346))
 *   D[lD+uD+1][q][mD]  D[lD+d][j][k] = D_k[j, j+d], d=-lD..uD (0 outside);
 *                 mD = m, or 1 when desc.uniform (all elements share one block,
 *                 the `Fill(D, m)` of src/continuouspolynomial.jl:286).
 */
#ifndef POP_ARROWHEAD_H
#define POP_ARROWHEAD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POP_ABI_VERSION 1

enum {
    POP_OK = 0,
    POP_ERR_ARG = -1,    /* bad argument / null pointer                      */
    POP_ERR_DIM = -2,    /* DimensionMismatch                                 */
    POP_ERR_STRUCT = -3, /* structure assertion of the reference violated
                            (src/arrowhead.jl:96-116,299,305,348,356)         */
    POP_ERR_CUDA = -4,   /* CUDA runtime error                                */
    POP_ERR_ALLOC = -5,  /* out of device / host memory                       */
    POP_ERR_NOGPU = -6   /* no CUDA device: there is NO CPU fallback          */
};

enum { POP_BASIS_C1 = 0, POP_BASIS_DIRICHLET = 1 }; /* ContinuousPolynomial{1} / DirichletPolynomial */
enum { POP_LEFT = 0, POP_RIGHT = 1 };               /* A*X, A\X  vs  X*A, X/A */
enum { POP_GENERAL = 0, POP_SYM_UPPER = 1 };        /* Symmetric(P) reads upper data, src/arrowhead.jl:154-167 */
enum { POP_UPPER = 0, POP_LOWER = 1 };
enum { POP_SOLVE_AUTO = 0, POP_SOLVE_STAGED = 1, POP_SOLVE_FUSED = 2,
       POP_SOLVE_FUSED_SMEM = 3 }; /* kernel choice for pop_solve: FUSED = best single-pass kernel,
                                       FUSED_SMEM = the shared-memory-resident variant */

typedef struct pop_ctx pop_ctx;
typedef struct pop_arrowhead pop_arrowhead; /* device-resident packed arrowhead or factor */

typedef struct pop_desc {
    int32_t m;       /* mesh elements = length(A.D)                           */
    int32_t nh;      /* hats = size(A.A,1)                                    */
    int32_t q;       /* bubbles per element = size(A.D[1],1) = N-1            */
    int32_t nB;      /* length(A.B)                                           */
    int32_t nC;      /* length(A.C)                                           */
    int32_t lD, uD;  /* bandwidths(A.D[1])                                    */
    int32_t uniform; /* 1: D given once ([bands][q]) and shared by all k      */
} pop_desc;

/* ---- context -------------------------------------------------------------------- */
int pop_abi_version(void);
const char *pop_last_error(void);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL to
 * let the context create its own non-blocking stream. */
int pop_create(pop_ctx **ctx, int device, void *stream);
int pop_destroy(pop_ctx *ctx);
/* use `stream` from now on, exactly as given: NULL (0) is the CUDA legacy default stream */
int pop_set_stream(pop_ctx *ctx, void *stream);
int pop_sync(pop_ctx *ctx);
/* number of kernels this context has launched so far (bench `gpu_launches`) */
int64_t pop_launch_count(const pop_ctx *ctx);

/* ---- device-resident panels --------------------------------------------------------
 * For hosts without a CUDA binding of their own (julia/PopArrowheadGPU.jl `DevicePanel`): a rows x cols Float64
 * matrix in device memory, column-major with the (even) leading dimension returned in *ld, zero-initialised.
 * Chains of calls on device pointers (pop_mul, pop_solve, pop_adi_poisson, pop_heat_steps ...) then move no data over
 * PCIe; the reference keeps its matrices in host memory, so this is the one place the drop-in adds a type.
 * upload / download copy a host matrix (leading dimension ldh) in / out and synchronise. */
int pop_panel_alloc(pop_ctx *ctx, int64_t rows, int64_t cols, double **dptr, int64_t *ld);
int pop_panel_free(pop_ctx *ctx, double *dptr);
int pop_panel_upload(pop_ctx *ctx, const double *host, int64_t ldh, double *dptr, int64_t ldd, int64_t rows,
                     int64_t cols);
int pop_panel_download(pop_ctx *ctx, const double *dptr, int64_t ldd, double *host, int64_t ldh, int64_t rows,
                       int64_t cols);

/* ---- container (src/arrowhead.jl:31-41,89-120) --------------------------------- */
/* BBBArrowheadMatrix(A,B,C,D): validates what the constructor asserts.  Any of
 * Au/Al/B/C may be NULL (= zero / empty). */
int pop_arrowhead_upload(pop_ctx *ctx, const pop_desc *desc, const double *Ad, const double *Au,
                         const double *Al, const double *B, const double *C, const double *D,
                         pop_arrowhead **out);
int pop_arrowhead_desc(const pop_arrowhead *A, pop_desc *desc);
/* copies the fields back in the layout `desc` of pop_arrowhead_desc describes
 * (D expanded to [bands][q][m] unless expand_uniform == 0).  NULL = skip field. */
int pop_arrowhead_download(pop_ctx *ctx, const pop_arrowhead *A, int expand_uniform, double *Ad,
                           double *Au, double *Al, double *B, double *C, double *D);
int pop_arrowhead_free(pop_ctx *ctx, pop_arrowhead *A);
/* The upper data of A already hold a reverse-Cholesky factor U (S = U U^T), e.g. `F.U` of a ReverseCholesky the
 * reference computed on the host (src/arrowhead.jl:276-292): make the handle usable by pop_solve / pop_solve_mul_axpy
 * without factoring again.  >0: 1-based index of a non-positive pivot. */
int pop_arrowhead_as_factor(pop_ctx *ctx, pop_arrowhead *A);
/* copy(A), src/arrowhead.jl:64 ; reversecholcopy, :266-272 */
int pop_arrowhead_copy(pop_ctx *ctx, const pop_arrowhead *A, pop_arrowhead **out);

/* ---- assembly on a uniform mesh (upper data, as the reference stores it) --------
 * mass:  grammatrix(C)[KR,KR]   src/continuouspolynomial.jl:274-290, src/dirichlet.jl:180-196
 * lap :  weaklaplacian(C)[KR,KR] src/continuouspolynomial.jl:348-357, src/dirichlet.jl:169-178
 * KR = Block.(oneto(N)) (src/arrowhead.jl:141-149); h = step(points).  Either output may be NULL. */
int pop_assemble(pop_ctx *ctx, int basis, int m, int N, double h, pop_arrowhead **mass,
                 pop_arrowhead **lap);

/* Non-uniform mesh: h_elems[k] = points[k+1]-points[k] (host array of m element lengths).  One D block per element
 * (desc.uniform = 0).  mass: the reference's generic grammatrix L'*G*L (src/continuouspolynomial.jl:259-265,293-297);
 * lap: -(diff C)' G (diff C) with the per-element scalings of src/continuouspolynomial.jl:318-325 (the reference
 * defines weaklaplacian itself for ranges only).  Synchronises the context's stream. */
int pop_assemble_mesh(pop_ctx *ctx, int basis, int m, int N, const double *h_elems, pop_arrowhead **mass,
                      pop_arrowhead **lap);

/* The same with a piecewise-constant diffusion coefficient a_elems[k] > 0 on element k (NULL = 1): lap becomes
 * -(D*C)' * (a .* (D*C)) of examples/heat.jl:77-79 for a in span of the piecewise constants (every element integral of
 * the weak Laplacian scaled by a_k); the mass matrix is unchanged. */
int pop_assemble_mesh_coef(pop_ctx *ctx, int basis, int m, int N, const double *h_elems, const double *a_elems,
                           pop_arrowhead **mass, pop_arrowhead **lap);

/* ---- algebra: alpha*X + beta*Y with ragged B/C and D bands (src/arrowhead.jl:392-417) */
int pop_axpby(pop_ctx *ctx, double alpha, const pop_arrowhead *X, double beta,
              const pop_arrowhead *Y, pop_arrowhead **out);

/* ---- reverse Cholesky S = U U^T (src/arrowhead.jl:276-389) ----------------------
 * reversecholesky(Symmetric(S)): reads the upper data of S, returns a new handle
 * holding U (same sparsity as triu(S)) plus reciprocal pivots. */
int pop_revchol(pop_ctx *ctx, const pop_arrowhead *S, pop_arrowhead **U);
/* reversecholesky!(Symmetric(S)): in place */
int pop_revchol_inplace(pop_ctx *ctx, pop_arrowhead *S);
/* factors of alpha[i]*K + sigma[i]*M for i < nshift in ONE batched launch sequence
 * (the per-shift `reversecholesky(Symmetric(C - B/p[j]))` of examples/adi_poisson.jl:54-55).
 * U[i] receives nshift handles; info[i] (may be NULL) the per-shift status. */
int pop_revchol_shifted_batch(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M,
                              const double *alpha, const double *sigma, int nshift,
                              pop_arrowhead **U, int32_t *info);

/* ---- products (src/arrowhead.jl:191-258) -----------------------------------------
 * side LEFT :  Y <- alpha*op(A)*X + beta*Y, X,Y are n x nrhs
 * side RIGHT:  Y <- alpha*X*op(A) + beta*Y, X,Y are nrhs x n
 * op(A) = A or A^T (trans); sym = POP_SYM_UPPER treats A as Symmetric(A). */
int pop_mul(pop_ctx *ctx, const pop_arrowhead *A, int side, int sym, int trans, double alpha,
            const double *dX, int64_t ldx, double beta, double *dY, int64_t ldy, int64_t nrhs);
int pop_mul_host(pop_ctx *ctx, const pop_arrowhead *A, int side, int sym, int trans, double alpha,
                 const double *X, int64_t ldx, double beta, double *Y, int64_t ldy, int64_t nrhs);

/* ---- triangular solves (src/arrowhead.jl:425-486) -------------------------------
 * X <- op(T)^{-1} X (LEFT) or X <- X op(T)^{-1} (RIGHT) with
 * T = UpperTriangular(A) / LowerTriangular(A) (uplo), Unit* when unit != 0. */
int pop_trsm(pop_ctx *ctx, const pop_arrowhead *A, int side, int uplo, int trans, int unit,
             double *dX, int64_t ldx, int64_t nrhs);
int pop_trsm_host(pop_ctx *ctx, const pop_arrowhead *A, int side, int uplo, int trans, int unit,
                  double *X, int64_t ldx, int64_t nrhs);

/* ---- ReverseCholesky solve:  F \ X  and  X / F  (call sites
 * test/test_arrowhead.jl:130, examples/adi_poisson.jl:54-55): X <- U^{-T} U^{-1} X.
 * `algo`: POP_SOLVE_AUTO picks the single-pass fused kernel when the column fits
 * on chip, else the staged kernels. */
int pop_solve(pop_ctx *ctx, const pop_arrowhead *U, int side, double *dX, int64_t ldx,
              int64_t nrhs, int algo);
int pop_solve_host(pop_ctx *ctx, const pop_arrowhead *U, int side, double *X, int64_t ldx,
                   int64_t nrhs, int algo);
/* `Y = F \ X` / `Y = X / F` with X left untouched (what the reference's allocating `\` does; pop_solve_host is `ldiv!`):
 * the two directions of the host link work on different buffers.  Y may be X. */
int pop_solve_host_to(pop_ctx *ctx, const pop_arrowhead *U, int side, const double *X, int64_t ldx, double *Y,
                      int64_t ldy, int64_t nrhs, int algo);
/* nshift systems, shift i acting on the panel dX + i*stride_shift (BASELINE config 3) */
int pop_solve_shift_batch(pop_ctx *ctx, pop_arrowhead *const *U, int nshift, double *dX,
                          int64_t ldx, int64_t stride_shift, int64_t nrhs, int algo);

/* ---- fused ADI half step (examples/adi_poisson.jl:54-55) ------------------------
 * dR <- T * (U^{-T} U^{-1} dR) + gamma * dF, column-wise, T = Symmetric(Tm) (upper
 * data).  T may be NULL (plain solve).  One read of R, one read of F, one write. */
int pop_solve_mul_axpy(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *Tm, double gamma,
                       const double *dF, int64_t ldf, double *dR, int64_t ldr, int64_t nrhs);

/* ---- right-hand-side pipeline of examples/adi_poisson.jl ------------------------------
 * analysis_2D (examples/adi_poisson.jl:60-76): dV is the (n p) x (n p) matrix of function values, entry
 * (i p + s, j p + t) = f at node s of cell i in x and node t of cell j in y; T_host (p x p, column-major, host)
 * is the analysis matrix of the cell grid (`plan_grid_transform(legendre(0..dx), (p,p))`: coefficients = T * values).
 * dF receives the tensor Legendre coefficients interlaced like a block vector: F[i + n a, j + n b]
 * (`F[i+1:n:n*p, j+1:n:n*p] = T * f_.(z, z')`).  p <= 64: one fused per-cell kernel; larger cells (the example's own p = 300):
 * two tiled passes through an (n p)^2 intermediate.  Synchronises the stream. */
int pop_legendre_analysis_2d(pop_ctx *ctx, int n, int p, const double *T_host, const double *dV, int64_t ldv,
                             double *dF, int64_t ldf);
/* (Q'P)[KR,KR] = (P\Q)' (P'P), Q = ContinuousPolynomial{1} or DirichletPolynomial, P = ContinuousPolynomial{0}
 * (examples/adi_poisson.jl:105; P\C: src/continuouspolynomial.jl:226-235, C\Q: src/dirichlet.jl:35-42, P'P:
 * src/continuouspolynomial.jl:259-272), KR = Block.(oneto(N)):
 *   side LEFT :  Y (n x cnt)  <- (Q'P) X ,   X is (m N) x cnt in P ordering (degree a of element k at a m + k)
 *   side RIGHT:  Y (cnt x n)  <- X (Q'P)',   X is cnt x (m N)
 * so F = (Q'P) fp (Q'P)' is one LEFT and one RIGHT call.  h: uniform step; h_elems (host, [m]) overrides it. */
int pop_weakform_apply(pop_ctx *ctx, int basis, int m, int N, double h, const double *h_elems, int side,
                       const double *dX, int64_t ldx, double *dY, int64_t ldy, int64_t cnt);

/* Legendre coefficients of a continuous piecewise polynomial (P ordering, N degrees per element) -> coefficients in
 * ContinuousPolynomial{1} / DirichletPolynomial blocks 1..N-1 (n = nh + m (N-2)): the `R \ dat` + interlacing of
 * ContinuousPolynomialTransform (src/continuouspolynomial.jl:84-146, adaptivetransform_ldiv :186-213;
 * src/dirichlet.jl:44-53,65-69 drops the end hats).  side LEFT: vectors are the columns of X ((m N) x cnt), RIGHT: the rows
 * (cnt x (m N)); a 2D field takes one call per side.  max_jump (host, may be NULL) receives the largest mismatch between
 * the values the two elements of an interior node imply (and, for DirichletPolynomial, the largest end value): the
 * reference throws ArgumentError("Discontinuity in data ...") when it exceeds 1000 M eps (:105-107).  A non-NULL
 * max_jump synchronises the stream. */
int pop_legendre_to_c1(pop_ctx *ctx, int basis, int m, int N, int side, const double *dX, int64_t ldx, double *dY,
                       int64_t ldy, int64_t cnt, double *max_jump);

/* INVERSE transforms (`Pl \ cfs`, src/continuouspolynomial.jl:129-146,171-184; src/dirichlet.jl:97-110,143-163;
 * test/test_continuouspolynomial.jl:19,26,33: `F \ (F * vals) ≈ vals`).
 * pop_c1_to_legendre: coefficients in ContinuousPolynomial{1} / DirichletPolynomial blocks 1..N-1 (n = nh + m (N-2)) ->
 * Legendre coefficients, N degrees per element in P ordering: the `Pl.R \ dat` step, i.e. the truncated conversion
 * (P \ C)[Block.(1:N), Block.(1:N-1)] of src/continuouspolynomial.jl:226-235.  side as in pop_legendre_to_c1 (its inverse).
 * pop_legendre_synthesis_2d: `legendretransform \ dat` for a 2D field: dF holds the tensor Legendre coefficients interlaced
 * as pop_legendre_analysis_2d writes them, S_host (p x p, column-major, host) is the synthesis matrix of the cell grid
 * (values = S * coefficients, S[s,a] = P_a(z_s), the inverse of the analysis matrix T); dV receives the values
 * (entry (i p + s, j p + t) = node s of cell i, node t of cell j).  Synchronises the stream. */
int pop_c1_to_legendre(pop_ctx *ctx, int basis, int m, int N, int side, const double *dX, int64_t ldx, double *dY,
                       int64_t ldy, int64_t cnt);
int pop_legendre_synthesis_2d(pop_ctx *ctx, int n, int p, const double *S_host, const double *dF, int64_t ldf,
                              double *dV, int64_t ldv);

/* out-of-place transpose dst (cols x rows, ldd) <- src^T, src rows x cols (lds) */
int pop_transpose(pop_ctx *ctx, const double *dsrc, int64_t lds, int64_t rows, int64_t cols,
                  double *ddst, int64_t ldd);

/* ---- drivers ----------------------------------------------------------------------
 * ADI shifts and iteration count: examples/adi_poisson.jl:11-40,47-49 (host arithmetic).
 * p,q must hold J doubles. */
int pop_adi_num_iterations(double a, double b, double c, double d, double tol);
int pop_adi_shifts(int J, double a, double b, double c, double d, double *p, double *q);
/* ADI(A=M, B=-M, C=K, F, a, b, c, d, tol) of examples/adi_poisson.jl:42-59 on one GPU:
 * solves K X M + M X K = F for X (n x n, device, column-major).  J <= 0: derive J from tol. */
int pop_adi_poisson(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, const double *dF,
                    int64_t ldf, double *dX, int64_t ldx, double a, double b, double c, double d,
                    double tol, int J);
/* the same with F and X in HOST memory (one copy in, the whole iteration on the device, one copy out) */
int pop_adi_poisson_host(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, const double *F,
                         int64_t ldf, double *X, int64_t ldx, double a, double b, double c, double d,
                         double tol, int J);
/* backward Euler with one reused factorisation, examples/heat.jl:58-63:
 * nsteps times  U <- S^{-1} (M U)  with S = M + dt*K; dims = 1: columns of dU (n x nrhs);
 * dims = 2: n x n field, U <- S^{-1} (M U M) S^{-1} (factored 2D extension). */
int pop_heat_steps(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, double dt,
                   int nsteps, int dims, double *dU, int64_t ldu, int64_t nrhs);

/* theta scheme (M + theta dt K) u' = (M - (1-theta) dt K) u with one reused factorisation: theta = 1 is pop_heat_steps,
 * theta = 1/2 the trapezoid rule `A \ (B * u)`, A = M - h/2 Delta, B = M + h/2 Delta of examples/heat.jl:26-45. */
int pop_heat_steps_theta(pop_ctx *ctx, const pop_arrowhead *K, const pop_arrowhead *M, double dt, double theta,
                         int nsteps, int dims, double *dU, int64_t ldu, int64_t nrhs);

/* ---- eigenvalues ----------------------------------------------------------------------
 * eigvals(Symmetric(A)) and eigvals(Symmetric(A), Symmetric(B)) of src/arrowhead.jl:536-537, and the dense
 * `A = U \ (U \ M)'; eigmin(A), eigmax(A)` of examples/adi_poisson.jl:91-94, by Lanczos with full
 * re-orthogonalisation on the arrowhead kernels.  UB = NULL: spectrum of Symmetric(A); otherwise UB is the
 * reverse-Cholesky factor of B and the spectrum is that of the pencil A x = lambda B x.  k steps (k <= 0 or
 * k >= n: n steps = every eigenvalue with its multiplicity); theta receives the Ritz values in ascending order,
 * resid (may be NULL) the bounds |lambda - theta_i| <= resid_i, *k_out their number. */
int pop_eigvals_lanczos(pop_ctx *ctx, const pop_arrowhead *A, const pop_arrowhead *UB, int k, uint64_t seed,
                        double *theta, double *resid, int *k_out);

/* ---- multi-GPU 2D drivers: one process per GPU, X distributed by column blocks ------------------
 * The x<->y switch between the row and column sweeps of examples/adi_poisson.jl:53-56 is a distributed
 * transpose: one kernel transposes tiles through shared memory and stores them into the owning
 * peer's buffer over NVLink (CUDA IPC mappings of library-owned buffers); device-side flags order
 * the exchange.  Usage: every rank calls pop_dist_create, exchanges the handles of
 * pop_dist_get_handle (pop_dist_handle_bytes each, rank order) by any host channel, then
 * pop_dist_open.  Buffers: 0, 1 work panels, 2 F, 3 F^T (n x w_max, leading dimension pop_dist_ld). */
typedef struct pop_dist pop_dist;
int pop_dist_handle_bytes(void);
int pop_dist_create(pop_ctx *ctx, int rank, int world, int64_t n, pop_dist **out);
int pop_dist_get_handle(pop_dist *d, void *handle_out);
int pop_dist_open(pop_dist *d, const void *all_handles);
int pop_dist_destroy(pop_ctx *ctx, pop_dist *d);
double *pop_dist_buffer(pop_dist *d, int which);
int64_t pop_dist_ld(const pop_dist *d);
int64_t pop_dist_lo(const pop_dist *d, int rank); /* first column of `rank` (rank == world: n) */
/* dst_all <- src_all^T across the ranks (collective: every rank calls it in the same order) */
int pop_dist_transpose(pop_ctx *ctx, pop_dist *d, int src_which, int dst_which);
/* ADI of examples/adi_poisson.jl:42-59 on the distributed matrix: dF_local = F[:, lo:hi] in,
 * dX_local = X[:, lo:hi] out (device pointers, column-major). */
int pop_adi_poisson_dist(pop_ctx *ctx, pop_dist *d, const pop_arrowhead *K, const pop_arrowhead *M,
                         const double *dF_local, int64_t ldf, double *dX_local, int64_t ldx, double a,
                         double b, double c, double dd, double tol, int J);

/* nsteps of backward Euler U <- S^{-1} M U M S^{-1}, S = M + dt K, on the distributed n x n field
 * (2D extension of examples/heat.jl:58-63); dU_local = U[:, lo:hi] in place. */
int pop_heat_steps_dist(pop_ctx *ctx, pop_dist *d, const pop_arrowhead *K, const pop_arrowhead *M, double dt,
                        int nsteps, double *dU_local, int64_t ldu);

/* ---- element-decomposed 2D drivers (csrc/pop_row.cu): X distributed over the GPUs by MESH ELEMENTS of
 * the column direction.  Rank g holds the n x nloc panel of the columns of its elements (all degrees)
 * and hats, in arrowhead order (pop_dd_column maps local -> global column).  F \ X and A*X
 * (examples/adi_poisson.jl:55) are local; X / F and X*A (:54) run along rows with lanes over the rows
 * and exchange only the hat-segment maps (O(n*world) doubles per half step) through peer memory.
 * Set-up as pop_dist: create, exchange pop_dd_get_handle blobs (pop_dist_handle_bytes each), open. */
typedef struct pop_dd pop_dd;
int pop_dd_create(pop_ctx *ctx, int rank, int world, int m, int nh, int q, pop_dd **out);
int pop_dd_get_handle(pop_dd *d, void *handle_out);
int pop_dd_open(pop_dd *d, const void *all_handles);
int pop_dd_destroy(pop_ctx *ctx, pop_dd *d);
int64_t pop_dd_nloc(const pop_dd *d);
int64_t pop_dd_ld(const pop_dd *d);
int64_t pop_dd_column(const pop_dd *d, int64_t local_column);
/* rows of dR (n x nloc, leading dimension pop_dd_ld):  r <- T (S^{-1} r) + gamma f, S = U U^T, T = Symmetric(Tm)
 * (Tm may be NULL: plain X / F); collective over the ranks. */
int pop_dd_row_half_step(pop_ctx *ctx, pop_dd *d, const pop_arrowhead *U, const pop_arrowhead *Tm, double gamma,
                         const double *dF, int64_t ldf, double *dR, int64_t ldr);
/* ADI of examples/adi_poisson.jl:42-59 without transposes; dF_local / dX_local are n x nloc panels */
int pop_adi_poisson_dd(pop_ctx *ctx, pop_dd *d, const pop_arrowhead *K, const pop_arrowhead *M, const double *dF_local,
                       int64_t ldf, double *dX_local, int64_t ldx, double a, double b, double c, double dd,
                       double tol, int J);

#ifdef __cplusplus
}
#endif
#endif /* POP_ARROWHEAD_H */
