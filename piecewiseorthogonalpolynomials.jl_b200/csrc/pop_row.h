// Shared by pop_row.cu (element-decomposed row kernels) and pop_row1.cuh (one-pass row kernel).
#pragma once
#include "pop_internal.h"

#define POP_DD_MAXWORLD 16

// ------------------------------------------------------------------------------------
// operator views handed to the row kernels
// ------------------------------------------------------------------------------------
struct RowOp {
    // factor U of the system being solved (upper data, shared D block)
    const double *rd, *w1, *w2;    // [q]
    const double *S;               // [nb][3][m]
    const double *rdA, *offA;      // [nh], [nh-1]
    int nb;
    // product operator T = Symmetric(upper data), shared D block (may be absent: nbT = -1)
    const double *t0, *t1, *t2;    // [q]
    const double *Ts;              // [nbT][3][m]
    const double *TAd, *TAu;
    int nbT;
    int has1;                      // the first super-diagonal of a D block (factor or product) may be non-zero
    double gamma;
    int m, nh, q, k0, ml, hA, nhl, hB, rank, world;
    int64_t nrows, ld, ldf;
    int kb[POP_DD_MAXWORLD + 1];
};


// One-pass row half step on ONE GPU (pop_row1.cuh): rows of R <- T (S^{-1} r) + gamma f with the row block resident in the
// registers of a thread-block cluster.  Returns POP_OK, an error, or POP_ROW1_UNSUPPORTED (caller falls back).
#define POP_ROW1_UNSUPPORTED 1
int pop_row1_half_step(pop_ctx *ctx, const RowOp &op, const double *F, double *R);
// X / F on a row panel (nrhs x n): POP_OK, an error, or POP_ROW1_UNSUPPORTED
int pop_row1_solve_right(pop_ctx *ctx, const pop_arrowhead *U, double *dX, int64_t ldx, int64_t nrhs);

// Single-pass row half step on ONE GPU with TMA tensor traffic and a cluster-wide hat broadcast (pop_row2.cu); same
// contract as pop_row1_half_step.  Tried first; the older kernels are the fallback for the shapes it does not cover.
// rem != null: the row direction is split over several GPUs (RowOp describes this GPU's panel); the CTAs that share a
// row block meet through mailboxes in peer-visible global memory (allocated by pop_dd_create for the plan below).
struct Row2Remote {
    int world, me, csv, mlc, rb;
    int standin;                                 // profiling aid (POP_DD_FAKE_PEERS): one GPU stands in for a rank, nobody answers
    double *mbox[POP_DD_MAXWORLD];               // [2][row blocks][csv readers][world * csv senders][mbw] doubles on every GPU, all bits set when empty
    unsigned long long *merr;                    // this GPU's error slot
    unsigned long long epoch;                    // one per half step, the same on every GPU
    int64_t mbox_doubles;                        // capacity of a mailbox
};
bool pop_row2_remote_plan(int m, int world, int *csv, int *mlc, int *rb, int *mbw);
int pop_row2_half_step(pop_ctx *ctx, const RowOp &op, const double *F, double *R, const Row2Remote *rem);
