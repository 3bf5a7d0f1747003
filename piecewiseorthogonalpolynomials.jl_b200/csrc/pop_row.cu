// Domain-decomposed 2D drivers: the unknown matrix X (n x n) is distributed over the GPUs by MESH
// ELEMENTS of the column direction.  Rank g owns the columns that belong to its elements [k0,k1)
// (all their bubble degrees) and to its hats [hA,hB): a local panel of n rows and
// nloc = nhl + ml*q columns, laid out like an arrowhead vector (hats, then degrees, element fastest).
//   * sweeps along the rows' index (F \ X, A*X of examples/adi_poisson.jl:55) act on whole columns:
//     they are local and use the fused column kernels unchanged;
//   * sweeps along the columns' index (X / F, X*A of examples/adi_poisson.jl:54) act on rows whose
//     entries are spread over the GPUs: the bubble chains of an element are local (lanes run over the
//     rows, so every access is a run of consecutive doubles), only the hat system couples the GPUs.
//     It is solved exactly like the hats of a column inside a thread-block cluster (pop_fused2.cuh):
//     edge couplings to the neighbours, then per direction one segment map per GPU, exchanged through
//     peer memory (CUDA IPC) and ordered by device-side epoch flags.
// Per row half step the GPUs exchange O(n * world) doubles instead of the O(n^2 / world) of a
// distributed transpose.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pop_internal.h"
#include "pop_row.h"

#define POP_DD_MAXQ 128
#define POP_DD_MAXHL 1032   // hats of one rank's segment kept in shared memory by the chain kernels

struct pop_dd {
    int rank, world;
    int m, nh, q;
    int k0, ml, hA, nhl;
    int64_t nrows, ld, nloc;
    int kb[POP_DD_MAXWORLD + 1];      // element bounds of every rank
    // peer-visible slab: [inbA 4*nrows][inbB world*nrows][inbC world*nrows][inbD 2*nrows][flags]
    void *slab;
    size_t slab_bytes;
    void *peer_slab[POP_DD_MAXWORLD];
    unsigned long long epoch;
    bool opened;
    bool fake_peers;   // profiling aid (POP_DD_FAKE_PEERS=1): every peer aliases this rank's own slab, results are meaningless
    // local scratch
    double *hx;      // [nhl+3][nrows]  hats hA-1, my hats, hat hB, one zero row
    double *gdn;     // [nrows] value that entered the downward sweep
    double *xb;      // [4][ml][nrows] x of the coupled bubbles (for the hat rows of the product)
    double *pconst;  // [2][world]
    double *work;    // nrows x nloc work panel (leading dimension ld)
    cudaStream_t pool_stream;   // the stream the local buffers were allocated on (stream-ordered pool)
    bool hats_merged;   // the hat phase runs as ONE kernel (k_row_hats): tile fits, grid co-resident
    int hats_ny;        // lanes over the hats in k_row_hats (block = 32 rows x hats_ny)
    size_t hats_smem;
    bool hats_planned;
    // mailboxes of the single-pass row kernel in remote mode (pop_row2.cu), inside the peer-visible slab
    bool r2_ok;
    int r2_csv, r2_mlc, r2_rb, r2_mbw;
    size_t r2_o_mbox, r2_o_flag;
    int64_t r2_mbox_doubles;
    unsigned long long r2_epoch;
};

struct DDPeers {
    double *inbA[POP_DD_MAXWORLD], *inbB[POP_DD_MAXWORLD], *inbC[POP_DD_MAXWORLD], *inbD[POP_DD_MAXWORLD];
    unsigned long long *flags[POP_DD_MAXWORLD];    // [world] epochs of whole-kernel rounds (+ error slot, + block counters)
    unsigned long long *bflags[POP_DD_MAXWORLD];   // [row blocks of 32][POP_DD_MAXWORLD] epochs of per-row-block rounds
};

#define POP_DD_FLAG_BYTES 256
static void dd_offsets(const pop_dd *d, size_t *oA, size_t *oB, size_t *oC, size_t *oD, size_t *oF) {
    const size_t nr = (size_t)d->nrows;
    *oA = 0;
    *oB = *oA + 4 * nr * 8;
    *oC = *oB + (size_t)d->world * nr * 8;
    *oD = *oC + (size_t)d->world * nr * 8;
    *oF = (*oD + 4 * nr * 8 + 255) & ~(size_t)255;
}

static void dd_peers(const pop_dd *d, DDPeers *p) {
    size_t oA, oB, oC, oD, oF;
    dd_offsets(d, &oA, &oB, &oC, &oD, &oF);
    for (int r = 0; r < d->world; ++r) {
        char *base = (char *)(r == d->rank ? d->slab : d->peer_slab[r]);
        p->inbA[r] = (double *)(base + oA);
        p->inbB[r] = (double *)(base + oB);
        p->inbC[r] = (double *)(base + oC);
        p->inbD[r] = (double *)(base + oD);
        p->flags[r] = (unsigned long long *)(base + oF);
        p->bflags[r] = (unsigned long long *)(base + oF + POP_DD_FLAG_BYTES);
    }
}

extern "C" int pop_dd_create(pop_ctx *ctx, int rank, int world, int m, int nh, int q, pop_dd **out) {
    POP_REQUIRE(ctx && out, POP_ERR_ARG, "pop_dd_create: null argument");
    POP_REQUIRE(world >= 1 && world <= POP_DD_MAXWORLD && rank >= 0 && rank < world, POP_ERR_ARG, "pop_dd_create: bad rank/world");
    POP_REQUIRE(m >= world && q >= 1 && q <= POP_DD_MAXQ && abs(nh - m) <= 1, POP_ERR_DIM, "pop_dd_create: need m >= world, 1 <= q <= %d", POP_DD_MAXQ);
    POP_REQUIRE((m + world - 1) / world + 1 <= POP_DD_MAXHL, POP_ERR_DIM, "pop_dd_create: more than %d hats per rank", POP_DD_MAXHL);
    pop_dd *d = new pop_dd();
    memset(d, 0, sizeof(*d));
    d->rank = rank; d->world = world; d->m = m; d->nh = nh; d->q = q;
    const int base = m / world, rem = m % world;
    for (int r = 0; r <= world; ++r) d->kb[r] = r * base + std::min(r, rem);
    d->k0 = d->kb[rank];
    d->ml = d->kb[rank + 1] - d->k0;
    d->hA = d->k0;
    const int hB = rank == world - 1 ? nh : std::min(nh, d->kb[rank + 1]);
    d->nhl = std::max(hB - d->hA, 0);
    d->nrows = (int64_t)nh + (int64_t)m * q;
    d->ld = (d->nrows + 1) & ~(int64_t)1;
    d->nloc = (int64_t)d->nhl + (int64_t)d->ml * q;
    size_t oA, oB, oC, oD, oF;
    dd_offsets(d, &oA, &oB, &oC, &oD, &oF);
    d->slab_bytes = oF + POP_DD_FLAG_BYTES + (size_t)((d->nrows + 31) / 32) * POP_DD_MAXWORLD * 8;
    {
        // mailboxes [2][row blocks of 8][world * csv][mbw] and flags of k_row2's remote mode (several GPUs; one GPU: POP_ROW2_REMOTE=1)
        const char *e = getenv("POP_ROW2_REMOTE");
        const bool want = world > 1 ? !(e && !atoi(e)) : (e && atoi(e));
        if (want && pop_row2_remote_plan(m, world, &d->r2_csv, &d->r2_mlc, &d->r2_rb, &d->r2_mbw)) {
            const int64_t nblk = (d->nrows + d->r2_rb - 1) / d->r2_rb, nr = (int64_t)world * d->r2_csv;
            d->r2_o_mbox = (d->slab_bytes + 255) & ~(size_t)255;
            d->r2_mbox_doubles = 2 * nblk * d->r2_csv * nr * d->r2_mbw;   // one mailbox per reading CTA
            d->r2_o_flag = d->r2_o_mbox + (size_t)d->r2_mbox_doubles * 8;
            d->slab_bytes = d->r2_o_flag;
            d->r2_ok = true;
        }
    }
    // The slab is what the peers map (CUDA IPC needs a cudaMalloc allocation); everything else is local and comes from the
    // stream-ordered pool, which keeps it between calls: the single-GPU entry points build a decomposition per solve.
    d->pool_stream = ctx->stream;
    cudaError_t e = cudaMalloc(&d->slab, d->slab_bytes);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d->hx, (size_t)(d->nhl + 3) * d->nrows * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d->gdn, (size_t)d->nrows * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d->xb, (size_t)4 * d->ml * d->nrows * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d->pconst, (size_t)2 * POP_DD_MAXWORLD * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d->work, (size_t)d->ld * d->nloc * 8, ctx->stream);
    if (e != cudaSuccess) {
        pop_set_error("pop_dd_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        pop_dd_destroy(ctx, d);   // nothing half-built survives
        return POP_ERR_ALLOC;
    }
    POP_CUDA(cudaMemsetAsync(d->slab, 0, d->slab_bytes, ctx->stream));
    if (d->r2_ok) POP_CUDA(cudaMemsetAsync((char *)d->slab + d->r2_o_mbox, 0xff, (size_t)d->r2_mbox_doubles * 8, ctx->stream));   // all mailbox entries empty
    if (const char *e = getenv("POP_DD_FAKE_PEERS")) {
        if (atoi(e) && world > 1) {
            // one GPU stands in for rank `rank` of `world`: all waits pass at once, all peer stores land locally
            d->fake_peers = true;
            for (int r = 0; r < world; ++r) d->peer_slab[r] = d->slab;
            POP_CUDA(cudaMemsetAsync((char *)d->slab + oF, 0x7f, POP_DD_MAXWORLD * 8, ctx->stream));
            POP_CUDA(cudaMemsetAsync((char *)d->slab + oF + POP_DD_FLAG_BYTES, 0x7f, (d->r2_ok ? d->r2_o_mbox : d->slab_bytes) - oF - POP_DD_FLAG_BYTES, ctx->stream));
            if (d->r2_ok) POP_CUDA(cudaMemsetAsync((char *)d->slab + d->r2_o_mbox, 0, (size_t)d->r2_mbox_doubles * 8, ctx->stream));   // stand-in: every entry reads as arrived
        }
    }
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    d->opened = world == 1 || d->fake_peers;
    *out = d;
    return POP_OK;
}

extern "C" int pop_dd_get_handle(pop_dd *d, void *handle_out) {
    POP_REQUIRE(d && handle_out, POP_ERR_ARG, "pop_dd_get_handle: null argument");
    cudaIpcMemHandle_t h;
    POP_CUDA(cudaIpcGetMemHandle(&h, d->slab));
    memcpy(handle_out, &h, sizeof(h));
    return POP_OK;
}

extern "C" int pop_dd_open(pop_dd *d, const void *all_handles) {
    POP_REQUIRE(d && all_handles, POP_ERR_ARG, "pop_dd_open: null argument");
    for (int r = 0; r < d->world; ++r) {
        if (r == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            pop_set_error("pop_dd_open: cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
            return POP_ERR_CUDA;
        }
        d->peer_slab[r] = p;
    }
    d->opened = true;
    return POP_OK;
}

extern "C" int pop_dd_destroy(pop_ctx *ctx, pop_dd *d) {
    if (!d) return POP_OK;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < d->world; ++r)
        if (!d->fake_peers && r != d->rank && d->peer_slab[r]) cudaIpcCloseMemHandle(d->peer_slab[r]);
    cudaFree(d->slab);
    const cudaStream_t ps = ctx ? ctx->stream : d->pool_stream;
    if (d->hx) cudaFreeAsync(d->hx, ps);
    if (d->gdn) cudaFreeAsync(d->gdn, ps);
    if (d->xb) cudaFreeAsync(d->xb, ps);
    if (d->pconst) cudaFreeAsync(d->pconst, ps);
    if (d->work) cudaFreeAsync(d->work, ps);
    delete d;
    return POP_OK;
}

extern "C" int64_t pop_dd_nloc(const pop_dd *d) { return d ? d->nloc : 0; }
extern "C" int64_t pop_dd_ld(const pop_dd *d) { return d ? d->ld : 0; }
// global column index of local column c (hats first, then degrees, element fastest); -1 if out of range
extern "C" int64_t pop_dd_column(const pop_dd *d, int64_t c) {
    if (!d || c < 0 || c >= d->nloc) return -1;
    if (c < d->nhl) return d->hA + c;
    const int64_t b = c - d->nhl;
    const int64_t j = b / d->ml, kl = b % d->ml;
    return (int64_t)d->nh + j * d->m + d->k0 + kl;
}

// Flags: the epoch protocol of pop_dist.cu, folded into the kernels (no separate signal / wait launches).
// A producing kernel ends with dd_signal_last: the last block to finish tells every rank "my data of epoch
// ep_signal has landed".  A consuming kernel starts with dd_wait_all: every block polls the local flags until
// all ranks have reached ep_wait.  The waited-for data is produced by kernels that never wait on this one, so
// the polling cannot deadlock whatever the residency of the grid.
struct DDSync {
    unsigned long long *peer_flags[POP_DD_MAXWORLD];   // the flag arrays of all ranks (mine included)
    unsigned int *cnt;                                 // block counter of the producing kernel (zero on entry)
    unsigned long long ep_wait, ep_signal;             // 0: no wait / no signal
    int world, me;
};

__device__ __forceinline__ void dd_wait_all(const DDSync &sy) {
    if (sy.ep_wait == 0) return;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t < sy.world) {
        volatile unsigned long long *fl = sy.peer_flags[sy.me];
        long long spins = 0;
        while (fl[t] < sy.ep_wait) {
            __nanosleep(64);
            if (++spins > 100000000LL) {   // a peer died: do not hang the device
                fl[POP_DD_MAXWORLD] = sy.ep_wait;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}

__device__ __forceinline__ void dd_signal_last(const DDSync &sy) {
    if (sy.ep_signal == 0) return;
    __syncthreads();
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    if (t == 0) {
        __threadfence_system();
        const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
        const unsigned prev = atomicAdd(sy.cnt, 1u);
        if (prev == nblocks - 1) {
            *sy.cnt = 0;
            __threadfence_system();
            for (int r = 0; r < sy.world; ++r) {
                volatile unsigned long long *p = sy.peer_flags[r] + sy.me;
                *p = sy.ep_signal;
            }
            __threadfence_system();
        }
    }
}

// multipliers of every rank's hat segment: Pdn[g] = prod(-off[i] rd[i]), Pup[g] = prod(-off[i-1] rd[i]) over its hats
// One block per factor: ptrs[2f] = rdA, ptrs[2f+1] = offA of factor f; pc + f*2*POP_DD_MAXWORLD receives its multipliers.
struct DDBounds {
    int kb[POP_DD_MAXWORLD + 1];
    int world, nh;
};
__global__ void k_dd_pconst(DDBounds bd, const double *const *__restrict__ ptrs, const double *rdA1, const double *offA1, double *pc) {
    const int g = threadIdx.x;
    if (g >= bd.world) return;
    const double *rdA = ptrs ? ptrs[2 * blockIdx.x] : rdA1, *offA = ptrs ? ptrs[2 * blockIdx.x + 1] : offA1;
    pc += (size_t)blockIdx.x * 2 * POP_DD_MAXWORLD;
    const int a = bd.kb[g], b = g == bd.world - 1 ? bd.nh : min(bd.nh, bd.kb[g + 1]);
    double pd = 1.0, pu = 1.0;
    for (int i = a; i < b; ++i) {
        const double rd = rdA[i];
        pd *= -(i < bd.nh - 1 ? offA[i] : 0.0) * rd;
        pu *= -(i > 0 ? offA[i - 1] : 0.0) * rd;
    }
    pc[g] = pd;
    pc[POP_DD_MAXWORLD + g] = pu;
}

// ---- K1: backward bubble sweep along the rows:  z_j = (b_j - w1_j z_{j+1} - w2_j z_{j+2}) rd_j ------------
// thread <-> (row r, local element kl); the 32 lanes of a warp are 32 consecutive rows.
template <int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB) k_row_bwd(RowOp op, double *__restrict__ X, DDPeers peers, DDSync sy) {
    __shared__ double srd[POP_DD_MAXQ], sw1[POP_DD_MAXQ], sw2[POP_DD_MAXQ];
    const int q = op.q;
    for (int j = threadIdx.y * 32 + threadIdx.x; j < q; j += 256) {
        srd[j] = op.rd[j];
        sw1[j] = (op.w1 && j + 1 < q) ? op.w1[j] : 0.0;
        sw2[j] = (op.w2 && j + 2 < q) ? op.w2[j] : 0.0;
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int kl = blockIdx.y * 8 + threadIdx.y;
    if (r < op.nrows && kl < op.ml) {
        double *p = X + r + ((int64_t)op.nhl + kl) * op.ld;
        const int64_t js = (int64_t)op.ml * op.ld;
        double z1 = 0.0, z2 = 0.0;
        int j = q - 1;
        for (; j >= DEPTH - 1; j -= DEPTH) {
            double v[DEPTH];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) v[u] = p[(int64_t)(j - u) * js];
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) {
                const double t = (v[u] - sw1[j - u] * z1 - sw2[j - u] * z2) * srd[j - u];
                z2 = z1;
                z1 = t;
                v[u] = t;
            }
#pragma unroll
            for (int u = 0; u < DEPTH; ++u) p[(int64_t)(j - u) * js] = v[u];
        }
        for (; j >= 0; --j) {
            const double t = (p[(int64_t)j * js] - sw1[j] * z1 - sw2[j] * z2) * srd[j];
            p[(int64_t)j * js] = t;
            z2 = z1;
            z1 = t;
        }
        // couplings of my edge elements to the neighbours' hats (the thread re-reads its own coupled bubbles)
        const int m = op.m;
        if (kl == 0 && op.rank > 0) {   // element k0, slot 0 -> hat k0-1 (last hat of the left neighbour)
            double acc = 0.0;
            for (int b = 0; b < op.nb; ++b) acc += op.S[((size_t)b * 3 + 0) * m + op.k0] * p[(int64_t)b * js];
            peers.inbA[op.rank - 1][op.nrows + r] = acc;
        }
        if (kl == op.ml - 1 && op.rank < op.world - 1) {   // element k1-1, slot 2 -> hat k1 (first hat of the right neighbour)
            double acc = 0.0;
            for (int b = 0; b < op.nb; ++b) acc += op.S[((size_t)b * 3 + 2) * m + op.k0 + kl] * p[(int64_t)b * js];
            peers.inbA[op.rank + 1][r] = acc;
        }
    }
    dd_signal_last(sy);
}

// ---- K2a: hat right-hand sides of my segment: one thread per (row, hat), fully parallel ------------------------
__global__ void __launch_bounds__(256) k_row_hat_gather(RowOp op, double *__restrict__ X, DDPeers peers, DDSync sy) {
    dd_wait_all(sy);
    const int64_t r = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int u = blockIdx.y * 8 + threadIdx.y;
    if (r >= op.nrows || u >= op.nhl) return;
    const int m = op.m, ml = op.ml, nhl = op.nhl, k0 = op.k0;
    const int i = op.hA + u;
    double v = X[r + (int64_t)u * op.ld];
    for (int b = 0; b < op.nb; ++b)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int kl = i - s + 1 - k0;
            if (kl >= 0 && kl < ml) v -= op.S[((size_t)b * 3 + s) * m + k0 + kl] * X[r + ((int64_t)nhl + (int64_t)b * ml + kl) * op.ld];
        }
    const double *inA = peers.inbA[op.rank];
    if (u == 0 && op.rank > 0) v -= inA[r];                                  // element k0-1 (slot 2) of the left neighbour
    if (i == k0 + ml - 1 && op.rank < op.world - 1) v -= inA[op.nrows + r];  // element k1 (slot 0) of the right neighbour
    X[r + (int64_t)u * op.ld] = v;
}

// The hat recurrences: one thread per row, sequential in the hat index, every access a run of consecutive
// doubles across the lanes.  Loads are issued 8 hats ahead of the dependent chain; the hat factor of the
// segment sits in shared memory.  c_dn[u] = off(u,u+1) rd[u], c_up[u] = off(u-1,u) rd[u].
#define DD_HAT_TABLES()                                                                                   \
    __shared__ double s_rd[POP_DD_MAXHL], s_cdn[POP_DD_MAXHL], s_cup[POP_DD_MAXHL];                       \
    for (int u = threadIdx.x; u < op.nhl; u += blockDim.x) {                                              \
        const int i = op.hA + u;                                                                          \
        const double rdv = op.rdA[i];                                                                     \
        s_rd[u] = rdv;                                                                                    \
        s_cdn[u] = (i < op.nh - 1 ? op.offA[i] : 0.0) * rdv;                                              \
        s_cup[u] = (i > 0 ? op.offA[i - 1] : 0.0) * rdv;                                                  \
    }                                                                                                     \
    (void)s_cdn;                                                                                          \
    (void)s_cup;                                                                                          \
    __syncthreads();

template <bool DOWN, bool STORE>
__device__ __forceinline__ double dd_chain(double *__restrict__ xr, const int64_t ld, const int nhl, const double *s_rd, const double *s_c, double hin) {
    for (int b0 = 0; b0 < nhl; b0 += 8) {
        double v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int u = DOWN ? nhl - 1 - (b0 + s) : b0 + s;
            v[s] = (b0 + s < nhl) ? xr[(int64_t)u * ld] : 0.0;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int u = DOWN ? nhl - 1 - (b0 + s) : b0 + s;
            if (b0 + s < nhl) {
                hin = fma(-s_c[u], hin, v[s] * s_rd[u]);
                if (STORE) xr[(int64_t)u * ld] = hin;
            }
        }
    }
    return hin;
}

// ---- K2b: first pass of the downward sweep (zero incoming): my segment's map for the ranks below me --------------
__global__ void __launch_bounds__(128) k_row_hat_dn1(RowOp op, double *__restrict__ X, DDPeers peers, DDSync sy) {
    DD_HAT_TABLES()
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < op.nrows) {
        const double t = dd_chain<true, false>(X + r, op.ld, op.nhl, s_rd, s_cdn, 0.0);
        for (int g = 0; g < op.rank; ++g) peers.inbB[g][(size_t)op.rank * op.nrows + r] = t;
    }
    dd_signal_last(sy);
}

// ---- K2c: downward sweep with the exact incoming value, first pass of the upward sweep -----------------------
__global__ void __launch_bounds__(128) k_row_hat_b(RowOp op, double *__restrict__ X, DDPeers peers, const double *__restrict__ pc, double *__restrict__ gdn,
                                                   DDSync sy) {
    DD_HAT_TABLES()
    dd_wait_all(sy);
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < op.nrows) {
        const double *inB = peers.inbB[op.rank];
        double G = 0.0;
        for (int g = op.world - 1; g > op.rank; --g) G = fma(pc[g], G, inB[(size_t)g * op.nrows + r]);
        gdn[r] = G;
        dd_chain<true, true>(X + r, op.ld, op.nhl, s_rd, s_cdn, G);
        if (op.world > 1) {
            const double t = dd_chain<false, false>(X + r, op.ld, op.nhl, s_rd, s_cup, 0.0);
            for (int g = op.rank + 1; g < op.world; ++g) peers.inbC[g][(size_t)op.rank * op.nrows + r] = t;
        }
    }
    dd_signal_last(sy);
}

// ---- K2d: upward sweep with the exact incoming value; hx = my hats with the neighbours' hats hA-1 and hB around them ----
// hx[(u+1)*nrows + r] = hat hA+u, hx[r] = hat hA-1, hx[(nhl+1)*nrows + r] = hat hB, one more zero row behind
__global__ void __launch_bounds__(128) k_row_hat_c(RowOp op, double *__restrict__ X, DDPeers peers, const double *__restrict__ pc, const double *__restrict__ gdn,
                                                   double *__restrict__ hx, DDSync sy) {
    DD_HAT_TABLES()
    dd_wait_all(sy);
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= op.nrows) return;
    const double *inC = peers.inbC[op.rank];
    double G = 0.0;
    for (int g = 0; g < op.rank; ++g) G = fma(pc[POP_DD_MAXWORLD + g], G, inC[(size_t)g * op.nrows + r]);
    double hin = G;
    const int nhl = op.nhl;
    for (int b0 = 0; b0 < nhl; b0 += 8) {
        double v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) v[s] = (b0 + s < nhl) ? X[r + (int64_t)(b0 + s) * op.ld] : 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int u = b0 + s;
            if (u < nhl) {
                hin = fma(-s_cup[u], hin, v[s] * s_rd[u]);
                X[r + (int64_t)u * op.ld] = hin;
                hx[(size_t)(u + 1) * op.nrows + r] = hin;
            }
        }
    }
    hx[r] = op.rank > 0 ? G : 0.0;                                                           // hat hA-1
    // hat hB = (y[hB] - off[hB-1] h[hB-1]) rd[hB], y[hB] being the value that entered my downward sweep
    hx[(size_t)(nhl + 1) * op.nrows + r] = (op.hB < op.nh && nhl > 0) ? (gdn[r] - op.offA[op.hB - 1] * hin) * op.rdA[op.hB] : 0.0;
    hx[(size_t)(nhl + 2) * op.nrows + r] = 0.0;
}

// ---- K2 (merged): the whole hat phase of 32 rows in one block ------------------------------------------------
// Block = 32 rows x ny lanes of hats (ny = 16: the gather is latency-bound, see profiles/r1_row1.txt).  The hat right-hand
// sides of the 32 rows are gathered into a shared-memory tile [nhl][32] by all threads; the recurrences then run on the tile (lane x of warp 0 owns row x: no global
// latency inside the dependent chains); the two segment-map exchanges use flags PER ROW BLOCK, so a block only
// waits for the same rows of its peers; the result leaves through all 128 threads again.  With peers the grid
// must be co-resident (checked by the host; the per-kernel path above is the fallback).
__device__ __forceinline__ void dd_block_exchange(const DDPeers &peers, int world, int me, unsigned long long ep) {
    __syncthreads();   // the block's peer stores are issued
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const size_t slot = (size_t)blockIdx.x * POP_DD_MAXWORLD;
    if (t == 0) {
        __threadfence_system();
        for (int r = 0; r < world; ++r) {
            volatile unsigned long long *p = peers.bflags[r] + slot + me;
            *p = ep;
        }
    }
    if (t < world) {
        volatile unsigned long long *fl = peers.bflags[me] + slot;
        long long spins = 0;
        while (fl[t] < ep) {
            if (++spins > 400000000LL) {   // a peer died: do not hang the device
                volatile unsigned long long *err = peers.flags[me];
                err[POP_DD_MAXWORLD] = ep;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_row_hats(RowOp op, double *__restrict__ X, DDPeers peers, const double *__restrict__ pc, double *__restrict__ hx,
                                                  DDSync sy, unsigned long long epB, unsigned long long epC) {
    extern __shared__ double sm_h[];
    const int nhl = op.nhl, nhlp = (nhl + 1) & ~1;
    double *s_rd = sm_h, *s_cdn = s_rd + nhlp, *s_cup = s_cdn + nhlp, *tile = s_cup + nhlp;
    const int x = threadIdx.x, y = threadIdx.y, t = y * 32 + x;
    const int ny = blockDim.y, nthr = 32 * ny;   // 32 rows x ny lanes over the hats
    for (int u = t; u < nhl; u += nthr) {
        const int i = op.hA + u;
        const double rdv = op.rdA[i];
        s_rd[u] = rdv;
        s_cdn[u] = (i < op.nh - 1 ? op.offA[i] : 0.0) * rdv;
        s_cup[u] = (i > 0 ? op.offA[i - 1] : 0.0) * rdv;
    }
    dd_wait_all(sy);   // the neighbours' edge couplings have landed
    const int64_t r = (int64_t)blockIdx.x * 32 + x;
    const bool rv = r < op.nrows;
    const int m = op.m, ml = op.ml, k0 = op.k0, me = op.rank, world = op.world;
    const bool multi = world > 1;
    // gather: hat right-hand sides minus the couplings to the (backward-swept) bubbles
    if (rv) {
        const double *inA = peers.inbA[me];
        for (int u = y; u < nhl; u += ny) {
            const int i = op.hA + u;
            double v = X[r + (int64_t)u * op.ld];
            for (int b = 0; b < op.nb; ++b)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int kl = i - s + 1 - k0;
                    if (kl >= 0 && kl < ml) v -= op.S[((size_t)b * 3 + s) * m + k0 + kl] * X[r + ((int64_t)nhl + (int64_t)b * ml + kl) * op.ld];
                }
            if (u == 0 && me > 0) v -= __ldcg(inA + r);
            if (i == k0 + ml - 1 && me < world - 1) v -= __ldcg(inA + op.nrows + r);
            tile[u * 32 + x] = v;
        }
    }
    __syncthreads();
    double *col = tile + x;
    double Gdn = 0.0, Gup = 0.0;
    if (multi) {
        if (y == 0 && rv) {   // my segment's downward map with zero incoming, for the ranks below me
            double h = 0.0;
            for (int u = nhl - 1; u >= 0; --u) h = fma(-s_cdn[u], h, col[u * 32] * s_rd[u]);
            for (int g = 0; g < me; ++g) peers.inbB[g][(size_t)me * op.nrows + r] = h;
        }
        dd_block_exchange(peers, world, me, epB);
        if (y == 0 && rv) {
            const double *inB = peers.inbB[me];
            for (int g = world - 1; g > me; --g) Gdn = fma(pc[g], Gdn, __ldcg(inB + (size_t)g * op.nrows + r));
        }
    }
    if (y == 0 && rv) {
        double h = Gdn;
        for (int u = nhl - 1; u >= 0; --u) {
            h = fma(-s_cdn[u], h, col[u * 32] * s_rd[u]);
            col[u * 32] = h;
        }
        if (multi) {
            h = 0.0;
            for (int u = 0; u < nhl; ++u) h = fma(-s_cup[u], h, col[u * 32] * s_rd[u]);
            for (int g = me + 1; g < world; ++g) peers.inbC[g][(size_t)me * op.nrows + r] = h;
        }
    }
    if (multi) {
        dd_block_exchange(peers, world, me, epC);
        if (y == 0 && rv) {
            const double *inC = peers.inbC[me];
            for (int g = 0; g < me; ++g) Gup = fma(pc[POP_DD_MAXWORLD + g], Gup, __ldcg(inC + (size_t)g * op.nrows + r));
        }
    }
    if (y == 0 && rv) {
        double h = Gup;
        for (int u = 0; u < nhl; ++u) {
            h = fma(-s_cup[u], h, col[u * 32] * s_rd[u]);
            col[u * 32] = h;
        }
        hx[r] = me > 0 ? Gup : 0.0;   // hat hA-1
        // hat hB = (y[hB] - off[hB-1] h[hB-1]) rd[hB], y[hB] being the value that entered my downward sweep
        hx[(size_t)(nhl + 1) * op.nrows + r] = (op.hB < op.nh && nhl > 0) ? (Gdn - op.offA[op.hB - 1] * h) * op.rdA[op.hB] : 0.0;
        hx[(size_t)(nhl + 2) * op.nrows + r] = 0.0;
    }
    __syncthreads();
    if (rv)
        for (int u = y; u < nhl; u += ny) {
            const double h = tile[u * 32 + x];
            X[r + (int64_t)u * op.ld] = h;
            hx[(size_t)(u + 1) * op.nrows + r] = h;
        }
}

// ---- K3: correction of the coupled bubbles, forward sweep, product with T, + gamma F; in place -----------------
//   x_j = (z_j - w1_{j-1} x_{j-1} - w2_{j-2} x_{j-2}) rd_j ;  r_j = sum_d T(j,j+d) x_{j+d} + coupling to the hats + gamma f_j
// Row j-2 of the product is emitted once x_j is known (window of five x values).  Blocks of 8 degrees: the 16
// loads of a block (z and f) are issued before its dependent chain.
template <int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB) k_row_fwd_mul(RowOp op, double *__restrict__ X, const double *__restrict__ F, const double *__restrict__ hx,
                                                     double *__restrict__ xb, DDPeers peers, DDSync sy) {
    __shared__ double srd[POP_DD_MAXQ + 12], sw1[POP_DD_MAXQ + 12], sw2[POP_DD_MAXQ + 12], st0[POP_DD_MAXQ + 12], st1[POP_DD_MAXQ + 12], st2[POP_DD_MAXQ + 12];
    const int q = op.q;
    const bool mul = op.nbT >= 0;
    // tables shifted by 2 so that index j-2 .. j+2 never needs a guard (zero outside [0,q))
    for (int j = threadIdx.y * 32 + threadIdx.x; j < q + 12; j += 256) {
        const int jj = j - 2;
        const bool in = jj >= 0 && jj < q;
        srd[j] = in ? op.rd[jj] : 0.0;
        sw1[j] = (in && op.w1 && jj + 1 < q) ? op.w1[jj] : 0.0;
        sw2[j] = (in && op.w2 && jj + 2 < q) ? op.w2[jj] : 0.0;
        st0[j] = (in && mul) ? op.t0[jj] : 0.0;
        st1[j] = (in && mul && op.t1 && jj + 1 < q) ? op.t1[jj] : 0.0;
        st2[j] = (in && mul && op.t2 && jj + 2 < q) ? op.t2[jj] : 0.0;
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int kl = blockIdx.y * 8 + threadIdx.y;
    if (r < op.nrows && kl < op.ml) {
        const int m = op.m, ml = op.ml, nhl = op.nhl, k = op.k0 + kl;
        const int64_t ld = op.ld, js = (int64_t)ml * ld;
        double *p = X + r + ((int64_t)nhl + kl) * ld;
        const double *f = F ? F + r + ((int64_t)nhl + kl) * op.ldf : nullptr;
        const int64_t fs = (int64_t)ml * op.ldf;
        // hats k-1, k, k+1 of this element
        const double h0 = hx[(size_t)kl * op.nrows + r], h1 = hx[(size_t)(kl + 1) * op.nrows + r], h2 = hx[(size_t)(kl + 2) * op.nrows + r];
        double xw0 = 0.0, xw1 = 0.0, xw2 = 0.0, xw3 = 0.0, xw4 = 0.0;   // x_j, x_{j-1}, ..., x_{j-4}
        double eL = 0.0, eR = 0.0;                                       // couplings of my edge element to a neighbour's product hat
        for (int j0 = 0; j0 < q + 2; j0 += DEPTH) {
            double zv[DEPTH], fv[DEPTH];
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) {
                const int j = j0 + s;
                zv[s] = j < q ? p[(int64_t)j * js] : 0.0;
                const int jo = j - 2;
                fv[s] = (f && jo >= 0 && jo < q) ? f[(int64_t)jo * fs] : 0.0;
            }
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) {
                const int j = j0 + s;
                double v = 0.0;
                if (j < q) {
                    v = zv[s];
                    if (j < op.nb) {
                        const double *sb = op.S + (size_t)j * 3 * m + k;
                        v -= sb[0] * h0 + sb[m] * h1 + sb[2 * m] * h2;
                    }
                    v = (v - sw1[j + 1] * xw0 - sw2[j] * xw1) * srd[j + 2];   // w1_{j-1} x_{j-1}, w2_{j-2} x_{j-2}
                    if (!mul) p[(int64_t)j * js] = v;
                    else if (j < op.nbT) {
                        xb[((size_t)j * ml + kl) * op.nrows + r] = v;
                        const double *tb = op.Ts + (size_t)j * 3 * m + k;
                        eL += tb[0] * v;
                        eR += tb[2 * m] * v;
                    }
                }
                xw4 = xw3; xw3 = xw2; xw2 = xw1; xw1 = xw0; xw0 = v;
                const int jo = j - 2;
                if (mul && jo >= 0 && jo < q) {
                    // T(jo,jo) x_jo + T(jo,jo+1) x_{jo+1} + T(jo-1,jo) x_{jo-1} + T(jo,jo+2) x_{jo+2} + T(jo-2,jo) x_{jo-2}
                    double acc = st0[jo + 2] * xw2 + st1[jo + 2] * xw1 + st1[jo + 1] * xw3 + st2[jo + 2] * xw0 + st2[jo] * xw4;
                    if (jo < op.nbT) {
                        const double *tb = op.Ts + (size_t)jo * 3 * m + k;
                        acc += tb[0] * h0 + tb[m] * h1 + tb[2 * m] * h2;
                    }
                    p[(int64_t)jo * js] = fma(op.gamma, fv[s], acc);
                }
            }
        }
        if (mul && op.world > 1) {
            if (kl == 0 && op.rank > 0) peers.inbD[op.rank - 1][op.nrows + r] = eL;            // element k0 -> hat k0-1
            if (kl == ml - 1 && op.rank < op.world - 1) peers.inbD[op.rank + 1][r] = eR;      // element k1-1 -> hat k1
        }
    }
    dd_signal_last(sy);
}

// ---- K4: hat rows of the product: one thread per (row, hat), reads the hats from hx, writes X -----------------------
__global__ void __launch_bounds__(256) k_row_hat_mul(RowOp op, double *__restrict__ X, const double *__restrict__ F, const double *__restrict__ hx,
                                                     const double *__restrict__ xb, DDPeers peers, DDSync sy) {
    dd_wait_all(sy);
    const int64_t r = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int u = blockIdx.y * 8 + threadIdx.y;
    if (r >= op.nrows || u >= op.nhl) return;
    const int m = op.m, ml = op.ml, k0 = op.k0;
    const int i = op.hA + u;
    const double hp = hx[(size_t)u * op.nrows + r], hc = hx[(size_t)(u + 1) * op.nrows + r], hn = hx[(size_t)(u + 2) * op.nrows + r];
    double acc = op.TAd[i] * hc;
    if (i + 1 < op.nh) acc += op.TAu[i] * hn;
    if (i > 0) acc += op.TAu[i - 1] * hp;
    for (int b = 0; b < op.nbT; ++b)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int kl = i - s + 1 - k0;
            if (kl >= 0 && kl < ml) acc += op.Ts[((size_t)b * 3 + s) * m + k0 + kl] * xb[((size_t)b * ml + kl) * op.nrows + r];
        }
    const double *inD = peers.inbD[op.rank];
    if (u == 0 && op.rank > 0) acc += inD[r];
    if (i == k0 + ml - 1 && op.rank < op.world - 1) acc += inD[op.nrows + r];
    X[r + (int64_t)u * op.ld] = F ? fma(op.gamma, F[r + (int64_t)u * op.ldf], acc) : acc;
}

// ------------------------------------------------------------------------------------
// one row half step:  R <- (R U^{-T}... ) i.e. rows of R:  r <- T (S^{-1} r) + gamma f   with S = U U^T
// ------------------------------------------------------------------------------------
static int dd_make_op(const pop_dd *d, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, int64_t ldf, RowOp *op) {
    const pop_desc &u = U->d;
    POP_REQUIRE(U->is_factor && U->mD == 1 && u.uD <= 2 && u.m == d->m && u.nh == d->nh && u.q == d->q, POP_ERR_DIM,
                "row half step: factor does not match the decomposition (or is not a shared-block factor with <= 2 bands)");
    memset(op, 0, sizeof(*op));
    const size_t plane = (size_t)u.q;
    op->rd = U->p.rdD;
    op->w1 = u.uD >= 1 ? U->p.D + (size_t)(u.lD + 1) * plane : nullptr;
    op->w2 = u.uD >= 2 ? U->p.D + (size_t)(u.lD + 2) * plane : nullptr;
    op->S = U->p.B;
    op->rdA = U->p.rdA;
    op->offA = U->p.Au;
    op->nb = std::min(u.nB, u.q);
    op->nbT = -1;
    if (T) {
        const pop_desc &t = T->d;
        POP_REQUIRE(T->mD == 1 && t.uD <= 2 && t.m == d->m && t.nh == d->nh && t.q == d->q, POP_ERR_DIM, "row half step: product operator does not match");
        const size_t tpl = (size_t)t.q;
        op->t0 = T->p.D + (size_t)t.lD * tpl;
        op->t1 = t.uD >= 1 ? T->p.D + (size_t)(t.lD + 1) * tpl : nullptr;
        op->t2 = t.uD >= 2 ? T->p.D + (size_t)(t.lD + 2) * tpl : nullptr;
        op->Ts = T->p.B;
        op->TAd = T->p.Ad;
        op->TAu = T->p.Au;
        op->nbT = std::min(t.nB, t.q);
        POP_REQUIRE(op->nbT <= 4, POP_ERR_DIM, "row half step: at most 4 coupling blocks");
    }
    op->has1 = ((u.uD >= 1 && !U->band1_zero) || (T && T->d.uD >= 1 && !T->band1_zero)) ? 1 : 0;
    op->gamma = gamma;
    op->m = d->m; op->nh = d->nh; op->q = d->q; op->k0 = d->k0; op->ml = d->ml; op->hA = d->hA; op->nhl = d->nhl;
    op->hB = d->hA + d->nhl;
    op->rank = d->rank; op->world = d->world;
    op->nrows = d->nrows; op->ld = d->ld; op->ldf = ldf;
    for (int r = 0; r <= d->world; ++r) op->kb[r] = d->kb[r];
    return POP_OK;
}

// one kernel for the hat phase if its tile fits in shared memory and (with peers) the whole grid is co-resident
static void dd_plan_hats(pop_ctx *ctx, pop_dd *d) {
    d->hats_planned = true;
    d->hats_merged = false;
    if (const char *e = getenv("POP_DD_SPLIT_HATS"))
        if (atoi(e)) return;
    const int nhlp = (d->nhl + 1) & ~1;
    const size_t smem = ((size_t)3 * nhlp + (size_t)d->nhl * 32) * 8;
    if (smem > (size_t)ctx->max_smem_optin || smem > 160 * 1024) return;
    if (cudaFuncSetAttribute(k_row_hats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    d->hats_ny = 16;
    if (const char *e = getenv("POP_DD_HATS_NY")) d->hats_ny = std::min(std::max(atoi(e), 1), 32);
    if (d->world > 1) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_row_hats, 32 * d->hats_ny, smem) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        if ((d->nrows + 31) / 32 > (int64_t)occ * ctx->sm_count) return;
    }
    d->hats_smem = smem;
    d->hats_merged = true;
}

// rows of R (nrows x nloc, leading dimension d->ld):  r <- T (S^{-1} r) + gamma f ; T may be null (plain solve)
static int dd_row_half_step(pop_ctx *ctx, pop_dd *d, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *F, int64_t ldf,
                            double *R, const double *pc) {
    RowOp op;
    int rc = dd_make_op(d, U, T, gamma, ldf, &op);
    if (rc) return rc;
    if (d->r2_ok) {
        // single pass with the hat exchange through mailboxes in peer memory (several GPUs, or forced on one)
        Row2Remote rem;
        memset(&rem, 0, sizeof(rem));
        rem.world = d->world; rem.me = d->rank; rem.csv = d->r2_csv; rem.mlc = d->r2_mlc; rem.rb = d->r2_rb;
        rem.standin = d->fake_peers ? 1 : 0;
        for (int r = 0; r < d->world; ++r) {
            char *base = (char *)(r == d->rank ? d->slab : d->peer_slab[r]);
            rem.mbox[r] = (double *)(base + d->r2_o_mbox);
        }
        size_t oA, oB, oC, oD, oF;
        dd_offsets(d, &oA, &oB, &oC, &oD, &oF);
        rem.merr = (unsigned long long *)((char *)d->slab + oF) + POP_DD_MAXWORLD;
        rem.epoch = d->r2_epoch + 1;
        rem.mbox_doubles = d->r2_mbox_doubles;
        rc = pop_row2_half_step(ctx, op, (T && gamma != 0.0) ? F : nullptr, R, &rem);
        if (rc == POP_OK) { ++d->r2_epoch; return rc; }
        if (rc != POP_ROW1_UNSUPPORTED) return rc;
    }
    if (d->world == 1) {
        // one GPU: the row block stays on chip for the whole half step (pop_row1.cuh), one pass over HBM
        rc = pop_row2_half_step(ctx, op, (T && gamma != 0.0) ? F : nullptr, R, nullptr);
        if (rc != POP_ROW1_UNSUPPORTED) return rc;
        rc = pop_row1_half_step(ctx, op, (T && gamma != 0.0) ? F : nullptr, R);
        if (rc != POP_ROW1_UNSUPPORTED) return rc;
    }
    DDPeers peers;
    dd_peers(d, &peers);
    const bool multi = d->world > 1;
    if (!d->hats_planned) dd_plan_hats(ctx, d);
    // epochs of the four exchange rounds: A edge couplings, B downward maps, C upward maps, D product edges
    const unsigned long long epA = multi ? ++d->epoch : 0, epB = multi ? ++d->epoch : 0, epC = multi ? ++d->epoch : 0,
                             epD = (multi && T) ? ++d->epoch : 0;
    size_t oA, oB, oC, oD, oF;
    dd_offsets(d, &oA, &oB, &oC, &oD, &oF);
    unsigned int *cnt = (unsigned int *)((char *)d->slab + oF + 192);
    auto sync = [&](unsigned long long wait, unsigned long long signal, int slot) {
        DDSync sy;
        for (int r = 0; r < POP_DD_MAXWORLD; ++r) sy.peer_flags[r] = r < d->world ? peers.flags[r] : nullptr;
        sy.cnt = cnt + slot;
        sy.ep_wait = wait;
        sy.ep_signal = signal;
        sy.world = d->world;
        sy.me = d->rank;
        return sy;
    };
    const dim3 blk2(32, 8), grd2((unsigned)((d->nrows + 31) / 32), (unsigned)((d->ml + 7) / 8));
    const dim3 grdh((unsigned)((d->nrows + 31) / 32), (unsigned)std::max((d->nhl + 7) / 8, 1));
    const unsigned grd1 = (unsigned)((d->nrows + 127) / 128);
    if (!pc) {
        DDBounds bd;
        for (int r = 0; r <= d->world; ++r) bd.kb[r] = d->kb[r];
        bd.world = d->world; bd.nh = d->nh;
        k_dd_pconst<<<1, 32, 0, ctx->stream>>>(bd, nullptr, op.rdA, op.offA, d->pconst);
        ctx->launches++;
        pc = d->pconst;
    }
    {
        static int bv = -1;
        if (bv < 0) {
            bv = 4;   // measured (profiles/r1_row1.txt): blocks of 4 degrees at 40 registers, 6 CTAs/SM
            if (const char *e = getenv("POP_DD_BWD_VARIANT")) bv = atoi(e);
        }
        const DDSync sy = sync(0, epA, 0);
        switch (bv) {
            case 1: k_row_bwd<4, 4><<<grd2, blk2, 0, ctx->stream>>>(op, R, peers, sy); break;
            case 2: k_row_bwd<8, 3><<<grd2, blk2, 0, ctx->stream>>>(op, R, peers, sy); break;
            case 3: k_row_bwd<16, 2><<<grd2, blk2, 0, ctx->stream>>>(op, R, peers, sy); break;
            case 4: k_row_bwd<4, 6><<<grd2, blk2, 0, ctx->stream>>>(op, R, peers, sy); break;
            default: k_row_bwd<8, 1><<<grd2, blk2, 0, ctx->stream>>>(op, R, peers, sy); break;
        }
    }
    ctx->launches++;
    if (d->hats_merged) {
        k_row_hats<<<(unsigned)((d->nrows + 31) / 32), dim3(32, d->hats_ny), d->hats_smem, ctx->stream>>>(op, R, peers, pc, d->hx, sync(epA, 0, 0), epB, epC);
        ctx->launches++;
    } else {
        k_row_hat_gather<<<grdh, blk2, 0, ctx->stream>>>(op, R, peers, sync(epA, 0, 0));
        ctx->launches++;
        if (multi) {
            k_row_hat_dn1<<<grd1, 128, 0, ctx->stream>>>(op, R, peers, sync(0, epB, 1));
            ctx->launches++;
        }
        k_row_hat_b<<<grd1, 128, 0, ctx->stream>>>(op, R, peers, pc, d->gdn, sync(epB, epC, 2));
        k_row_hat_c<<<grd1, 128, 0, ctx->stream>>>(op, R, peers, pc, d->gdn, d->hx, sync(epC, 0, 0));
        ctx->launches += 2;
    }
    {
        static int variant = -1;
        if (variant < 0) {
            variant = 0;
            if (const char *e = getenv("POP_DD_FWD_VARIANT")) variant = atoi(e);
        }
        const double *Fp = (T && gamma != 0.0) ? F : nullptr;
        const DDSync sy = sync(0, epD, 3);
        // measured (profiles/r1_dd_standin.txt): blocks of 4 degrees at 64 registers (4 CTAs/SM) edge out blocks of 8 at 128
        switch (variant) {
            case 1: k_row_fwd_mul<8, 2><<<grd2, blk2, 0, ctx->stream>>>(op, R, Fp, d->hx, d->xb, peers, sy); break;
            case 2: k_row_fwd_mul<8, 4><<<grd2, blk2, 0, ctx->stream>>>(op, R, Fp, d->hx, d->xb, peers, sy); break;
            default: k_row_fwd_mul<4, 4><<<grd2, blk2, 0, ctx->stream>>>(op, R, Fp, d->hx, d->xb, peers, sy); break;
        }
    }
    ctx->launches++;
    if (T) {
        k_row_hat_mul<<<grdh, blk2, 0, ctx->stream>>>(op, R, gamma != 0.0 ? F : nullptr, d->hx, d->xb, peers, sync(epD, 0, 0));
        ctx->launches++;
    }
    POP_CUDA(cudaGetLastError());
    return POP_OK;
}

// After the final stream synchronisation of a driver: did a bounded spin-wait on a peer give up?  (dd_wait_all,
// dd_block_exchange record the epoch in the error slot and carry on with incomplete data.)
static int dd_check_timeout(pop_ctx *ctx, pop_dd *d, const char *who) {
    if (d->world == 1 || d->fake_peers) return POP_OK;
    size_t oA, oB, oC, oD, oF;
    dd_offsets(d, &oA, &oB, &oC, &oD, &oF);
    unsigned long long *slot = (unsigned long long *)((char *)d->slab + oF) + POP_DD_MAXWORLD;
    unsigned long long mark = 0;
    POP_CUDA(cudaMemcpyAsync(&mark, slot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    POP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (mark == 0) return POP_OK;
    cudaMemsetAsync(slot, 0, 8, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    pop_set_error("%s: rank %d timed out waiting for a peer at exchange epoch %llu; the result is incomplete", who, d->rank, mark);
    return POP_ERR_CUDA;
}

// exposed for tests / drivers: one row half step on a caller-owned local panel
extern "C" int pop_dd_row_half_step(pop_ctx *ctx, pop_dd *d, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *dF,
                                    int64_t ldf, double *dR, int64_t ldr) {
    POP_REQUIRE(ctx && d && U && dR, POP_ERR_ARG, "pop_dd_row_half_step: null argument");
    POP_REQUIRE(d->opened, POP_ERR_ARG, "pop_dd_row_half_step: peers not opened");
    POP_REQUIRE(ldr == d->ld, POP_ERR_DIM, "pop_dd_row_half_step: the panel must use the leading dimension pop_dd_ld()");
    POP_REQUIRE(!(T && gamma != 0.0) || (dF && ldf >= d->nrows), POP_ERR_DIM, "pop_dd_row_half_step: F panel missing or too small");
    int rc = dd_row_half_step(ctx, d, U, T, gamma, dF, ldf, dR, nullptr);
    if (rc == POP_OK && d->world > 1 && !d->fake_peers) {
        POP_CUDA(cudaStreamSynchronize(ctx->stream));
        rc = dd_check_timeout(ctx, d, "pop_dd_row_half_step");
    }
    return rc;
}

__global__ void k_dd_scale_copy(const double *__restrict__ src, int64_t lds, double *__restrict__ dst, int64_t ldd, int64_t rows, int64_t cols, double s) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) dst[c * ldd + r] = s * src[c * lds + r];
}

// ------------------------------------------------------------------------------------
// ADI of examples/adi_poisson.jl:42-59 on the element-decomposed matrix: no transposes.
//   R <- -F/p_1 ; for j: rows: R <- (R S1_j^{-1}) T2_j - F/q_j ; columns: R <- T1_{j+1} (S2_j^{-1} R) - F/p_{j+1}
//   (last iteration: X = S2_J^{-1} R).  dF_local / dX_local: n x nloc panels in the local column order.
// ------------------------------------------------------------------------------------
extern "C" int pop_adi_poisson_dd(pop_ctx *ctx, pop_dd *d, const pop_arrowhead *K, const pop_arrowhead *M, const double *dF_local, int64_t ldf,
                                  double *dX_local, int64_t ldx, double a, double b, double c, double dd, double tol, int J) {
    POP_REQUIRE(ctx && d && K && M && dF_local && dX_local, POP_ERR_ARG, "pop_adi_poisson_dd: null argument");
    POP_REQUIRE(d->opened, POP_ERR_ARG, "pop_adi_poisson_dd: peers not opened");
    const int64_t n = K->n();
    POP_REQUIRE(n == d->nrows && K->d.m == d->m && K->d.q == d->q, POP_ERR_DIM, "pop_adi_poisson_dd: operator does not match the decomposition");
    POP_REQUIRE(ldf >= n && ldx >= n && !(ldf & 1) && !((uintptr_t)dF_local & 15), POP_ERR_DIM,
                "pop_adi_poisson_dd: panels need leading dimensions >= n (F: even, 16-byte aligned)");
    if (J <= 0) J = pop_adi_num_iterations(a, b, c, dd, tol);
    POP_REQUIRE(J >= 1 && J <= 2000, POP_ERR_ARG, "pop_adi_poisson_dd: iteration count %d out of range", J);
    std::vector<double> p(J), q(J);
    int rc = pop_adi_shifts(J, a, b, c, dd, p.data(), q.data());
    if (rc) return rc;
    std::vector<double> al(2 * J), sg(2 * J), ta(2 * J), tb(2 * J);
    for (int j = 0; j < J; ++j) {
        al[2 * j] = 1.0; sg[2 * j] = 1.0 / p[j];
        al[2 * j + 1] = 1.0; sg[2 * j + 1] = -1.0 / q[j];
        ta[2 * j] = -1.0; tb[2 * j] = 1.0 / p[j];
        ta[2 * j + 1] = -1.0; tb[2 * j + 1] = -1.0 / q[j];
    }
    std::vector<pop_arrowhead *> fac(2 * J, nullptr), top(2 * J, nullptr);
    auto cleanup = [&]() {
        for (auto *h : fac) pop_arrowhead_free(ctx, h);
        for (auto *h : top) pop_arrowhead_free(ctx, h);
    };
    rc = pop_revchol_shifted_batch(ctx, K, M, al.data(), sg.data(), 2 * J, fac.data(), nullptr);
    if (rc) { cleanup(); return rc; }
    {
        pop_arrowhead Ku = *K, Mu = *M;   // upper data (Symmetric views)
        Ku.d.nC = 0; Mu.d.nC = 0;
        Ku.p.D = K->p.D + (size_t)K->d.lD * K->d.q * K->mD; Ku.d.lD = 0;
        Mu.p.D = M->p.D + (size_t)M->d.lD * M->d.q * M->mD; Mu.d.lD = 0;
        rc = pop_launch_axpby_batch(ctx, &Ku, &Mu, ta.data(), tb.data(), 2 * J, top.data());
        if (rc) { cleanup(); return rc; }
    }
    const int64_t ld = d->ld, nloc = d->nloc;
    double *R = d->work;
    {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)std::min<int64_t>(nloc, 32768));
        k_dd_scale_copy<<<grid, 256, 0, ctx->stream>>>(dF_local, ldf, R, ld, n, nloc, -1.0 / p[0]);
        ctx->launches++;
    }
    // segment multipliers of all row factors in one launch
    double *pcall = nullptr;
    const double **dptrs = nullptr;
    {
        std::vector<const double *> hp(2 * J);
        for (int j = 0; j < J; ++j) {
            hp[2 * j] = fac[2 * j]->p.rdA;
            hp[2 * j + 1] = fac[2 * j]->p.Au;
        }
        cudaError_t e = cudaMallocAsync((void **)&pcall, (size_t)J * 2 * POP_DD_MAXWORLD * 8, ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&dptrs, (size_t)2 * J * sizeof(double *), ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dptrs, hp.data(), (size_t)2 * J * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // hp goes out of scope
        if (e != cudaSuccess) {
            pop_set_error("pop_adi_poisson_dd: %s", cudaGetErrorString(e));
            cleanup();
            return POP_ERR_CUDA;
        }
        DDBounds bd;
        for (int r = 0; r <= d->world; ++r) bd.kb[r] = d->kb[r];
        bd.world = d->world; bd.nh = d->nh;
        k_dd_pconst<<<J, 32, 0, ctx->stream>>>(bd, dptrs, nullptr, nullptr, pcall);
        ctx->launches++;
    }
    for (int j = 0; j < J && rc == POP_OK; ++j) {
        if ((rc = dd_row_half_step(ctx, d, fac[2 * j], top[2 * j + 1], -1.0 / q[j], dF_local, ldf, R, pcall + (size_t)j * 2 * POP_DD_MAXWORLD))) break;
        if (j + 1 < J) rc = pop_solve_mul_axpy(ctx, fac[2 * j + 1], top[2 * j + 2], -1.0 / p[j + 1], dF_local, ldf, R, ld, nloc);
        else rc = pop_solve(ctx, fac[2 * j + 1], POP_LEFT, R, ld, nloc, POP_SOLVE_AUTO);
    }
    if (rc == POP_OK) {
        cudaError_t e = cudaMemcpy2DAsync(dX_local, ldx * 8, R, ld * 8, n * 8, nloc, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { pop_set_error("pop_adi_poisson_dd: %s", cudaGetErrorString(e)); rc = POP_ERR_CUDA; }
    }
    cudaFreeAsync(pcall, ctx->stream);
    cudaFreeAsync((void *)dptrs, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (rc == POP_OK) rc = dd_check_timeout(ctx, d, "pop_adi_poisson_dd");
    return rc;
}
