// k_row2 -- single-pass row half step of the 2D ADI on ONE GPU (24 B per entry):
//     rows of R  <-  T (S^{-1} r) + gamma f ,   S = U U^T      (X / F followed by X * A, examples/adi_poisson.jl:54)
//
// A block of 8 consecutive rows of the n x n panel (every memory access is a 64 B run; two neighbouring blocks fill a
// 128 B line) is owned by a thread-block CLUSTER of cs CTAs for the whole half step.  CTA c owns the mesh elements
// [c*mlc, (c+1)*mlc) of the row direction and their hats.  All global traffic is TMA TENSOR traffic
// (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG / UTMAPF) issued by one thread:
//   * the bubble part of the panel is a 3-D tensor (row, element, degree); the next block lands in a shared-memory
//     buffer [degree][element][row] while the current one is worked on in REGISTERS: thread (parity, element, row)
//     keeps the even or the odd degrees of its element's chain (the chains decouple: the first super-diagonal of
//     every D block of K + sigma M is zero) between the backward sweep (src/arrowhead.jl:439-441), the hat solve
//     (:449-453, :472-477) and the forward sweep (:480-482);
//   * F streams through a ring of 8-degree chunks (prefetched into L2 one block ahead), each chunk is overwritten in
//     place by the result T x + gamma f and leaves by a TMA store.
// The hat system of the 8 rows couples the CTAs.  Every CTA broadcasts, with ONE bulk copy per peer through
// distributed shared memory (cp.async.bulk.shared::cluster + mbarrier complete_tx), the right-hand sides of its hats
// with the couplings of its own elements taken off, plus the few backward-swept bubble values of its two edge
// elements; every CTA then solves the whole bidiagonal pair for the 8 rows redundantly (4 lanes per row, segmented
// affine scan) -- one exchange per block, no cluster barrier in the steady state, no second round for the product.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "pop_ptx.cuh"
#include "pop_row.h"

#define R2_RB 8          // rows per block
#define R2_NBMAX 4       // coupling blocks (factor and product)
#define R2_MAXCS 8
#define R2_NS_MAX 8

struct R2Params {
    CUtensorMap mapX3, mapXh, mapF3, mapFh;   // bubbles (row, element, degree) and hats (row, hat) of the panel and of F
    // factor U (upper data, shared D block)
    const double *rd, *w2;         // [q]
    const double *S;               // [nb][3][m]
    const double *rdA, *offA;      // [nh], [nh-1]
    int nb;
    // product operator T = Symmetric(upper data); nbT < 0: plain solve
    const double *t0, *t2;
    const double *Ts;
    const double *TAd, *TAu;
    int nbT;
    double gamma;
    int useF;
    double *X;
    const double *F;
    int64_t ld, ldf, nrows;
    int m, nh, q;
    int cs, mlc, nch, nblk, pf_dist;
    int pairing;                   // clusters 2p and 2p+1 work on the row blocks 2b and 2b+1 at the same time (the two halves of every 128 B line)
    int hfs;                       // doubles per row of HF: HF[row][2 + hat], zeros in front of hat 0 and behind hat nh-1 (even, hfs/2 odd: rows land in different banks)
    int hat_mode;                  // 1: the compute warps solve the hat systems themselves, one warp per row, hat_L hats per lane (Kogge-Stone over the lanes)
    int hat_L;
    int o_ks;                      // tables of the warp-per-row hat scan (hat_mode 1): [2][5][32] step multipliers, then cd, cu, rd as [hat_L][32] (entry (i, lane) = hat lane*L + i - 1)
    int o_pos;                     // int [hfs]: position of hat h inside a row of HF at index h + 1 (hat -1 and the hats behind the last are zero pads)
    int lp;                        // hats per scan segment (8 segments; tables and HF are padded to 8*lp)
    // shared-memory offsets (bytes)
    int o_in, o_inh, o_fh, o_ring, o_mb, o_hf, o_ct, o_tab, o_ttab, o_ftab, o_hrd, o_hcd, o_hcu, o_ss, o_st, total;
    int mbw;                       // doubles per mailbox message
    int f_evict_first;             // F loads carry the L2 evict_first policy
    int use_senders;               // remote mode: the idle warps of the producer group carry the messages to the mailboxes
    long long *dbg;                // POP_ROW2_TRACE: phase time stamps of CTA 0 (clock64), [block][16]
    // remote mode (several GPUs, or forced): the row direction is split over nr = world * csv CTAs that meet through
    // mailboxes in (peer) global memory instead of a cluster's shared memory
    int remote, world, me, csv, nr;
    int kbase;                     // first element of this GPU's panel (tensor coordinates are panel-local)
    unsigned long long epoch;      // of this half step; mailbox / flag parity = epoch & 1
    double *mbox[POP_DD_MAXWORLD];             // mailboxes of all GPUs: [2][nblk][csv readers][nr senders][mbw]
    unsigned long long *merr;      // error slot (a bounded wait gave up)
};

// ---- TMA tensor helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Shared-memory layout.  Everything whose size follows from (CH, DC, mlc) comes first, so that a kernel instantiated for
// a fixed mlc addresses it with compile-time offsets; the arrays that depend on nh, m, cs and on the number of ring slots
// follow (their offsets travel in R2Params).
struct R2Fixed {
    int in, inh, fh, ct, tab, ttab, ftab, gtab, ss, st, se, ring, cb, hbs;
};
__host__ __device__ constexpr int r2_al(int x) { return (x + 127) / 128 * 128; }
__host__ __device__ constexpr R2Fixed r2_fixed(int ch, int dc, int mlc, int rb) {
    R2Fixed f = {};
    const int nchmax = (2 * ch + dc - 1) / dc;
    f.cb = dc * mlc * rb * 8;
    f.hbs = r2_al((mlc + 1) * rb * 8);
    int o = 1024;                                // mbarriers
    f.in = o; o += nchmax * f.cb;
    f.inh = o; o += 2 * f.hbs;
    f.fh = o; o += 2 * f.hbs;
    f.ct = o; o += r2_al(4 * mlc * rb * 8);
    f.tab = o; o += r2_al((2 * ch + 2) * 16);
    f.ttab = o; o += r2_al((2 * ch + 2) * 16);
    f.ftab = o; o += r2_al((2 * ch + 2) * 16);
    f.gtab = o; o += r2_al((2 * ch + 2) * 16);
    f.ss = o; o += r2_al(R2_NBMAX * 3 * mlc * 8);
    f.st = o; o += r2_al(R2_NBMAX * 3 * mlc * 8);
    f.se = o; o += r2_al(4 * R2_NBMAX * 3 * 8 + (2 * mlc + 8) * 8);
    f.ring = o;
    return f;
}

// shared-memory accesses of the hot loops by 32-bit address (no generic 64-bit pointers in the register-critical code)
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// A mailbox entry.  Cluster mode: shared memory.  Remote mode: an 8-byte word in global memory that another CTA -- of this
// or of another GPU -- stores exactly once per half step; it holds R2_EMPTY (a NaN no arithmetic produces) until then, the
// single reader spins on it and puts R2_EMPTY back.  Data and "flag" are the same atomic 8-byte store: no fence, no second
// round trip over NVLink (the slot is reused two half steps later, ordered by the messages in between).
#define R2_EMPTY 0xffffffffffffffffull
// the wait itself (rare: a message usually arrived long before it is read) stays out of line, the kernel inlines only the test
__device__ __noinline__ unsigned long long r2_mbspin(volatile unsigned long long *q, unsigned long long *err) {
    unsigned long long v;
    long long spins = 0;
    while ((v = *q) == R2_EMPTY) {
        if (++spins > 100000000LL) {   // a peer died: do not hang the device
            *err = 1;
            v = 0;
            break;
        }
    }
    return v;
}
__device__ __forceinline__ double r2_mbld(const double *p, int mode, unsigned long long *err) {   // mode 0: shared memory, 1: remote, 2: remote stand-in (profiling: entries are never emptied)
    if (!mode) return *p;
    volatile unsigned long long *q = reinterpret_cast<volatile unsigned long long *>(const_cast<double *>(p));
    unsigned long long v = *q;
    if (v == R2_EMPTY) v = r2_mbspin(q, err);
    if (mode == 1) *q = R2_EMPTY;
    return __longlong_as_double((long long)v);
}

// The same in two steps, so that several entries are requested before the first one is waited for (a mailbox read is an
// L2 or NVLink round trip: issued one after the other they add up on the critical path of the row block).
// (measured on 8 GPUs: ld.relaxed.gpu / ld.global.cg loads and plain or 16-byte stores change nothing, profiles/r2_dd8_variants.txt)
__device__ __forceinline__ unsigned long long r2_mbpeek(const double *p) {
    return *reinterpret_cast<volatile const unsigned long long *>(p);
}
__device__ __forceinline__ double r2_mbtake(const double *p, unsigned long long v, int mode, unsigned long long *err) {
    volatile unsigned long long *q = reinterpret_cast<volatile unsigned long long *>(const_cast<double *>(p));
    if (v == R2_EMPTY) v = r2_mbspin(q, err);
    if (mode == 1) *q = R2_EMPTY;
    return __longlong_as_double((long long)v);
}

// CH chain entries per thread: thread (par, e, r) owns the degrees j = 2 i + par, i < CH, of element e in row r.
// DC degrees per TMA box / ring slot (8, or 4 when shared memory is short: more, smaller slots).
// Threads [0, 16 mlc) compute; one more warp (the last) is the TMA PRODUCER: it alone issues and waits for the tensor
// copies, and meets the compute threads only through mbarriers (no block-wide barrier on the data path).
#define R2_FG 4            // chunks per proxy fence in the output phase
#define R2_BAR_COMPUTE 1   // named barrier of the compute threads
#define R2_STAMP(slot) do { if (P.dbg && blockIdx.x == 0 && tid == 0 && t < 64) P.dbg[t * 16 + (slot)] = clock64(); } while (0)
#define R2_BAR_HATS 2      // compute threads + the two hat warps: hats solved
#define R2_BAR_MSG 4       // 4, 5 (by block parity): compute threads arrive, the sender warps wait: the CTA's message is in the staging buffer
#define R2_BAR_FREE 6      // 6, 7 (by block parity): the sender warps arrive, the compute threads wait: the staging buffer may be rewritten
#define R2_SENDERS 96      // remote mode with the hats on the compute warps: the three idle warps of the producer group send the messages
#define R2_COMPUTE_SLOTS 512   // threads [0, 512): compute (the first 16 mlc of them work), [512, 640): producer warp group
// MLC > 0: elements per CTA fixed at compile time (every hot shared-memory address is then one register plus an immediate).
// RB rows per block: 8 (64 B runs; clusters / groups walk in pairs) or 16 (128 B runs; for panels of few elements per CTA).
// REM: remote mode (mailboxes in peer-visible global memory, hat systems solved by the compute warps) compiled as a kernel of
// its own: the cluster-mode kernel carries none of its code (measured: the shared source cost the one-GPU ADI 5 %).
template <int CH, int DC, int MLC, int RB, bool REM>
__global__ void __launch_bounds__(640, 1) k_row2(const __grid_constant__ R2Params P) {
    extern __shared__ __align__(128) unsigned char r2_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(r2_raw);
    constexpr int NCM = (2 * CH + DC - 1) / DC;   // chunk regions of the row block buffer
    uint64_t *barH = bars;             // [2] the hats of a block (and of F) have landed (by block parity)
    uint64_t *barMB = bars + 2;        // [2] the peers' messages of a block (by block parity)
    uint64_t *barX = bars + 4;         // [NCM] chunk c of the row block has landed (bytes)
    uint64_t *barXfree = barX + NCM;   // [NCM] every compute warp has taken chunk c into registers: its region is free
    uint64_t *barF = barXfree + NCM;   // [NCM] F of chunk c has landed in region c (bytes)
    uint64_t *barR = barF + NCM;       // [NCM] every compute warp has written its results of chunk c into region c
    const int tid = threadIdx.x;
    const int cs = REM ? 1 : P.cs, mlc = MLC ? MLC : P.mlc, nch = P.nch;   // remote mode: no cluster
    const int nthr = 2 * RB * mlc;     // compute threads
    const R2Fixed LF = r2_fixed(CH, DC, mlc, RB);   // compile-time constants when MLC > 0
    constexpr bool remote = REM;
    const int mbmode = REM ? P.remote : 0;         // 0: cluster (shared-memory mailboxes), 1: remote, 2: remote stand-in
    const bool hatm = REM && P.hat_mode != 0;      // the compute warps solve the hat systems (one warp per row)
    const int G = remote ? P.csv : cs;             // CTAs that share a row block on this GPU
    const int nr = remote ? P.nr : cs;             // ... and over all GPUs
    const int rank = remote ? P.me * P.csv + (int)(blockIdx.x % (unsigned)P.csv) : (cs > 1 ? (int)cluster_ctarank() : 0);
    const int m = P.m, nh = P.nh, q = P.q;
    const bool mul = P.nbT >= 0;
    const bool useF = P.useF != 0;
    double *IN = reinterpret_cast<double *>(r2_raw + LF.in);
    double *INH = reinterpret_cast<double *>(r2_raw + LF.inh);     // [2][(mlc+1)*8]
    double *FH = reinterpret_cast<double *>(r2_raw + LF.fh);       // [2][(mlc+1)*8]
    double *MB = reinterpret_cast<double *>(r2_raw + P.o_mb);       // [2][cs][mbw]
    double *HF = reinterpret_cast<double *>(r2_raw + P.o_hf);       // [RB][hfs], entry 2 + i of a row = hat i, zeros around
    const int HFS = P.hfs;
    // row stride of HF: hat_mode 1 keeps a row together ([row][position]), hat_mode 0 keeps the rows of a hat together
    // ([position][row], the layout the two hat warps walk without bank conflicts); posT holds position * (stride of a position)
    const int HSR = hatm ? HFS : 1;
    double *ksT = reinterpret_cast<double *>(r2_raw + P.o_ks);
    // hat h of a row sits at HF[row * HFS + posT(h + 1)]: hat_mode 1 interleaves the hats over the lanes of the scan ((g % L) * 32 + g / L
    // with g = h + 1: the warp's accesses are conflict-free), hat_mode 0 keeps them in order (2 + h)
    const int *posTab = reinterpret_cast<const int *>(r2_raw + P.o_pos);
    auto posT = [&](int g) -> int { return REM ? posTab[g] : (g + 1) * RB; };   // cluster mode: [position][row] with one more pad in front
    double *ct = reinterpret_cast<double *>(r2_raw + LF.ct);       // [3][mlc][8] (both parities summed); later xb [4][mlc][8]
    double2 *tab = reinterpret_cast<double2 *>(r2_raw + LF.tab);   // (rd_j, w2_j)
    double2 *ttab = reinterpret_cast<double2 *>(r2_raw + LF.ttab); // (t0_j, t2_j)
    double2 *ftab = reinterpret_cast<double2 *>(r2_raw + LF.ftab); // (rd_j, w2_{j-2}): the forward sweep's pair
    double2 *gtab = reinterpret_cast<double2 *>(r2_raw + LF.gtab); // impulse responses of the forward sweep to unit changes of the chain entries 0 and 1
    double *hrd = reinterpret_cast<double *>(r2_raw + P.o_hrd);     // [nh]
    double *hcd = reinterpret_cast<double *>(r2_raw + P.o_hcd);     // off(i,i+1) rd_i
    double *hcu = reinterpret_cast<double *>(r2_raw + P.o_hcu);     // off(i-1,i) rd_i
    double *hwd = hcu + (64 / RB) * P.lp;                                   // weights of the downward segment maps
    double *hwu = hwd + (64 / RB) * P.lp;                                   // weights of the upward segment maps (times rd_i)
    double *sS = reinterpret_cast<double *>(r2_raw + LF.ss);       // [4][3][mlc]
    double *sT = reinterpret_cast<double *>(r2_raw + LF.st);       // [4][3][mlc]
    double *sEdge = reinterpret_cast<double *>(r2_raw + LF.se);    // [2][4][3] couplings of the elements k0+mlc (side 0) and k0-1 (side 1): the neighbours' edge elements
    double *tEdge = sEdge + 2 * R2_NBMAX * 3;                      // the same of the product operator
    double *tAd = tEdge + 2 * R2_NBMAX * 3;                        // [mlc+1] diagonal of the product's hat block at my hats
    double *tAu = tAd + mlc + 1;                                   // [mlc+2] its off-diagonal: entry u = coupling of the hats hA+u-1 and hA+u
    const int CBD = LF.cb / 8;                // doubles per chunk of DC degrees
    const int HBD = (mlc + 1) * RB;        // doubles per hat box
    const int HBS = LF.hbs / 8;               // distance of the two parity buffers (TMA destinations: 128 B aligned)
    const int mbw = P.mbw;
    const int k0 = rank * mlc;
    const int hA = k0, hB = rank == nr - 1 ? nh : min(nh, k0 + mlc);
    const int kx0 = k0 - P.kbase;                  // element / hat coordinate inside this GPU's panel
    const int nhl = max(hB - hA, 0);
    const bool producer = tid >= nthr;   // set-up: everybody but the compute threads idles

    // ---- one-time set-up ------------------------------------------------------------------------------
    pdl_launch_dependents();   // the next kernel may set itself up while this one drains
    if (tid == 0) {
        mbar_init(barH, 1);
        mbar_init(barH + 1, 1);
        mbar_init(barMB, 1);
        mbar_init(barMB + 1, 1);
        for (int c = 0; c < NCM; ++c) {
            mbar_init(barX + c, 1);
            mbar_init(barXfree + c, (uint32_t)(nthr / 32));   // one arrival per compute warp
            mbar_init(barF + c, 1);
            mbar_init(barR + c, (uint32_t)(nthr / 32));
        }
        fence_mbar_init();
    }
    if (!producer) {
        for (int j = tid; j < 2 * CH + 2; j += nthr) {
            const bool in = j < q;
            tab[j] = make_double2(in ? P.rd[j] : 0.0, (in && P.w2 && j + 2 < q) ? P.w2[j] : 0.0);
            ttab[j] = make_double2((in && mul) ? P.t0[j] : 0.0, (in && mul && P.t2 && j + 2 < q) ? P.t2[j] : 0.0);
            ftab[j] = make_double2(in ? P.rd[j] : 0.0, (in && P.w2 && j >= 2) ? P.w2[j - 2] : 0.0);
        }
        for (int i = tid; i < (64 / RB) * P.lp; i += nthr) {   // padded to the 64 / RB scan segments of lp hats: zeros behind the last hat
            const double rdv = i < nh ? P.rdA[i] : 0.0;
            hrd[i] = rdv;
            hcd[i] = (i < nh - 1 ? P.offA[i] : 0.0) * rdv;
            hcu[i] = (i > 0 && i < nh ? P.offA[i - 1] : 0.0) * rdv;
        }
        for (int x = tid; x < R2_NBMAX * 3 * mlc; x += nthr) {
            const int b = x / (3 * mlc), t = (x / mlc) % 3, e = x % mlc;
            sS[x] = b < P.nb ? P.S[((size_t)b * 3 + t) * m + k0 + e] : 0.0;
            sT[x] = (mul && b < P.nbT) ? P.Ts[((size_t)b * 3 + t) * m + k0 + e] : 0.0;
        }
        for (int x = tid; x < 2 * R2_NBMAX * 3; x += nthr) {
            const int side = x / (R2_NBMAX * 3), b = (x / 3) % R2_NBMAX, t = x % 3;
            const int kk = side == 0 ? k0 + mlc : k0 - 1;   // side 0: the right neighbour's first element, side 1: the left neighbour's last
            const bool in = kk >= 0 && kk < m;
            sEdge[x] = (in && b < P.nb) ? P.S[((size_t)b * 3 + t) * m + kk] : 0.0;
            tEdge[x] = (in && mul && b < P.nbT) ? P.Ts[((size_t)b * 3 + t) * m + kk] : 0.0;
        }
        for (int x = tid; x < mlc + 2; x += nthr) {
            const int i = hA + x;
            if (x <= mlc) tAd[x] = (mul && i < nh) ? P.TAd[i] : 0.0;
            tAu[x] = (mul && i >= 1 && i < nh) ? P.TAu[i - 1] : 0.0;
        }
        for (int x = tid; x < HFS * RB; x += nthr) HF[x] = 0.0;   // the pads around the hats stay zero for the whole kernel
        for (int g = tid; g < HFS; g += nthr) {
            int *pw = reinterpret_cast<int *>(r2_raw + P.o_pos);
            pw[g] = hatm ? ((g < 32 * P.hat_L) ? (g % P.hat_L) * 32 + g / P.hat_L : 32 * P.hat_L) : min(g + 1, HFS - 1) * RB;
        }
    }
    __syncthreads();
    if (tid < 64 / RB) {
        // segment maps of the two hat recurrences: h_bottom = sum_i wd_i ytil_i + Ad h_in (downward),
        // x_top = sum_i wu_i h_i + Au x_in (upward, wu includes rd_i); the products Ad, Au are rebuilt by the hat lanes
        const int sg = tid, a0 = sg * P.lp;
        double pd = 1.0;
        for (int i = 0; i < P.lp; ++i) {
            hwd[a0 + i] = pd;
            pd *= -hcd[a0 + i];
        }
        double pu = 1.0;
        for (int i = P.lp - 1; i >= 0; --i) {
            hwu[a0 + i] = pu * hrd[a0 + i];
            pu *= -hcu[a0 + i];
        }
    }
    if (hatm && tid >= 32 && tid < 64) {
        // warp-per-row hat scan: lane l owns the hats [l L, (l+1) L); its multipliers and the Kogge-Stone step multipliers
        const int lane = tid - 32, L = P.hat_L;
        double ad = 1.0, au = 1.0;
        for (int i = 0; i < L; ++i) {
            const int hi = lane * L + i - 1;   // entry 0 of lane 0 is the zero pad in front of hat 0
            const bool in = hi >= 0 && hi < nh;
            const double rdv = in ? P.rdA[hi] : 0.0;
            const double cdv = (in && hi < nh - 1 ? P.offA[hi] : 0.0) * rdv, cuv = (in && hi > 0 ? P.offA[hi - 1] : 0.0) * rdv;
            ksT[(10 + i) * 32 + lane] = cdv;
            ksT[(10 + L + i) * 32 + lane] = cuv;
            ksT[(10 + 2 * L + i) * 32 + lane] = rdv;
            ad *= -cdv;
            au *= -cuv;
        }
        // step st of the suffix scan multiplies what comes from lane l + 2^st by the product of the multipliers of the lanes
        // [l, l + 2^st); the prefix scan (upward recurrence) mirrors it
#pragma unroll 1
        for (int st = 0; st < 5; ++st) {
            double pd = 1.0, pu = 1.0;
#pragma unroll 1
            for (int j = 0; j < (1 << st); ++j) {
                pd *= __shfl_sync(POP_FULL_MASK, ad, (lane + j) & 31);
                pu *= __shfl_sync(POP_FULL_MASK, au, (lane - j) & 31);
            }
            ksT[st * 32 + lane] = (lane + (1 << st) < 32) ? pd : 0.0;   // zero where the step has no partner: the scan needs no branch
            ksT[(5 + st) * 32 + lane] = (lane >= (1 << st)) ? pu : 0.0;
        }
    }
    if (tid < 2) {
        // x = forward sweep of z is linear: a change of z_0 (z_1) of a chain by -d moves x_i by -d * G0_i (G1_i).  The hats enter
        // the forward sweep only through the first two chain entries (src/arrowhead.jl:475-477), so the sweep can run before the
        // hats are known and be corrected afterwards.
        double g0 = 0.0, g1 = 0.0;
        for (int i = 0; i < CH; ++i) {
            const double2 cf = ftab[2 * i + tid];
            g0 = ((i == 0 ? 1.0 : 0.0) - cf.y * g0) * cf.x;
            g1 = ((i == 1 ? 1.0 : 0.0) - cf.y * g1) * cf.x;
            gtab[2 * i + tid] = make_double2(g0, g1);
        }
    }
    __syncthreads();
    if (cs > 1) {   // every CTA's mbarriers are initialised before a peer sends to them
        cluster_arrive();
        cluster_wait();
    }
    pdl_wait();   // the set-up above read the factor only; panel, F and mailboxes belong to the previous kernel until here

    // row blocks of this cluster: block t is number bbase + t * bstride.  With pairing, two clusters walk through
    // neighbouring row blocks side by side, so that both 64 B halves of every 128 B line are requested within
    // microseconds of each other (the second request hits L2 instead of fetching the line from HBM again).
    const int ncl = gridDim.x / G, cl = blockIdx.x / G;
    int bbase, bstride, nloc;
    if (P.pairing) {
        const int npt = (P.nblk + 1) / 2, npc = ncl / 2;
        const int per = (npt + npc - 1) / npc;
        const int p0 = (cl >> 1) * per, p1 = min(npt, p0 + per);
        bbase = 2 * p0 + (cl & 1);
        bstride = 2;
        nloc = max(p1 - p0, 0);
        if (nloc > 0 && bbase + 2 * (nloc - 1) >= P.nblk) --nloc;
    } else {
        const int per = (P.nblk + ncl - 1) / ncl;
        bbase = cl * per;
        bstride = 1;
        nloc = max(min(P.nblk, bbase + per) - bbase, 0);
    }

    // Register budget (5 warps per scheduler: 96 registers each at launch): the producer warp group (threads 512 .. 639:
    // the TMA producer lane and the two hat warps) hands registers back to the CTA pool, the four compute warp groups take
    // 112 each from it: 4 warps x (96 - 32) released = 16 warps x (112 - 96) taken -- the pool holds nothing else)
    if (tid >= R2_COMPUTE_SLOTS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        // =========================== TMA producer (one lane) =====================================================
        if (tid == R2_COMPUTE_SLOTS && nloc > 0) {
            // The row block buffer IN is time-multiplexed: chunk c of X(t) lives in region NCM-1-c until the backward sweep
            // has taken it into registers; F(t) of chunk r then lands in region r (requested as the regions drain, i.e. a
            // whole hat phase before it is needed), is overwritten in place by the results and leaves by a TMA store; as
            // soon as the store has read region r, chunk NCM-1-r of X(t+1) is requested into it.  The backward sweep
            // walks the chunks downwards and the output phase upwards, so both find their data in the order of arrival.
            const uint32_t cbytes = (uint32_t)(CBD * 8);
            auto load_x = [&](int t, int c) {      // chunk c of block t -> region NCM-1-c
                mbar_expect_tx(barX + c, cbytes);
                tma_load_3d(IN + (NCM - 1 - c) * CBD, &P.mapX3, (bbase + t * bstride) * RB, kx0, DC * c, barX + c);
            };
            const uint64_t polF = l2_policy_evict_first();
            auto load_f = [&](int r, int row0) {   // F of chunk r -> region r (a stream without reuse: evict_first when asked)
                if (P.f_evict_first)
                    tma_load_3d_hint(IN + r * CBD, &P.mapF3, row0, kx0, DC * r, barF + r, polF);
                else
                    tma_load_3d(IN + r * CBD, &P.mapF3, row0, kx0, DC * r, barF + r);
            };
            auto load_h = [&](int t) {             // hats of block t (and of F)
                mbar_expect_tx(barH + (t & 1), (uint32_t)(HBD * 8 * (useF ? 2 : 1)));
                tma_load_2d(INH + (t & 1) * HBS, &P.mapXh, (bbase + t * bstride) * RB, kx0, barH + (t & 1));
                if (useF) tma_load_2d(FH + (t & 1) * HBS, &P.mapFh, (bbase + t * bstride) * RB, kx0, barH + (t & 1));
            };
            load_h(0);
            for (int c = nch - 1; c >= 0; --c) load_x(0, c);
            for (int t = 0; t < nloc; ++t) {
                const int row0 = (bbase + t * bstride) * RB;
                const uint32_t ph = (uint32_t)t & 1u;
                const bool more = t + 1 < nloc;
                // phase 1: regions drain (backward sweep) -> F(t) in, or X(t+1) where the region never holds results
                if (useF)
                    for (int r = 0; r < NCM - nch && r < nch; ++r) {   // regions that never hold X (q below the class maximum)
                        mbar_expect_tx(barF + r, cbytes);
                        load_f(r, row0);
                    }
                for (int c = nch - 1; c >= 0; --c) {
                    const int r = NCM - 1 - c;
                    mbar_wait(barXfree + c, ph);
                    if (r < nch) {
                        if (useF) {
                            mbar_expect_tx(barF + r, cbytes);
                            load_f(r, row0);
                        }
                    } else if (more) {
                        load_x(t + 1, c);
                    }
                }
                if (more) load_h(t + 1);
                // phase 2: results out, X(t+1) in behind them
                for (int r = 0; r < nch; ++r) {
                    mbar_wait(barR + r, ph);
                    tma_store_3d(&P.mapX3, row0, kx0, DC * r, IN + r * CBD);
                    tma_commit();
                    if (r >= 1) {
                        tma_wait_read<1>();            // the store of region r-1 has read it
                        const int c = NCM - 1 - (r - 1);
                        if (more && c < nch) load_x(t + 1, c);
                    }
                }
                tma_wait_read<0>();
                {
                    const int c = NCM - 1 - (nch - 1);
                    if (more && c < nch) load_x(t + 1, c);
                }
            }
            tma_wait_all();
        }
        __syncwarp();
        if (tid >= R2_COMPUTE_SLOTS + 32 && tid < R2_COMPUTE_SLOTS + 96) {
            // =========================== two hat warps ===========================================================
            // The two bidiagonal recurrences of the hat block (src/arrowhead.jl:453, :472) for the 8 rows of the block:
            // each warp takes 4 rows, lane = seg*4 + row, segment seg of 8 owns the hats [seg*lp, (seg+1)*lp) (tables padded
            // with zeros).  Both recurrences h_i = y_i - c_i h_next are affine in the incoming value: every segment builds
            // its map (A, B), the maps are composed across the 8 segments by shuffles, then the segment re-runs with the
            // exact incoming value -- and builds the map of the second recurrence in the same pass.  (These warps run
            // on 32 registers: the compute warps need the rest.)
            constexpr int RW = RB / 2, NSEG = 32 / RW;   // rows per hat warp, scan segments per row
            const int lane = tid & 31, l4 = lane % RW, rr = l4 + RW * ((tid - R2_COMPUTE_SLOTS - 32) >> 5), seg = lane / RW;
            const int lp = P.lp;
            double *col = HF + rr + (2 + seg * lp) * RB;   // hat_mode 0 layout: [position][row]
            const double *cd = hcd + seg * lp, *cu = hcu + seg * lp, *rdp = hrd + seg * lp;
            const double *wd = hwd + seg * lp, *wu = hwu + seg * lp;
            double Ad = 1.0, Au = 1.0;   // the segment's multipliers of the incoming values
            for (int i = 0; i < lp; ++i) {
                Ad *= -cd[i];
                Au *= -cu[i];
            }
            for (int t = 0; t < (hatm ? 0 : nloc); ++t) {   // hat_mode 1: the compute warps solve the hats, these two warps idle
                named_bar_sync(R2_BAR_HATS, nthr + 64);   // the right-hand sides (times rd) of all hats are in HF
                // pass 1: what leaves my segment at the bottom with zero coming in: a dot product, no chain
                double b0 = 0.0, b1 = 0.0;
                for (int i0 = 0; i0 < lp; i0 += 4) {
                    b0 = fma(wd[i0], col[i0 * RB], b0);
                    b1 = fma(wd[i0 + 1], col[(i0 + 1) * RB], b1);
                    b0 = fma(wd[i0 + 2], col[(i0 + 2) * RB], b0);
                    b1 = fma(wd[i0 + 3], col[(i0 + 3) * RB], b1);
                }
                b0 += b1;
                double inc = 0.0;
                {
                    double res = 0.0;   // value leaving the segments above, built from the top
#pragma unroll
                    for (int s2 = NSEG - 1; s2 >= 0; --s2) {
                        const double As = __shfl_sync(POP_FULL_MASK, Ad, s2 * RW + l4), Bs = __shfl_sync(POP_FULL_MASK, b0, s2 * RW + l4);
                        if (s2 == seg) inc = res;
                        res = fma(As, res, Bs);
                    }
                }
                // pass 2, descending: the final h_i = ytil_i - cd_i h_{i+1}; on the way the dot product of the upward map
                double h = inc;
                b0 = 0.0;
                for (int i0 = lp - 2; i0 >= 0; i0 -= 2) {
                    const double y1 = col[(i0 + 1) * RB], y0 = col[i0 * RB];
                    const double c1 = cd[i0 + 1], c0 = cd[i0], w1 = wu[i0 + 1], w0 = wu[i0];
                    h = fma(-c1, h, y1);
                    col[(i0 + 1) * RB] = h;
                    b0 = fma(w1, h, b0);
                    h = fma(-c0, h, y0);
                    col[i0 * RB] = h;
                    b0 = fma(w0, h, b0);
                }
                {
                    double res = 0.0;   // value entering from the segments below, built from the bottom
                    inc = 0.0;
#pragma unroll
                    for (int s2 = 0; s2 < NSEG; ++s2) {
                        const double As = __shfl_sync(POP_FULL_MASK, Au, s2 * RW + l4), Bs = __shfl_sync(POP_FULL_MASK, b0, s2 * RW + l4);
                        if (s2 == seg) inc = res;
                        res = fma(As, res, Bs);
                    }
                }
                // pass 3, ascending: the final x_i = h_i rd_i - cu_i x_{i-1}
                h = inc;
                for (int i0 = 0; i0 < lp; i0 += 2) {
                    const double y0 = col[i0 * RB] * rdp[i0], y1 = col[(i0 + 1) * RB] * rdp[i0 + 1];
                    const double c0 = cu[i0], c1 = cu[i0 + 1];
                    h = fma(-c0, h, y0);
                    col[i0 * RB] = h;
                    h = fma(-c1, h, y1);
                    col[(i0 + 1) * RB] = h;
                }
                __syncwarp();
                named_bar_sync(R2_BAR_HATS, nthr + 64);   // the hats of the 8 rows are solved
            }
        }
        if (REM && hatm && P.use_senders && tid >= R2_COMPUTE_SLOTS + 32) {
            // =========================== three sender warps (remote mode) ===============================================
            // The message of a block goes to every CTA that shares the row block, on every GPU: 8 byte stores into their
            // mailboxes.  These warps take that off the compute warps, which go on with the forward sweep; the staging
            // buffer alternates with the block parity, two named barriers per parity hand it back and forth.
            const int st = tid - (R2_COMPUTE_SLOTS + 32);
            for (int t = 0; t < nloc; ++t) {
                const int pb = t & 1;
                named_bar_sync(R2_BAR_MSG + pb, nthr + R2_SENDERS);   // the message of block t is complete
                const double *src = MB + (size_t)pb * mbw;
                const size_t mslot = ((size_t)(P.epoch & 1) * P.nblk + (size_t)(bbase + t * bstride)) * G;
                for (int gi = 0; gi < P.world; ++gi) {
                    const int g = (P.me + 1 + gi) % P.world;   // every GPU starts with another target; its own mailboxes come last
                    for (int dc = 0; dc < G; ++dc) {
                        double *dst = P.mbox[g] + ((mslot + dc) * nr + rank) * mbw;
                        // two entries per store (mbw is even, the mailboxes are 16-byte aligned); every 8-byte word still arrives whole
                        for (int x = 2 * st; x < mbw; x += 2 * R2_SENDERS)
                            asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(dst + x), "d"(src[x]), "d"(src[x + 1]) : "memory");
                    }
                }
                named_bar_arrive(R2_BAR_FREE + pb, nthr + R2_SENDERS);   // the staging buffer of this parity is free again
            }
        }
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
      if (tid < nthr) {
        // =========================== compute threads ============================================================
        const int par = tid / (mlc * RB);          // 0 / 1
        const int tp = tid - par * mlc * RB;
        const int e = tp / RB, r = tp % RB;
        const int k = k0 + e;
        // 32-bit shared addresses of this thread's entries: one base register, the rest are immediates when MLC > 0
        const uint32_t sbase = smem_u32(r2_raw);
        const uint32_t off_er = sbase + (uint32_t)(((par * mlc + e) * RB + r) * 8);   // (parity, element, row) inside a [degree][element][row] box
        const uint32_t dstride = (uint32_t)(2 * mlc * RB * 8);                       // two degrees further
        const uint32_t a_tab = sbase + LF.tab + par * 16, a_ttab = sbase + LF.ttab + par * 16, a_ftab = sbase + LF.ftab + par * 16;
        const uint32_t a_in = off_er + LF.in;
        const uint32_t a_ss = sbase + LF.ss + (par * 3 * mlc + e) * 8, a_st = sbase + LF.st + (par * 3 * mlc + e) * 8;
        const uint32_t a_hfr = sbase + P.o_hf + r * HSR * 8;              // row r of HF; hats k-1, k, k+1 through posT
        for (int t = 0; t < nloc; ++t) {
            const int pb = t & 1;
            const uint32_t phMB = (uint32_t)(t >> 1) & 1u;
            const int row0 = (bbase + t * bstride) * RB;
            const double *inh = INH + pb * HBS, *fh = FH + pb * HBS;
            const bool senders = REM && hatm && P.use_senders;   // the sender warps carry the message to the mailboxes
            double *mbme = remote ? MB + (senders ? (size_t)pb * mbw : 0) : MB + ((size_t)pb * cs + rank) * mbw;   // remote: staging of my message only
            const size_t mslot = ((size_t)(P.epoch & 1) * P.nblk + (size_t)(bbase + t * bstride)) * G;   // mailboxes of this row block: one per reading CTA
            const double *mbp = remote ? P.mbox[P.me] + (mslot + (blockIdx.x % (unsigned)G)) * nr * mbw : MB + (size_t)pb * cs * mbw;
            // ---- A+B: the row block, shared memory -> registers, chunk by chunk from the top degrees down, and the
            //      backward sweep  z_j = (b_j - w2_j z_{j+2}) rd_j  on the way
            R2_STAMP(0);
            double z[CH];
            {
                double2 ca = lds_v2f64(a_tab + (CH - 1) * 32), cb = lds_v2f64(a_tab + (CH - 2) * 32);   // coefficients two steps ahead
#pragma unroll
                for (int cc = NCM - 1; cc >= 0; --cc) {
                    if (cc < nch) {
                        mbar_wait(barX + cc, (uint32_t)t & 1u);
                        if (cc == NCM - 1) R2_STAMP(1);
#pragma unroll
                        for (int i4 = DC / 2 - 1; i4 >= 0; --i4) {
                            const int i = (DC / 2) * cc + i4;
                            if (i < CH) z[i] = lds_f64(a_in + (uint32_t)((NCM - 1 - cc) * (LF.cb)) + (uint32_t)i4 * dstride);
                        }
                        __syncwarp();
                        if ((tid & 31) == 0) mbar_arrive(barXfree + cc);
                    } else {
#pragma unroll
                        for (int i4 = 0; i4 < DC / 2; ++i4)
                            if ((DC / 2) * cc + i4 < CH) z[(DC / 2) * cc + i4] = 0.0;
                    }
#pragma unroll
                    for (int i4 = DC / 2 - 1; i4 >= 0; --i4) {
                        const int i = (DC / 2) * cc + i4;
                        if (i < CH) {
                            const double2 cf = ca;
                            ca = cb;
                            if (i >= 2) cb = lds_v2f64(a_tab + (i - 2) * 32);
                            double v = z[i];
                            if (i + 1 < CH) v = fma(-cf.y, z[i + 1], v);
                            z[i] = v * cf.x;
                        }
                    }
                }
            }
            R2_STAMP(2);
            if (senders && t >= 2) named_bar_sync(R2_BAR_FREE + pb, nthr + R2_SENDERS);   // the sender warps are done with this staging buffer (block t - 2)
            {
                double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    const int b = 2 * ii + par;
                    const uint32_t sb = a_ss + ii * 6 * mlc * 8;
                    c0 = fma(lds_f64(sb), z[ii], c0);
                    c1 = fma(lds_f64(sb + mlc * 8), z[ii], c1);
                    c2 = fma(lds_f64(sb + 2 * mlc * 8), z[ii], c2);
                    // the backward-swept coupled bubbles of my two edge elements travel with the message
                    if (e == 0) mbme[(mlc + 2) * RB + (0 * R2_NBMAX + b) * RB + r] = z[ii];
                    if (e == mlc - 1) mbme[(mlc + 2) * RB + (1 * R2_NBMAX + b) * RB + r] = z[ii];
                }
                // the couplings of both parities of an (element, row) are summed in shared memory: odd degrees first
                if (par == 1) {
                    ct[(0 * mlc + e) * RB + r] = c0;
                    ct[(1 * mlc + e) * RB + r] = c1;
                    ct[(2 * mlc + e) * RB + r] = c2;
                }
                named_bar_sync(R2_BAR_COMPUTE, nthr);
                if (par == 0) {
                    ct[(0 * mlc + e) * RB + r] += c0;
                    ct[(1 * mlc + e) * RB + r] += c1;
                    ct[(2 * mlc + e) * RB + r] += c2;
                }
            }
            named_bar_sync(R2_BAR_COMPUTE, nthr);
            // ---- C: my message: hat right-hand sides minus the couplings of MY elements, for hats k0-1 .. k0+mlc --
            R2_STAMP(3);
            mbar_wait(barH + pb, phMB);   // the hats of this block (loaded a block ahead)
            if (tid < (mlc + 2) * RB) {
                const int u = tid / RB - 1, rr = tid % RB;
                double v = (u >= 0 && u < nhl) ? inh[u * RB + rr] : 0.0;
#pragma unroll
                for (int tt = 0; tt < 3; ++tt) {
                    const int el = u + 1 - tt;   // slot tt of element el couples to hat el + tt - 1
                    if (el >= 0 && el < mlc) v -= ct[(tt * mlc + el) * RB + rr];
                }
                mbme[tid] = v;
            }
            if (remote && senders) {
                named_bar_arrive(R2_BAR_MSG + pb, nthr + R2_SENDERS);   // the sender warps take it from here
            } else if (remote) {
                // my message into the mailbox of every CTA that shares this row block, on every GPU (mine included)
                named_bar_sync(R2_BAR_COMPUTE, nthr);
                for (int g = 0; g < P.world; ++g)
                    for (int dc = 0; dc < G; ++dc) {
                        volatile double *dst = P.mbox[g] + ((mslot + dc) * nr + rank) * mbw;
                        for (int x = tid; x < mbw; x += nthr) dst[x] = mbme[x];
                    }
            } else if (cs > 1) {
                fence_proxy_async();
                named_bar_sync(R2_BAR_COMPUTE, nthr);
                if (tid < cs && tid != rank) dsmem_bulk_copy(mapa_u32(mbme, tid), mbme, (uint32_t)(mbw * 8), mapa_u32(barMB + pb, tid));
                if (tid == 0) mbar_expect_tx(barMB + pb, (uint32_t)(mbw * 8 * (cs - 1)));
            }
            // ---- E1: forward sweep (:480-482) WITHOUT the hat corrections, while the hat warps wait for the peers' messages
            //      and solve the hat systems
            R2_STAMP(4);
            {
                double2 ca = lds_v2f64(a_ftab), cb = lds_v2f64(a_ftab + 32);   // (rd_j, w2_{j-2}) two steps ahead
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const double2 cf = ca;
                    ca = cb;
                    if (i + 2 < CH) cb = lds_v2f64(a_ftab + (i + 2) * 32);
                    double v = z[i];
                    if (i >= 1) v = fma(-cf.y, z[i - 1], v);
                    z[i] = v * cf.x;
                }
            }
            R2_STAMP(5);
            // the messages of the peers have landed long ago (plain CTA-scope wait: the bytes were written into this CTA's
            // shared memory by the async proxy and completed on this CTA's mbarrier, as for a TMA load)
            if (remote) {
                // nothing to wait for here: the mailbox loads below spin on the entries themselves
            } else if (cs > 1) {
                mbar_wait(barMB + pb, phMB);
            } else {
                named_bar_sync(R2_BAR_COMPUTE, nthr);
            }
            R2_STAMP(10);
            {
                // right-hand sides (times rd) of ALL hats of the 8 rows, redundantly in every CTA
                if (REM) {
                    // remote mailboxes: every thread takes the main entries (hat right-hand sides) of all messages in ONE linear,
                    // coalesced sweep, four requests in flight before the first is waited for.  Slot u + 1 of message c is hat
                    // c mlc + u; the slots 0 and mlc + 1 belong to the neighbours' edge hats: they are parked in HE (the idle
                    // coupling buffer) and added behind a barrier, so no entry is read twice and no address depends on a branch.
                    const int per = (mlc + 2) * RB, tot = nr * per;
                    double *HE = ct;   // [2][nr][RB]
                    for (int y0 = tid; y0 < tot; y0 += 4 * nthr) {
                        unsigned long long va[4];
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int y = y0 + kk * nthr;
                            const int c = y / per;
                            va[kk] = y < tot ? r2_mbpeek(mbp + c * mbw + (y - c * per)) : 0ull;
                        }
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int y = y0 + kk * nthr;
                            if (y < tot) {
                                const int c = y / per, w = y - c * per, slot = w / RB, rr = w % RB;
                                const double val = r2_mbtake(mbp + c * mbw + w, va[kk], mbmode, P.merr);
                                const int i = c * mlc + slot - 1;
                                if (slot == 0) HE[c * RB + rr] = val;
                                else if (slot == mlc + 1) HE[(nr + c) * RB + rr] = val;
                                else if (i < nh) HF[rr * HSR + posT(i + 1)] = val * hrd[i];
                            }
                        }
                    }
                    named_bar_sync(R2_BAR_COMPUTE, nthr);
                    for (int x = tid; x < 2 * nr * RB; x += nthr) {
                        const int side = x / (nr * RB), c = (x / RB) % nr, rr = x % RB;
                        const int i = side == 0 ? c * mlc - 1 : (c + 1) * mlc;   // the hat next to message c's range
                        if (i >= 0 && i < nh) {
                            double *dst = HF + rr * HSR + posT(i + 1);
                            const double val = HE[x] * hrd[i];
                            if (side == 1 && c == nr - 1) *dst = val;   // the last hat of a ContinuousPolynomial{1} basis has no other entry
                            else *dst += val;
                        }
                    }
                } else
                for (int x = tid; x < nh * RB; x += nthr) {
                    const int i = x / RB, rr = x % RB;
                    const int c = min(i / mlc, nr - 1), u = i - c * mlc;
                    double v = r2_mbld(mbp + c * mbw + (u + 1) * RB + rr, mbmode, P.merr);
                    if (u == 0 && c > 0) v += r2_mbld(mbp + (c - 1) * mbw + (mlc + 1) * RB + rr, mbmode, P.merr);
                    if (u == mlc - 1 && c < nr - 1) v += r2_mbld(mbp + (c + 1) * mbw + rr, mbmode, P.merr);
                    HF[rr * HSR + posT(i + 1)] = v * hrd[i];
                }
            }
            R2_STAMP(11);
            if (hatm) {
                // The two bidiagonal recurrences of the hat block (src/arrowhead.jl:453, :472), one warp per row, lane l owns the
                // hats [l L, (l+1) L): local recurrence with zero coming in, Kogge-Stone over the lanes with the multipliers of
                // the factor (tables, the same for every row), local recurrence again with the exact incoming value.
                named_bar_sync(R2_BAR_COMPUTE, nthr);   // the right-hand sides (times rd) of all hats are in HF
                R2_STAMP(12);
                const int w = tid >> 5, lane = tid & 31;
                if (w < RB) {
                    const int L = P.hat_L;
                    double *rowp = HF + w * HFS + lane;   // entry (i, lane) at rowp[i * 32]
                    const double *cdT = ksT + 10 * 32 + lane, *cuT = cdT + L * 32, *rdT = cuT + L * 32;
                    // (every pass re-reads the row from shared memory instead of keeping it in registers: the chain of the
                    // thread's element occupies them)
                    double tdn = 0.0;
                    for (int i = L - 1; i >= 0; --i) tdn = fma(-cdT[i * 32], tdn, rowp[i * 32]);
#pragma unroll
                    for (int st = 0; st < 5; ++st) {
                        const double o = __shfl_down_sync(POP_FULL_MASK, tdn, 1 << st);
                        tdn = fma(ksT[st * 32 + lane], o, tdn);
                    }
                    double hin = __shfl_down_sync(POP_FULL_MASK, tdn, 1);
                    if (lane == 31) hin = 0.0;
                    for (int i = L - 1; i >= 0; --i) {
                        hin = fma(-cdT[i * 32], hin, rowp[i * 32]);   // h_i = ytil_i - cd_i h_{i+1}
                        rowp[i * 32] = hin * rdT[i * 32];
                    }
                    double tup = 0.0;
                    for (int i = 0; i < L; ++i) tup = fma(-cuT[i * 32], tup, rowp[i * 32]);
#pragma unroll
                    for (int st = 0; st < 5; ++st) {
                        const double o = __shfl_up_sync(POP_FULL_MASK, tup, 1 << st);
                        tup = fma(ksT[(5 + st) * 32 + lane], o, tup);
                    }
                    double xin = __shfl_up_sync(POP_FULL_MASK, tup, 1);
                    if (lane == 0) xin = 0.0;
                    for (int i = 0; i < L; ++i) {
                        xin = fma(-cuT[i * 32], xin, rowp[i * 32]);   // x_i = h_i rd_i - cu_i x_{i-1}
                        rowp[i * 32] = xin;
                    }
                }
                R2_STAMP(13);
                named_bar_sync(R2_BAR_COMPUTE, nthr);   // the hats of all rows are solved
            } else {
                named_bar_sync(R2_BAR_HATS, nthr + 64);   // the hat warps take over ...
                named_bar_sync(R2_BAR_HATS, nthr + 64);   // ... and have solved the hat systems of the rows
            }
            // ---- E2: the corrections of the coupled bubbles (:475-477) carried through the sweep by its impulse responses,
            //      product, output
            R2_STAMP(6);
            const double h0 = lds_f64(a_hfr + posT(k) * 8), h1 = lds_f64(a_hfr + posT(k + 1) * 8), h2 = lds_f64(a_hfr + posT(k + 2) * 8);
            double d0, d1;
            {
                const uint32_t sb = a_ss;
                d0 = lds_f64(sb) * h0;
                d0 = fma(lds_f64(sb + mlc * 8), h1, d0);
                d0 = fma(lds_f64(sb + 2 * mlc * 8), h2, d0);
                const uint32_t sc = a_ss + 6 * mlc * 8;
                d1 = lds_f64(sc) * h0;
                d1 = fma(lds_f64(sc + mlc * 8), h1, d1);
                d1 = fma(lds_f64(sc + 2 * mlc * 8), h2, d1);
            }
            const uint32_t a_gtab = sbase + LF.gtab + par * 16;
            auto corrected = [&](int i) -> double {   // x_i = x0_i - d0 G0_i - d1 G1_i
                const double2 g = lds_v2f64(a_gtab + i * 32);
                return fma(-d1, g.y, fma(-d0, g.x, z[i]));
            };
            double *xb = ct;   // [4][mlc][8]: x of the coupled bubbles, for the hat rows of the product
            double xm = 0.0, x0 = corrected(0), xp = CH > 1 ? corrected(1) : 0.0;   // x_{i-1}, x_i, x_{i+1}
            if (mul) {
                xb[(par * mlc + e) * RB + r] = x0;
                if (CH > 1) xb[((2 + par) * mlc + e) * RB + r] = xp;
            }
            // chunk cc holds the chain entries i = cc*DC/2 .. (cc+1)*DC/2 - 1 (registers are indexed statically)
            R2_STAMP(7);
#pragma unroll
            for (int cc = 0; cc < (2 * CH + DC - 1) / DC; ++cc) {
                if (cc < nch) {
                    const uint32_t slot = a_in + (uint32_t)(cc * LF.cb);   // region cc: F of this chunk in, results out
                    // this chunk's product coefficients (t0_j, t2_j), t2_{j-2}: loaded before the wait for F
                    double2 tc[DC / 2];
                    double tm[DC / 2];
                    if (mul) {
#pragma unroll
                        for (int i4 = 0; i4 < DC / 2; ++i4) {
                            const int i = (DC / 2) * cc + i4;
                            if (i < CH) {
                                tc[i4] = lds_v2f64(a_ttab + i * 32);
                                tm[i4] = i >= 1 ? lds_v2f64(a_ttab + (i - 1) * 32).y : 0.0;
                            }
                        }
                    }
                    if (useF) mbar_wait(barF + cc, (uint32_t)t & 1u);      // F of this chunk has landed
#pragma unroll
                    for (int i4 = 0; i4 < DC / 2; ++i4) {
                        const int i = (DC / 2) * cc + i4;
                        if (i < CH) {
                            const uint32_t sp = slot + i4 * dstride;
                            double out;
                            if (mul) {
                                double acc = tc[i4].x * x0;
                                if (i + 1 < CH) acc = fma(tc[i4].y, xp, acc);
                                if (i >= 1) acc = fma(tm[i4], xm, acc);
                                if (i < 2) {
                                    const uint32_t tb = a_st + i * 6 * mlc * 8;
                                    acc = fma(lds_f64(tb), h0, acc);
                                    acc = fma(lds_f64(tb + mlc * 8), h1, acc);
                                    acc = fma(lds_f64(tb + 2 * mlc * 8), h2, acc);
                                }
                                out = useF ? fma(P.gamma, lds_f64(sp), acc) : acc;
                            } else {
                                out = x0;
                            }
                            sts_f64(sp, out);
                            xm = x0;
                            x0 = xp;
                            if (i + 2 < CH) xp = corrected(i + 2);
                        }
                    }
                    // one proxy fence (a CTA-wide memory barrier in SASS) per group of R2_FG chunks, then the group is handed
                    // to the producer
                    if ((cc % R2_FG) == R2_FG - 1 || cc == nch - 1) {
                        fence_proxy_async();           // my writes are visible to the TMA store
                        __syncwarp();
                        if ((tid & 31) == 0) {
#pragma unroll
                            for (int c2 = cc - (cc % R2_FG); c2 <= cc; ++c2) mbar_arrive(barR + c2);
                        }
                    }
                }
            }
            R2_STAMP(8);
            named_bar_sync(R2_BAR_COMPUTE, nthr);   // xb complete
            // hat rows: x = h (plain solve) or T_A h + sum_b T_b x_b + gamma f, for my hats
            if (tid < nhl * RB) {
                const int u = tid / RB, rr = tid % RB, i = hA + u;
                const int64_t rw = (int64_t)row0 + rr;
                const double *hrow = HF + rr * HSR;       // hat h of this row: hrow[posT(h + 1)]
                const double hc = hrow[posT(i + 1)];
                double out = hc;
                if (mul) {
                    const double hp = hrow[posT(i)], hn = hrow[posT(i + 2)];
                    double acc = tAd[u] * hc;
                    acc = fma(tAu[u + 1], hn, acc);   // zero beyond the last hat
                    acc = fma(tAu[u], hp, acc);       // zero in front of the first
#pragma unroll
                    for (int tt = 0; tt < 3; ++tt) {
                        const int el = u + 1 - tt, kk = k0 + el;   // slot tt of element kk couples to hat i
                        if (kk < 0 || kk >= m) continue;
                        if (el >= 0 && el < mlc) {
                            for (int b = 0; b < P.nbT; ++b) acc = fma(sT[(b * 3 + tt) * mlc + el], xb[(b * mlc + el) * RB + rr], acc);
                        } else {
                            // an edge element of a neighbour CTA: its x_b follow from its backward-swept z_b (in the
                            // neighbour's message) and the hats, exactly as its owner computes them
                            const int nbr = el < 0 ? rank - 1 : rank + 1, side = el < 0 ? 1 : 0;
                            const double *zz = mbp + nbr * mbw + (mlc + 2) * RB + side * R2_NBMAX * RB + rr;
                            const double g0 = hrow[posT(kk)], g1 = hrow[posT(kk + 1)], g2 = hrow[posT(kk + 2)];
                            double xs[R2_NBMAX];
                            unsigned long long zv[R2_NBMAX];
#pragma unroll
                            for (int b = 0; b < R2_NBMAX; ++b) zv[b] = mbmode ? r2_mbpeek(zz + b * RB) : 0ull;   // all four requested at once
#pragma unroll
                            for (int b = 0; b < R2_NBMAX; ++b) {
                                double v = mbmode ? r2_mbtake(zz + b * RB, zv[b], mbmode, P.merr) : zz[b * RB];
                                {   // couplings of the neighbour's edge element (zero beyond P.nb), cached at set-up
                                    const double *sg = sEdge + (side * R2_NBMAX + b) * 3;
                                    v = fma(-sg[0], g0, v);
                                    v = fma(-sg[1], g1, v);
                                    v = fma(-sg[2], g2, v);
                                }
                                if (b >= 2) v = fma(-tab[b - 2].y, xs[b - 2], v);
                                xs[b] = v * tab[b].x;
                            }
#pragma unroll
                            for (int b = 0; b < R2_NBMAX; ++b)
                                acc = fma(tEdge[(side * R2_NBMAX + b) * 3 + tt], xs[b], acc);
                        }
                    }
                    out = useF ? fma(P.gamma, fh[u * RB + rr], acc) : acc;
                }
                if (rw < P.nrows) P.X[rw + (int64_t)(i - P.kbase) * P.ld] = out;   // hats come first in the panel
            }
            named_bar_sync(R2_BAR_COMPUTE, nthr);   // HF, ct / xb and the mailboxes of this parity are rewritten by later blocks
            R2_STAMP(9);
        }
      }
    }
    __syncthreads();
    if (cs > 1) {   // no CTA exits while a peer may still send into its shared memory
        cluster_arrive();
        cluster_wait();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_t r2_encoder() {
    static encode_tiled_t fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_t>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

static int r2_env(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// bubbles of a panel as (row, element, degree), hats as (row, hat); boxes (8, mlc, 8) and (8, mlc + 1)
static bool r2_make_maps(encode_tiled_t enc, const double *base, int64_t ld, int64_t nrows, int m, int nh, int q, int mlc, int dc, int rb, CUtensorMap *m3, CUtensorMap *mh) {
    const CUtensorMapL2promotion promo = r2_env("POP_ROW2_L2PROMO", 0) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    {
        cuuint64_t dims[3] = {(cuuint64_t)nrows, (cuuint64_t)m, (cuuint64_t)q};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * 8 * (cuuint64_t)m};
        cuuint32_t box[3] = {(cuuint32_t)rb, (cuuint32_t)mlc, (cuuint32_t)dc};
        cuuint32_t es[3] = {1, 1, 1};
        if (enc(m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)(base + (int64_t)nh * ld), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)nrows, (cuuint64_t)std::max(nh, 1)};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
        cuuint32_t box[2] = {(cuuint32_t)rb, (cuuint32_t)(mlc + 1)};
        cuuint32_t es[2] = {1, 1};
        if (enc(mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

typedef void (*r2_kernel_t)(const R2Params);
template <bool REM>
static r2_kernel_t r2_kernel_t8(int ch, int dc, int mlc) {   // 8 rows per block
    if (mlc == 32 && ch == 32) return dc == 8 ? k_row2<32, 8, 32, 8, REM> : k_row2<32, 4, 32, 8, REM>;   // the ADI shapes (m = 128: cluster of 4, m = 256: of 8)
    switch (ch) {
        case 8: return dc == 8 ? k_row2<8, 8, 0, 8, REM> : k_row2<8, 4, 0, 8, REM>;
        case 16: return dc == 8 ? k_row2<16, 8, 0, 8, REM> : k_row2<16, 4, 0, 8, REM>;
        case 24: return dc == 8 ? k_row2<24, 8, 0, 8, REM> : k_row2<24, 4, 0, 8, REM>;
        case 32: return dc == 8 ? k_row2<32, 8, 0, 8, REM> : k_row2<32, 4, 0, 8, REM>;
    }
    return nullptr;
}
static r2_kernel_t r2_kernel(int ch, int dc, int mlc, int rb, bool rem) {
    if (!rem) return rb == 8 ? r2_kernel_t8<false>(ch, dc, mlc) : nullptr;   // cluster mode walks 8-row blocks only
    if (rb == 16) {   // panels of few elements per CTA (8 GPUs: 16 elements each): 16 rows per block
        if (mlc == 16 && ch == 32) return dc == 8 ? k_row2<32, 8, 16, 16, true> : k_row2<32, 4, 16, 16, true>;
        switch (ch) {
            case 8: return dc == 8 ? k_row2<8, 8, 0, 16, true> : k_row2<8, 4, 0, 16, true>;
            case 16: return dc == 8 ? k_row2<16, 8, 0, 16, true> : k_row2<16, 4, 0, 16, true>;
            case 24: return dc == 8 ? k_row2<24, 8, 0, 16, true> : k_row2<24, 4, 0, 16, true>;
            case 32: return dc == 8 ? k_row2<32, 8, 0, 16, true> : k_row2<32, 4, 0, 16, true>;
        }
        return nullptr;
    }
    return r2_kernel_t8<true>(ch, dc, mlc);
}

static int r2_align(int x, int a) { return (x + a - 1) / a * a; }

// Remote mode plan for a decomposition of m elements over `world` GPUs: CTAs per row block and GPU (csv), elements per
// CTA (mlc), doubles per message.  false: the shape is not covered (the two-pass kernels of pop_row.cu serve it).
bool pop_row2_remote_plan(int m, int world, int *csv, int *mlc, int *rb, int *mbw) {
    if (world < 1 || m % world) return false;
    const int ml = m / world;
    for (int c = 1; c <= 8; c <<= 1) {
        if (ml % c) continue;
        const int e = ml / c;
        if (e < 2 || (e & 1) || 16 * e > 512) continue;
        *csv = c;
        *mlc = e;
        *rb = (e <= 16 && r2_env("POP_ROW2_RB", 16) == 16) ? 16 : 8;   // 2 * rb * e <= 512 compute threads
        *mbw = (e + 2) * *rb + 2 * R2_NBMAX * *rb;
        return true;
    }
    return false;
}

// POP_OK, an error, or POP_ROW1_UNSUPPORTED (the caller falls back to the older row kernels)
int pop_row2_half_step(pop_ctx *ctx, const RowOp &op, const double *F, double *R, const Row2Remote *rem) {
    if ((op.world != 1 && !rem) || r2_env("POP_ROW2", 1) == 0) return POP_ROW1_UNSUPPORTED;
    const bool strict = r2_env("POP_ROW2_STRICT", 0) != 0;
    auto unsupported = [&](const char *why) -> int {
        if (strict) {
            pop_set_error("pop_row2: %s (m=%d q=%d nh=%d)", why, op.m, op.q, op.nh);
            return POP_ERR_DIM;
        }
        return POP_ROW1_UNSUPPORTED;
    };
    if (op.has1) return unsupported("the first super-diagonal of a D block is not zero");
    if (op.q < 2 || op.q > 64 || op.nb > R2_NBMAX || op.nbT > R2_NBMAX) return unsupported("chain length / coupling blocks out of range");
    if ((op.ld & 1) || ((uintptr_t)R & 15)) return unsupported("panel not 16-byte aligned with an even leading dimension");
    const bool useF = op.nbT >= 0 && F != nullptr && op.gamma != 0.0;
    if (useF && ((op.ldf & 1) || ((uintptr_t)F & 15))) return unsupported("F not 16-byte aligned with an even leading dimension");
    encode_tiled_t enc = r2_encoder();
    if (!enc) return unsupported("cuTensorMapEncodeTiled is not available");
    const int ch = op.q <= 16 ? 8 : op.q <= 32 ? 16 : op.q <= 48 ? 24 : 32;
    const int want_cs = r2_env("POP_ROW2_CS", 0);
    int cs = 0, mlc = 0;
    int rb = R2_RB;
    if (rem) {
        // several GPUs (or forced): no cluster, the plan was fixed when the mailboxes were allocated
        if (op.ml * op.world != op.m || op.k0 != op.rank * op.ml || rem->csv * rem->mlc != op.ml) return unsupported("uneven element decomposition");
        cs = 1;
        mlc = rem->mlc;
        rb = rem->rb;
    } else {
        for (int c = 1; c <= R2_MAXCS; c <<= 1) {
            if (want_cs && c != want_cs) continue;
            if (op.m % c) continue;
            const int e = op.m / c;
            if (e < 1 || 16 * e > 512) continue;
            if (c > 1 && e < 2) continue;
            cs = c;
            break;
        }
        if (!cs) return unsupported("no cluster size with m % cs == 0 and 16 * m / cs <= 512 threads");
        mlc = op.m / cs;
    }
    if (mlc + 1 > 256) return unsupported("hat box too wide");
    const int nthr = 2 * rb * mlc;
    if (nthr % 32) return unsupported("elements per CTA must be even");   // whole warps: the named barrier and the hat scan need them
    R2Params P;
    memset(&P, 0, sizeof(P));
    P.mbw = (mlc + 2) * rb + 2 * R2_NBMAX * rb;
    const int nseg = 64 / rb;                                // scan segments per row (two hat warps share the rows)
    const int lp = (((op.nh + nseg - 1) / nseg) + 3) & ~3;  // hats per scan segment, a multiple of 4
    // warp-per-row hat scan by the compute warps: needs a warp per row and at most 8 hats per lane
    const int hat_L = (std::max(op.nh, op.m) + 3 + 31) / 32;   // 32 L entries of a row: the pad of hat -1, the hats, at least one pad behind (hats up to m + 1 are read)
    // default: remote mode only (measured on one box: cluster mode 71.1 ms per ADI solve with the two hat warps, 74.5 with the
    // compute warps scanning; remote mode 8-GPU shape 15.4 against 15.25)
    const int hat_mode = (rem && nthr / 32 >= rb && hat_L <= 16 && r2_env("POP_ROW2_HATMODE", 1)) ? 1 : 0;
    int hfs = std::max(hat_mode ? 32 * hat_L : nseg * lp, op.m + 2) + 4;   // zero entries in front of hat 0 and behind the last hat
    if (hat_mode) {
        hfs = (hfs + 1) & ~1;
        if (!((hfs / 2) & 1)) hfs += 2;                                     // even with hfs / 2 odd: the rows start in different banks
    }
    // (shared memory beyond 196 KB costs the next carve-out step: 28 KB of L1 instead of 60 KB, and the spills of the compute
    // warps live there -- measured 71.1 -> 74.5 ms per ADI solve on one GPU when these tables pushed the CTA over it)
    const int ks_bytes = hat_mode ? (10 + 3 * hat_L) * 32 * 8 : 0;
    // behind the ring: HF, the padded hat tables, the scan tables, the mailboxes
    const int tail = r2_align(hfs * rb * 8, 128) + 2 * r2_align(nseg * lp * 8, 128) + r2_align(3 * nseg * lp * 8, 128) + r2_align(ks_bytes, 128) + r2_align(hfs * 4, 128) +
                     r2_align(2 * cs * P.mbw * 8, 128);
    (void)want_cs;
    // degrees per chunk (TMA box, region of the row block buffer): 4 keeps the hand-over between the blocks fine-grained
    // (16-row blocks of the remote mode: 8, measured 12.50 -> 12.22 ms per ADI solve at the 8-GPU shape; 8-row blocks: 4, 8 costs 3-7 %)
    const int dc = r2_env("POP_ROW2_DC", (rem && rb == 16) ? 8 : 4) == 8 ? 8 : 4;
    const int nch = (op.q + dc - 1) / dc;
    const R2Fixed LF = r2_fixed(ch, dc, mlc, rb);
    if (LF.ring + tail > ctx->max_smem_optin) return unsupported("shared memory too small for the row block");
    if (4 + 4 * ((2 * ch + dc - 1) / dc) > 128) return unsupported("too many barriers");
    int o = LF.ring;
    P.o_hf = o; o += r2_align(hfs * rb * 8, 128);
    P.o_hrd = o; o += r2_align(nseg * lp * 8, 128);
    P.o_hcd = o; o += r2_align(nseg * lp * 8, 128);
    P.o_hcu = o; o += r2_align(3 * nseg * lp * 8, 128);   // hcu, hwd, hwu
    P.o_ks = o; o += r2_align(ks_bytes, 128);
    P.o_pos = o; o += r2_align(hfs * 4, 128);
    P.o_mb = o; o += r2_align(2 * cs * P.mbw * 8, 128);
    P.total = o;
    if ((size_t)P.total > (size_t)ctx->max_smem_optin) return unsupported("shared memory too small");
    P.lp = lp;
    P.hfs = hfs;
    P.hat_mode = hat_mode;
    P.hat_L = hat_L;
    P.f_evict_first = r2_env("POP_F_EVICT_FIRST", 1);
    // measured on 8 real GPUs, same box (profiles/r2z_dd8_senders.txt): 12.85 ms per ADI solve with the compute warps sending, 13.25 with the
    // sender warps (the one-GPU stand-in, where a message is an L2 round trip, had shown the opposite: 12.08 -> 11.9): off
    P.use_senders = r2_env("POP_ROW2_SENDERS", 0);
    P.pf_dist = r2_env("POP_ROW2_PF", 0);   // L2 prefetch distance of F in row blocks (0: off)

    const int pm = rem ? op.ml : op.m, pnh = rem ? op.nhl : op.nh;   // the panel: all elements, or this GPU's
    if (!r2_make_maps(enc, R, op.ld, op.nrows, pm, pnh, op.q, mlc, dc, rb, &P.mapX3, &P.mapXh)) return unsupported("cuTensorMapEncodeTiled rejected the panel");
    if (useF) {
        if (!r2_make_maps(enc, F, op.ldf, op.nrows, pm, pnh, op.q, mlc, dc, rb, &P.mapF3, &P.mapFh)) return unsupported("cuTensorMapEncodeTiled rejected F");
    } else {
        P.mapF3 = P.mapX3;
        P.mapFh = P.mapXh;
    }
    P.rd = op.rd; P.w2 = op.w2; P.S = op.S; P.rdA = op.rdA; P.offA = op.offA; P.nb = op.nb;
    P.t0 = op.t0; P.t2 = op.t2; P.Ts = op.Ts; P.TAd = op.TAd; P.TAu = op.TAu; P.nbT = op.nbT;
    P.gamma = op.gamma;
    P.useF = useF ? 1 : 0;
    P.X = R; P.F = F; P.ld = op.ld; P.ldf = op.ldf; P.nrows = op.nrows;
    P.m = op.m; P.nh = op.nh; P.q = op.q;
    P.cs = cs; P.mlc = mlc; P.nch = nch;
    P.nblk = (int)((op.nrows + rb - 1) / rb);
    if (rem) {
        P.remote = rem->standin ? 2 : 1; P.world = rem->world; P.me = rem->me; P.csv = rem->csv; P.nr = rem->world * rem->csv;
        P.kbase = op.k0;
        P.epoch = rem->epoch;
        for (int g = 0; g < rem->world; ++g) P.mbox[g] = rem->mbox[g];
        P.merr = rem->merr;
        if ((int64_t)2 * P.nblk * P.csv * P.nr * P.mbw > rem->mbox_doubles) return unsupported("mailbox too small");
        if (P.nr > 2 * mlc) return unsupported("more CTAs per row block than the edge buffer holds");   // HE aliases the coupling buffer [4][mlc][rb]
    }

    r2_kernel_t kern = r2_kernel(ch, dc, mlc, rb, rem != nullptr);
    if (!kern) return unsupported("no kernel for this chain class");
    POP_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P.total));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // set-up under the tail of the previous kernel (pdl_wait in the kernel)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(640);   // 512 compute slots (16 mlc of them work) + the TMA producer warp group
    cfg.dynamicSmemBytes = (size_t)P.total;
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;   // the occupancy query wants the cluster attribute alone
    static std::mutex plan_mu;
    static int cached_key[5] = {0, 0, 0, 0, -1}, cached_ncl = 0;
    const int key[5] = {ch * 16 + dc + (mlc == 32 ? 1024 : 0) + (rem ? 4096 : 0) + (rb == 16 ? 8192 : 0), rem ? rem->csv : cs, nthr, P.total, ctx->device};
    int ncl;
    {
        std::lock_guard<std::mutex> lk(plan_mu);
        ncl = cached_ncl;
        if (!std::equal(key, key + 5, cached_key) || ncl <= 0) {
            int nc = 0;
            if (rem) {
                // every CTA of the grid must be resident: the CTAs of a row block wait for each other through the mailboxes
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, 640, (size_t)P.total) != cudaSuccess || per_sm < 1) {
                    cudaGetLastError();
                    return unsupported("kernel does not fit an SM");
                }
                nc = ctx->sm_count / rem->csv;   // one CTA per SM, the same grid on every GPU
            } else {
                cfg.gridDim = dim3((unsigned)(cs * ctx->sm_count));
                if (cudaOccupancyMaxActiveClusters(&nc, (const void *)kern, &cfg) != cudaSuccess || nc <= 0) {
                    cudaGetLastError();
                    return unsupported("no co-resident cluster");
                }
            }
            ncl = nc;
            std::copy(key, key + 5, cached_key);
            cached_ncl = ncl;
        }
    }
    if (const int cap = r2_env("POP_ROW2_NCL", 0)) ncl = std::min(ncl, cap);
    ncl = std::min(ncl, P.nblk);
    P.pairing = (rb == 8 && ncl >= 2 && r2_env("POP_ROW2_PAIR", 1)) ? 1 : 0;   // 16 rows are whole 128 B lines already
    if (P.pairing) ncl &= ~1;
    cfg.gridDim = dim3((unsigned)(ncl * (rem ? rem->csv : cs)));
    if (r2_env("POP_ROW2_VERBOSE", 0))
        fprintf(stderr, "pop_row2: chain %d per thread, cluster %d x %d threads, %d B smem, chunks of %d degrees, %d clusters, %d row blocks, F %d\n", ch, cs, nthr, P.total, dc, ncl,
                P.nblk, P.useF);
    long long *dbg = nullptr;
    if (r2_env("POP_ROW2_TRACE", 0)) {
        POP_CUDA(cudaMalloc(&dbg, 64 * 16 * sizeof(long long)));
        POP_CUDA(cudaMemset(dbg, 0, 64 * 16 * sizeof(long long)));
        P.dbg = dbg;
    }
    cfg.numAttrs = r2_env("POP_PDL", 0) ? 2 : 1;   // measured: programmatic dependent launch costs 5 % here (70.1 -> 74.2 ms, 12.1 -> 12.7 ms per ADI solve): off
    POP_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    ctx->launches++;
    if (dbg) {   // debugging aid: phase durations of CTA 0 in SM clock cycles
        long long h[64 * 16];
        POP_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        static const char *names[8] = {"wait X", "load+backward", "couplings", "message", "forward", "wait hats", "corrections", "chunks"};
        static const char *sub[5] = {"messages in", "rhs built", "barrier", "scan", "barrier"};   // stamps 10..14 inside "wait hats"
        double sum[10] = {0}, ssum[5] = {0};
        int nb = 0;
        for (int t = 2; t < 64 && h[t * 16 + 9]; ++t, ++nb) {
            for (int i = 0; i < 9; ++i) sum[i] += (double)(h[t * 16 + i + 1] - h[t * 16 + i]);
            long long prev = h[t * 16 + 5];
            for (int i = 0; i < 5; ++i)
                if (h[t * 16 + 10 + i]) {
                    ssum[i] += (double)(h[t * 16 + 10 + i] - prev);
                    prev = h[t * 16 + 10 + i];
                }
        }
        if (nb) {
            fprintf(stderr, "pop_row2 trace (CTA 0, thread 0, mean SM cycles over %d row blocks):", nb);
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %s %.0f |", names[i], sum[i] / nb);
            fprintf(stderr, " hat rows + end %.0f | block %.0f\n", sum[8] / nb, (double)(h[(2 + nb - 1) * 16 + 9] - h[2 * 16]) / nb);
            fprintf(stderr, "  inside wait hats:");
            for (int i = 0; i < 5; ++i) fprintf(stderr, " %s %.0f |", sub[i], ssum[i] / nb);
            fprintf(stderr, "\n");
        }
    }
    return POP_OK;
}
