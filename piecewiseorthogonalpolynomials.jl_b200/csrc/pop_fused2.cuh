// Kernel template of the register-resident fused solve (see pop_fused2.cu for the design notes).
// Instantiated per row class in pop_fused2_rNN.cu so the classes compile in parallel.
#pragma once
#include "pop_internal.h"
#include "pop_ptx.cuh"

#define F2_NBMAX 4        // coupling blocks B handled in registers
#define F2_BAR_BYTES 1152 // 144 mbarrier words at the start of shared memory
#define F2_MAXSLOTS 32
#define F2_MAXFSLOTS 16
#define F2_G 8            // rows per ring chunk
#define F2_CLASS 8        // a row class covers q in (ROWS*NPAR - F2_CLASS*NPAR, ROWS*NPAR]

struct F2Params {
#ifdef F2_PROF
    unsigned long long *prof;   // [2 observers][16 phases] accumulated SM cycles (debug build only)
#endif
    // factor U (upper data; D block shared by all elements)
    const double *rdD, *W1, *W2;   // [q] reciprocal pivots, first / second super-diagonal (may be null)
    const double *Sb;              // [nb][3][m] hat <-> bubble coupling
    const double *rdA, *offA;      // [nh], [nh-1]
    int m, nh, q, nb;
    // optional product operator T = Symmetric(upper data), D block shared
    const double *Td0, *Td1, *Td2; // [q] diagonal, first, second super-diagonal (may be null)
    const double *Trow;            // [nbT][3][m]
    const double *TAd, *TAu;       // [nh], [nh-1]
    int nbT;
    double gamma;
    const double *F;
    int64_t ldf;
    // panel
    double *X;
    int64_t ldx, nrhs;
    // plan
    int f_evict_first;     // F is a stream without reuse: its loads carry the L2 evict_first policy (the panel then survives in L2 between half steps when it fits)
    int stagger_ns;        // start-up offset between clusters (de-phases the clusters' memory bursts)
    int cs, mloc, rp, nch, nslots, slot_stride, ncE, contiguous, hcap, nf;   // nf: slots of the F ring (fused product)
    int o_tab, o_tab1, o_tT, o_hrd, o_hoff, o_H, o_ct, o_xb, o_inb, o_wagg, o_hw, o_ring, o_fring;   // offsets in doubles
    int o_kst, hat_L;      // single-CTA columns: tables of the one-warp hat scan, [10][32] step multipliers + cd, cu, rd as [hat_L][32]
};

typedef void (*f2_kernel_t)(const F2Params);
// variant: 0 -> no D coupling, 1 -> second super-diagonal only (even/odd chains decoupled), 2 -> general (<= 2 bands)
f2_kernel_t f2_kernel_r8(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r16(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r24(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r32(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r40(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r48(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r56(int npar, int variant, bool mul, bool cl);
f2_kernel_t f2_kernel_r64(int npar, int variant, bool mul, bool cl);
// largest block of a row class (registers: the chain of an element lives in 2*ROWS registers)
__host__ __device__ constexpr int f2_tmax(int rows) { return rows <= 40 ? 512 : rows <= 48 ? 384 : 256; }

// layout of the small cluster mailboxes (doubles, relative to o_inb)
#define F2_INB_A 0     // [2 buffers][2 sides][2 parities]  boundary couplings of the neighbours' edge elements
#define F2_INB_B 8     // [2 buffers][8 ranks][2]           zero-incoming segment maps of the downward / upward hat sweep (16 B pairs)
#define F2_INB_P 40    // [3][8 ranks]                      constants of the factor: multipliers down / up, response of the up map to the value entering the down sweep
#define F2_INB_SIZE 64
#define F2_WAGG 192    // scratch of the scan warps: [0,64) down, [64,128) up, [128,192) reductions

#ifdef __CUDACC__
// The hats are DISTRIBUTED over the cluster: CTA r owns hats [hA,hB) (hA = first element of the CTA).
//   down: y[i] = (H[i] - off[i] y[i+1]) rd[i], i = nh-1..0   (U solve,   src/arrowhead.jl:453)
//   up  : h[i] = (y[i] - off[i-1] h[i-1]) rd[i], i = 0..nh-1 (U' solve,  src/arrowhead.jl:472)
// Each is a segmented affine scan: pass 1 reduces a thread's hats to the map  out = t + Pm * in ;
// shuffles compose the maps over a warp and over the warp aggregates; the CTA's map (its additive
// part; the multipliers are constants of the factor) goes to the peers that need it through
// st.async + mbarrier; pass 2 is the reference recurrence itself with the exact incoming value.
// hoffl[u] couples local hats u-1 and u (hoffl[0]: the left neighbour's last hat).
// Returns the value that entered this CTA's segment.
#ifndef F2_EARLY_COEF
#define F2_EARLY_COEF 0   // 1: load the correction coefficients before the hat solve (measured: the 24 extra live registers spill, -10 %)
#endif
#ifndef F2_STORE_NA
#define F2_STORE_NA 0   // 1: result stores bypass L1 allocation (measured: no gain for cluster columns, -4 % for single-CTA columns)
#endif
#if F2_STORE_NA
#define F2_STORE(p, v) stg_na_f64((p), (v))
#else
#define F2_STORE(p, v) (*(p) = (v))
#endif
#define F2_SCAN_WARPS 4   // warps that run the hat scans (the others wait at the block barrier)
#ifndef F2_OLD_SCAN
#define F2_OLD_SCAN 1     // 0: single-CTA columns solve their hats on ONE warp (Kogge-Stone over the lanes with tabulated multipliers, as k_row2 does);
                          // measured equal (0.314 ms per half step, 70.0 / 12.1 ms per ADI solve at the 1- / 8-GPU shapes) and parity-green: the hat
                          // stage of a column is covered by the other resident CTA, so the generic segmented scan stays
#endif
// Pass 1 + composition over the scan warps: thread-local map (t, Pm), then (at, aP) = composite of the scan warps
// from this lane's warp on (lanes < nsw hold warp `lane`'s suffix / prefix composite).
template <bool BW>
__device__ __forceinline__ void f2_scan_local(const double *Hs, const double *hrd, const double *hoffl, const int lo, const int hi, const int lane,
                                              const int warp, const int nsw, double *wagg, double &t, double &Pm, double &at, double &aP) {
    t = 0.0;
    Pm = 1.0;
    for (int ii = lo; ii < hi; ++ii) {
        const int u = BW ? (hi - 1 - (ii - lo)) : ii;
        const double off = BW ? hoffl[u + 1] : hoffl[u];
        const double rd = hrd[u];
        t = (Hs[u] - off * t) * rd;
        Pm = -off * rd * Pm;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t2 = BW ? __shfl_down_sync(POP_FULL_MASK, t, o) : __shfl_up_sync(POP_FULL_MASK, t, o);
        const double P2 = BW ? __shfl_down_sync(POP_FULL_MASK, Pm, o) : __shfl_up_sync(POP_FULL_MASK, Pm, o);
        const bool ok = BW ? (lane + o < 32) : (lane - o >= 0);
        if (ok) {
            t = fma(Pm, t2, t);
            Pm = Pm * P2;
        }
    }
    // lane now holds the composite of lanes lane..31 (BW) / 0..lane (forward)
    at = t;
    aP = Pm;
    if (nsw > 1) {
        if (lane == (BW ? 0 : 31)) {
            wagg[warp] = t;
            wagg[32 + warp] = Pm;
        }
        named_bar_sync(1, nsw * 32);
        // compose the warp aggregates (every scan warp redundantly, by shuffles)
        at = lane < nsw ? wagg[lane] : 0.0;
        aP = lane < nsw ? wagg[32 + lane] : 1.0;
#pragma unroll
        for (int o = 1; o < F2_SCAN_WARPS; o <<= 1) {
            const double t2 = BW ? __shfl_down_sync(POP_FULL_MASK, at, o) : __shfl_up_sync(POP_FULL_MASK, at, o);
            const double P2 = BW ? __shfl_down_sync(POP_FULL_MASK, aP, o) : __shfl_up_sync(POP_FULL_MASK, aP, o);
            const bool ok = BW ? (lane + o < 32) : (lane - o >= 0);
            if (ok) {
                at = fma(aP, t2, at);
                aP = aP * P2;
            }
        }
    }
}

// Pass 2: the reference recurrence itself, started from the exact value Gc that enters this CTA's segment.
template <bool BW>
__device__ __forceinline__ void f2_scan_finish(double *Hs, const double *hrd, const double *hoffl, const int lo, const int hi, const int lane, const int warp,
                                               const int nsw, const double t, const double Pm, const double at, const double aP, const double Gc,
                                               double &hlast) {
    // value entering this warp: composite of the scan warps after (BW) / before (forward) it applied to Gc
    double Gin = Gc;
    if (nsw > 1) {
        const int src = BW ? warp + 1 : warp - 1;
        const double ats = __shfl_sync(POP_FULL_MASK, at, src & 31), aPs = __shfl_sync(POP_FULL_MASK, aP, src & 31);
        if (BW ? (warp + 1 < nsw) : (warp > 0)) Gin = fma(aPs, Gc, ats);
    }
    const double tn = BW ? __shfl_down_sync(POP_FULL_MASK, t, 1) : __shfl_up_sync(POP_FULL_MASK, t, 1);
    const double Pn = BW ? __shfl_down_sync(POP_FULL_MASK, Pm, 1) : __shfl_up_sync(POP_FULL_MASK, Pm, 1);
    double hin = (BW ? (lane == 31) : (lane == 0)) ? Gin : fma(Pn, Gin, tn);
    for (int ii = lo; ii < hi; ++ii) {
        const int u = BW ? (hi - 1 - (ii - lo)) : ii;
        const double off = BW ? hoffl[u + 1] : hoffl[u];
        hin = (Hs[u] - off * hin) * hrd[u];
        Hs[u] = hin;
    }
    hlast = hin;   // the last hat this thread wrote
}

// Set-up helper: in-place inclusive scan of the affine maps x -> A[u] x + B[u] in shared memory (Hillis-Steele,
// all threads of the block): afterwards (A[u], B[u]) = map_u o ... o map_0, or map_u o ... o map_{n-1} if REV.
template <bool REV>
__device__ __forceinline__ void f2_setup_scan(double *A, double *B, double *tA, double *tB, const int n, const int tid, const int nthr) {
    double *sa = A, *sb = B, *da = tA, *db = tB;
    for (int d = 1; d < n; d <<= 1) {
        for (int u = tid; u < n; u += nthr) {
            const int v = REV ? u + d : u - d;
            double a = sa[u], b = sb[u];
            if (v >= 0 && v < n) {
                b = fma(a, sb[v], b);
                a = a * sa[v];
            }
            da[u] = a;
            db[u] = b;
        }
        __syncthreads();
        double *x = sa; sa = da; da = x;
        x = sb; sb = db; db = x;
    }
    if (sa != A) {
        for (int u = tid; u < n; u += nthr) {
            A[u] = sa[u];
            B[u] = sb[u];
        }
    }
    __syncthreads();
}

// TMA loads of one chunk (rows j0 .. j0+nr-1 of column xcol) into its ring slot; one warp.
__device__ __forceinline__ void f2_issue_chunk(const F2Params &P, uint64_t *full, double *ring, const int lane, const int k0, const int ml,
                                               const double *xcol, const int chunk_from_top, const int slot) {
    const int ch = P.nch - 1 - chunk_from_top;
    const int j0 = ch * F2_G;
    const int nr = min(F2_G, P.q - j0);
    double *dst = ring + (size_t)slot * P.slot_stride;
    if (nr <= 0) {   // a chunk above the last row (wide classes of the even/odd split): nothing to load
        if (lane == 0) mbar_expect_tx(full + slot, 0);
    } else if (P.contiguous) {
        if (lane == 0) {
            const int64_t gs = (int64_t)P.nh + (int64_t)j0 * P.m;
            const int64_t a0 = gs & ~(int64_t)1, a1 = (gs + (int64_t)nr * P.m + 1) & ~(int64_t)1;
            const uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
            mbar_expect_tx(full + slot, bytes);
            tma_load_1d(dst, xcol + a0, bytes, full + slot);
        }
    } else {
        const int64_t g0 = (int64_t)P.nh + (int64_t)(j0 + lane) * P.m + k0;
        const int64_t a0 = g0 & ~(int64_t)1, a1 = (g0 + ml + 1) & ~(int64_t)1;
        const uint32_t bytes = lane < nr ? (uint32_t)(a1 - a0) * 8u : 0u;
        const uint32_t total = __reduce_add_sync(POP_FULL_MASK, bytes);
        if (lane == 0) mbar_expect_tx(full + slot, total);
        __syncwarp();
        if (lane < nr) tma_load_1d(dst + (size_t)lane * P.rp, xcol + a0, bytes, full + slot);
    }
}

// TMA load of F chunk `ch` (rows 8*ch .., ascending order) of column fcol into F-ring slot `slot`; one thread.
__device__ __forceinline__ void f2_issue_fchunk(const F2Params &P, uint64_t *ffull, double *fring, const double *fcol, const int ch, const int slot) {
    const int j0 = ch * F2_G;
    const int nr = min(F2_G, P.q - j0);
    if (nr <= 0) {
        mbar_expect_tx(ffull + slot, 0);
        return;
    }
    const int64_t gs = (int64_t)P.nh + (int64_t)j0 * P.m;
    const int64_t a0 = gs & ~(int64_t)1, a1 = (gs + (int64_t)nr * P.m + 1) & ~(int64_t)1;
    const uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
    mbar_expect_tx(ffull + slot, bytes);
    if (P.f_evict_first)
        tma_load_1d_hint(fring + (size_t)slot * P.slot_stride, fcol + a0, bytes, ffull + slot, l2_policy_evict_first());
    else
        tma_load_1d(fring + (size_t)slot * P.slot_stride, fcol + a0, bytes, ffull + slot);
}

// CL: the kernel may run as a thread-block cluster (a column spread over cs CTAs).  Columns that fit one CTA get a kernel of
// their own without the cluster set-up, mailboxes and segment-map exchange: a third less SASS to stream through the
// instruction cache per column (profiles/r2_fused2_dd8_ncu.txt: no_instruction is the top stall of the short ADI panels).
template <int ROWS, int NPAR, int ND, bool SKIP1, bool MUL, bool CL>
__global__ void __launch_bounds__(f2_tmax(ROWS), 1) k_fused2(const __grid_constant__ F2Params P) {
    static_assert(!(MUL && CL), "the fused product runs on single-CTA columns");
    static_assert(NPAR == 1 || (ND != 1 && (ND == 0 || SKIP1)), "the even/odd split needs decoupled chains");
    static_assert(!(MUL && NPAR == 2 && !(ND == 0 || SKIP1)), "fused product with split chains: T must not couple j and j+1");
    static_assert(ROWS % F2_G == 0, "row classes are multiples of the chunk height");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int JMAX = ROWS * NPAR;
    constexpr int G = F2_G;
    constexpr int NCH = JMAX / G;                    // chunks per column of this class (the top ones may be partial)
    constexpr int JSAFE = JMAX - F2_CLASS * NPAR;    // rows j <= JSAFE exist for every q of the class
    constexpr int NBI = (F2_NBMAX + NPAR - 1) / NPAR;

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *full = bars, *hfull = bars + F2_MAXSLOTS, *hxA = hfull + 2, *hxB = hfull + 4, *hxC = hfull + 6;
    uint64_t *ffull = bars + 48, *fempty = bars + 64;   // F ring of the fused product
    double *sm = reinterpret_cast<double *>(smem_raw + F2_BAR_BYTES);
    double2 *tab = reinterpret_cast<double2 *>(sm + P.o_tab);   // (rd_j, w2_j), zero beyond q
    double *tab1 = sm + P.o_tab1;                               // w1_j
    double *tT = sm + P.o_tT;                                   // [3][JMAX+2] product bands
    double *hrd = sm + P.o_hrd;                                 // [nhl+1] reciprocal pivots of my hats (+ the right neighbour's first)
    double *hoffl = sm + P.o_hoff;                              // [nhl+1] hoffl[u] couples local hats u-1 and u
    double *Hb = sm + P.o_H;
    double *ct = sm + P.o_ct;
    double *xb = sm + P.o_xb;                                   // x of the coupled bubbles (fused product)
    double *inb = sm + P.o_inb;
    double *wagg = sm + P.o_wagg;
    double *hwd = sm + P.o_hw, *hwu = hwd + P.hcap;   // cs > 1: weights of the segment maps (see the set-up below)
    double *ring = sm + P.o_ring;
    double *fring = sm + P.o_fring;
    double *kst = sm + P.o_kst;   // CL == false only

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    const int cs = CL ? P.cs : 1;
    const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int m = P.m, nh = P.nh, q = P.q, nb = P.nb;
    const int ncE = P.ncE;
    const int k0 = rank * P.mloc;
    const int ml = min(m, k0 + P.mloc) - k0;
    const int hA = k0, hB = rank == cs - 1 ? nh : min(nh, k0 + ml);   // my hats
    const int nhl = max(hB - hA, 0);
    const int nslots = P.nslots;
    const int ncl = gridDim.x / cs;
    const int c_first = blockIdx.x / cs;
    const int ncols = c_first < P.nrhs ? (int)((P.nrhs - c_first + ncl - 1) / ncl) : 0;   // columns of this cluster
    const int hA2 = hA & ~1;
    const uint32_t hbytes = (uint32_t)(((hB + 1) & ~1) - hA2) * 8u;
    const int64_t colstep = (int64_t)ncl * P.ldx;

    // ---- one-time setup ----------------------------------------------------------------------
    pdl_launch_dependents();   // the next kernel may set itself up while this one drains
    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) mbar_init(full + s, 1);
        if (MUL)
            for (int s = 0; s < P.nf; ++s) {
                mbar_init(ffull + s, 1);
                mbar_init(fempty + s, nw);
            }
        for (int s = 0; s < 2; ++s) {
            mbar_init(hfull + s, 1);
            mbar_init(hxA + s, 1);
            mbar_init(hxB + s, 1);
            mbar_init(hxC + s, 1);
        }
        fence_mbar_init();
    }
    for (int j = tid; j < JMAX + 2; j += nthr) {
        const bool in = j < q;
        tab[j] = make_double2(in ? P.rdD[j] : 0.0, (in && P.W2 && j + 2 < q) ? P.W2[j] : 0.0);
        tab1[j] = (in && P.W1 && j + 1 < q) ? P.W1[j] : 0.0;
        if (MUL) {
            tT[j] = in ? P.Td0[j] : 0.0;
            tT[(JMAX + 2) + j] = (in && P.Td1 && j + 1 < q) ? P.Td1[j] : 0.0;
            tT[2 * (JMAX + 2) + j] = (in && P.Td2 && j + 2 < q) ? P.Td2[j] : 0.0;
        }
    }
    for (int u = tid; u <= nhl; u += nthr) {
        const int i = hA + u;
        hrd[u] = i < nh ? P.rdA[i] : 0.0;
        hoffl[u] = (i >= 1 && i < nh) ? P.offA[i - 1] : 0.0;
    }
    // the hat scans run on the first nsw warps: scan thread t owns hats [lo,hi)
    constexpr bool KSCAN = !CL && !F2_OLD_SCAN;          // single-CTA columns: one warp solves the hats (tables below)
    const int nsw = KSCAN ? 1 : min(nw, F2_SCAN_WARPS);
    const int nst = nsw * 32;
    const int sg = (nhl + nst - 1) / nst;
    const int lo = tid < nst ? min(tid * sg, nhl) : 0, hi = tid < nst ? min(lo + sg, nhl) : 0;
    __syncthreads();
    if (KSCAN && warp == 0) {
        // One-warp hat solve of a single-CTA column: lane l owns the hats [l L, (l+1) L).  Both bidiagonal recurrences
        //   y_i = H_i rd_i - cd_i y_{i+1},  cd_i = off(i,i+1) rd_i      (src/arrowhead.jl:453)
        //   h_i = y_i rd_i - cu_i h_{i-1},  cu_i = off(i-1,i) rd_i      (src/arrowhead.jl:472)
        // are affine in the incoming value with multipliers that depend on the factor only: they are tabulated here, together
        // with the products the Kogge-Stone steps over the lanes need (zero where a step has no partner).
        const int L = P.hat_L;
        double ad = 1.0, au = 1.0;
        for (int i = 0; i < L; ++i) {
            const int u = lane * L + i;
            const bool in = u < nhl;
            const double rdv = in ? hrd[u] : 0.0, cdv = in ? hoffl[u + 1] * rdv : 0.0, cuv = in ? hoffl[u] * rdv : 0.0;
            kst[(10 + i) * 32 + lane] = cdv;
            kst[(10 + L + i) * 32 + lane] = cuv;
            kst[(10 + 2 * L + i) * 32 + lane] = rdv;
            ad *= -cdv;
            au *= -cuv;
        }
#pragma unroll 1
        for (int st = 0; st < 5; ++st) {
            double pd = 1.0, pu = 1.0;
#pragma unroll 1
            for (int j = 0; j < (1 << st); ++j) {
                pd *= __shfl_sync(POP_FULL_MASK, ad, (lane + j) & 31);
                pu *= __shfl_sync(POP_FULL_MASK, au, (lane - j) & 31);
            }
            kst[st * 32 + lane] = (lane + (1 << st) < 32) ? pd : 0.0;
            kst[(5 + st) * 32 + lane] = (lane >= (1 << st)) ? pu : 0.0;
        }
    }
    if (cs > 1) {
        // multipliers of every rank's segment map (constants of the factor), shared through DSMEM
        double pd = 1.0, pu = 1.0;
        for (int u = lo; u < hi; ++u) {
            pd *= -hoffl[u + 1] * hrd[u];
            pu *= -hoffl[u] * hrd[u];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pd *= __shfl_xor_sync(POP_FULL_MASK, pd, o);
            pu *= __shfl_xor_sync(POP_FULL_MASK, pu, o);
        }
        if (lane == 0) {
            wagg[warp] = pd;
            wagg[32 + warp] = pu;
        }
        __syncthreads();
        // With zero incoming values both segment maps are LINEAR FUNCTIONALS of the hat right-hand sides H of
        // my segment (a_u = -off(u,u+1) rd_u, b_u = -off(u-1,u) rd_u):
        //   down:  y_0     = sum_k hwd[k] H[k] + pd * Gdn,   hwd[k] = rd_k prod_{u<k} a_u
        //   up:    h_{n-1} = sum_k wup[k] y_k  + pu * Gup,   wup[k] = rd_k prod_{u>k} b_u
        //          with y_k = y0_k + (prod_{u>=k} a_u) Gdn:   = sum_j hwu[j] H[j] + kappa * Gdn + pu * Gup,
        //          hwu[j] = rd_j c_j,  c_j = wup[j] + a_{j-1} c_{j-1},  kappa = sum_k wup[k] prod_{u>=k} a_u.
        // So ONE exchange of two dot products per column replaces the two dependent map exchanges.
        {
            const int n = nhl, hc = P.hcap;
            double *s0 = ring, *s1 = s0 + hc, *s2 = s1 + hc, *s3 = s2 + hc, *s4 = s3 + hc;
            for (int u = tid; u < n; u += nthr) {
                s0[u] = -hoffl[u + 1] * hrd[u];
                s1[u] = 0.0;
            }
            __syncthreads();
            f2_setup_scan<false>(s0, s1, s2, s3, n, tid, nthr);   // s0[k] = prod_{u<=k} a_u
            for (int u = tid; u < n; u += nthr) hwd[u] = hrd[u] * (u > 0 ? s0[u - 1] : 1.0);
            __syncthreads();
            for (int u = tid; u < n; u += nthr) {
                s0[u] = -hoffl[u] * hrd[u];
                s1[u] = 0.0;
            }
            __syncthreads();
            f2_setup_scan<true>(s0, s1, s2, s3, n, tid, nthr);    // s0[k] = prod_{u>=k} b_u
            for (int u = tid; u < n; u += nthr) s4[u] = hrd[u] * (u + 1 < n ? s0[u + 1] : 1.0);   // wup
            __syncthreads();
            for (int u = tid; u < n; u += nthr) {
                s0[u] = -hoffl[u + 1] * hrd[u];
                s1[u] = 0.0;
            }
            __syncthreads();
            f2_setup_scan<true>(s0, s1, s2, s3, n, tid, nthr);    // s0[k] = prod_{u>=k} a_u
            double kp = 0.0;
            for (int u = tid; u < n; u += nthr) kp = fma(s4[u], s0[u], kp);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kp += __shfl_xor_sync(POP_FULL_MASK, kp, o);
            if (lane == 0) wagg[128 + warp] = kp;
            __syncthreads();
            for (int u = tid; u < n; u += nthr) {
                s0[u] = u > 0 ? -hoffl[u] * hrd[u - 1] : 0.0;
                s1[u] = s4[u];
            }
            __syncthreads();
            f2_setup_scan<false>(s0, s1, s2, s3, n, tid, nthr);   // s1[j] = c_j
            for (int u = tid; u < n; u += nthr) hwu[u] = hrd[u] * s1[u];
            fence_proxy_async();   // the scratch lives in the ring: order these writes before the first TMA refill
        }
        if (tid == 0) {
            pd = 1.0;
            pu = 1.0;
            for (int w = 0; w < nsw; ++w) {
                pd *= wagg[w];
                pu *= wagg[32 + w];
            }
            double kp = 0.0;
            for (int w = 0; w < nw; ++w) kp += wagg[128 + w];
            for (int r = 0; r < cs; ++r) {
                st_cluster_f64(mapa_u32(inb + F2_INB_P + rank, r), pd);
                st_cluster_f64(mapa_u32(inb + F2_INB_P + 8 + rank, r), pu);
                st_cluster_f64(mapa_u32(inb + F2_INB_P + 16 + rank, r), kp);
            }
        }
        cluster_arrive();
        cluster_wait();
    }
    pdl_wait();   // everything above read the factor only; the panel (and F) belong to the previous kernel until here
    if (P.stagger_ns > 0) {
        // clusters would otherwise run in phase (same work, same start): all loading, then all in the hat
        // scans, then all storing.  Offsetting their starts spreads the HBM traffic over the column period.
        const unsigned ns = (unsigned)((c_first % 16) * (int64_t)P.stagger_ns / 16);
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)ns * 2) {
        }
    }
    // The ring holds exactly one column: chunk c (counted from the top of the degree chain) lives in
    // slot c and its mbarrier phase is the parity of the local column number.  Warp w issues chunk w.
    const double *xbase = P.X + (int64_t)c_first * P.ldx;
    const double *fbase = MUL ? P.F + (int64_t)c_first * P.ldf : nullptr;
    const int64_t fstep = (int64_t)ncl * P.ldf;
    if (ncols > 0) {
        if (tid == 0) {
            mbar_expect_tx(hfull, hbytes);
            tma_load_1d(Hb + 2, xbase + hA2, hbytes, hfull);
        }
        for (int c = warp; c < NCH; c += nw) f2_issue_chunk(P, full, ring, lane, k0, ml, xbase, c, c);
        if (MUL && P.gamma != 0.0 && tid == 0)
            for (int f = 0; f < P.nf && f < (int64_t)ncols * NCH; ++f) f2_issue_fchunk(P, ffull, fring, fbase + (f / NCH) * fstep, f % NCH, f);
    }
    // F ring cursor: next chunk of the global F sequence (column-major over columns, chunks ascending)
    int fslot = 0;
    uint32_t fph = 0;

    // thread <-> element: shifted so that a warp's 32 elements start on a 32 B boundary of the column
    const int par = NPAR == 1 ? 0 : tid / ncE;            // warp-uniform
    const int tp = NPAR == 1 ? tid : tid - par * ncE;
    const bool active = tp < ml;
    const int e = (tp + ((4 - ((nh + k0) & 3)) & 3)) % ml;
    const int el = active ? e : ml - 1;
    const int k = k0 + el;
    // rows were loaded as 16 B aligned supersets: element e of a row sits at offset e + par0
    const int par0 = P.contiguous ? (nh & 1) : ((nh + k0) & 1);
    const int off0 = el + par0;
    const int rp = P.rp;
    const double2 *tabp = tab + par;
    const double *tab1p = tab1 + par;
    // coupling entries of this element (rows b = NPAR*i + par), reloaded per use from L1/L2
    const double *sbk = P.Sb + k;
    const int expA = cs > 1 ? NPAR * 8 * ((rank > 0) + (rank < cs - 1)) : 0;

#ifdef F2_PROF
    const int pobs = (blockIdx.x == 0 && lane == 0) ? (warp == 0 ? 0 : (warp == nw - 1 ? 1 : -1)) : -1;
    long long pt = clock64();
#define F2_TICK(ph)                                                         \
    if (pobs >= 0) {                                                        \
        const long long now = clock64();                                    \
        atomicAdd(P.prof + pobs * 16 + (ph), (unsigned long long)(now - pt)); \
        pt = now;                                                           \
    }
#else
#define F2_TICK(ph)
#endif
    for (int it = 0; it < ncols; ++it) {
        double *xcol = P.X + ((int64_t)c_first + (int64_t)it * ncl) * P.ldx;
        const int hb = (int)(it & 1);
        const uint32_t hph = (uint32_t)(it >> 1) & 1;
        double *Hs = Hb + (size_t)hb * P.hcap + 2 + (hA & 1);   // Hs[u] = hat hA + u; Hs[-1], Hs[nhl] hold the neighbours' hats
        double *inbA = inb + F2_INB_A + hb * 4;

        // ---- column -> registers, chunk by chunk in backward-sweep order, each chunk swept as soon as it is in:
        //      z_j = (b_j - W1(j) z_{j+1} - W2(j) z_{j+2}) rd_j  (:439-441); the dependent chain runs under the wait
        //      for the next chunk --------------------------------------------------------------------------
        double z[ROWS];
#pragma unroll
        for (int cc = NCH - 1; cc >= 0; --cc) {
            mbar_wait(full + (NCH - 1 - cc), (uint32_t)it & 1);
            const double *pr = ring + (size_t)(NCH - 1 - cc) * P.slot_stride + off0;
#pragma unroll
            for (int r = 0; r < G; ++r) {
                const int j = cc * G + r;
                if (NPAR == 1) {
                    z[j] = (j <= JSAFE || j < q) ? *pr : 0.0;
                } else {
                    if (par == (j & 1)) z[j >> 1] = (j <= JSAFE || j < q) ? *pr : 0.0;
                }
                pr += rp;
            }
            constexpr int BL = G / NPAR;   // chain entries of this thread per chunk
#pragma unroll
            for (int i = (cc + 1) * BL - 1; i >= cc * BL; --i) {
                const double2 cf = tabp[NPAR * i];
                double v = z[i];
                if (NPAR == 1) {
                    if (ND >= 1 && !SKIP1 && i + 1 < ROWS) v = fma(-tab1p[i], z[i + 1], v);
                    if (ND >= 2 && i + 2 < ROWS) v = fma(-cf.y, z[i + 2], v);
                } else {
                    if (ND >= 2 && i + 1 < ROWS) v = fma(-cf.y, z[i + 1], v);
                }
                z[i] = v * cf.x;
            }
        }
        F2_TICK(0)   // column -> registers + backward sweep
        // ---- coupling to the hats: per-element contributions (:449-451) -----------------------
        {
            double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
            for (int i = 0; i < NBI; ++i) {
                const int b = NPAR * i + par;
                if (i < ROWS && b < nb) {
                    const double *sb = sbk + (size_t)b * 3 * m;
                    c0 = fma(__ldg(sb), z[i], c0);
                    c1 = fma(__ldg(sb + m), z[i], c1);
                    c2 = fma(__ldg(sb + 2 * m), z[i], c2);
                }
            }
            if (active) {
                ct[(par * 3 + 0) * ncE + e] = c0;
                ct[(par * 3 + 1) * ncE + e] = c1;
                ct[(par * 3 + 2) * ncE + e] = c2;
                if (cs > 1) {
                    // my edge elements also couple to a neighbour's hat
                    if (e == 0 && rank > 0) st_async_f64(mapa_u32(inbA + 2 + par, rank - 1), c0, mapa_u32(hxA + hb, rank - 1));
                    if (e == ml - 1 && rank < cs - 1) st_async_f64(mapa_u32(inbA + par, rank + 1), c2, mapa_u32(hxA + hb, rank + 1));
                }
            }
        }
        F2_TICK(1)   // backward sweep + couplings
        fence_proxy_async();   // this thread's earlier writes to the hat buffer vs. the TMA refill below
        __syncthreads();
        F2_TICK(2)   // barrier A
        if (expA && tid == 0) mbar_expect_tx(hxA + hb, (uint32_t)expA);
        if (expA) mbar_wait(hxA + hb, hph);
        mbar_wait(hfull + hb, hph);
        F2_TICK(4)   // wait for the neighbours' edge couplings and the hat load
        // ---- right-hand sides of my hats: hat hA+u collects elements u-1 (t=2), u (t=1), u+1 (t=0) ----
        // cs > 1: the two zero-incoming segment maps are dot products of these right-hand sides with constant weights
        // (see the set-up); every thread adds its hats, the warp sums meet the scan warps behind barrier B
        double sd = 0.0, su = 0.0;
        for (int u = tid; u < nhl; u += nthr) {
            double v = 0.0;
#pragma unroll
            for (int pp = 0; pp < NPAR; ++pp) {
                if (u + 1 < ml) v += ct[(pp * 3 + 0) * ncE + u + 1];
                else if (u + 1 == ml && rank < cs - 1) v += inbA[2 + pp];
                if (u < ml) v += ct[(pp * 3 + 1) * ncE + u];
                if (u >= 1) {
                    if (u - 1 < ml) v += ct[(pp * 3 + 2) * ncE + u - 1];
                } else if (rank > 0) {
                    v += inbA[pp];
                }
            }
            const double h = Hs[u] - v;
            Hs[u] = h;
            if (cs > 1) {
                sd = fma(hwd[u], h, sd);
                su = fma(hwu[u], h, su);
            }
        }
        if (cs > 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sd += __shfl_xor_sync(POP_FULL_MASK, sd, o);
                su += __shfl_xor_sync(POP_FULL_MASK, su, o);
            }
            if (lane == 0) {
                wagg[128 + warp] = sd;
                wagg[160 + warp] = su;
            }
        }
        F2_TICK(5)   // hat right-hand sides
        __syncthreads();
        F2_TICK(6)   // barrier B
#if F2_EARLY_COEF
        // coupling entries for the correction after the hat solve: issued now, their L2 latency passes under the scans
        double sc[NBI][3];
#pragma unroll
        for (int i = 0; i < NBI; ++i) {
            const int b = NPAR * i + par;
            const double *sb = sbk + (size_t)(b < nb ? b : 0) * 3 * m;
            sc[i][0] = __ldg(sb);
            sc[i][1] = __ldg(sb + m);
            sc[i][2] = __ldg(sb + 2 * m);
        }
#endif
        // every row of this column is in registers: refill the ring (and the other hat buffer).  The LAST warps
        // issue (warp nw-1 first): issuing a chunk of a cluster column (8 bulk copies from divergent lanes) stalls
        // the warp for thousands of cycles, so the scan warps stay out of it when other warps exist.
        {
            const int iw = nw - 1 - warp;
            const int nissue = nw > nsw ? nw - nsw : nw;
            if (iw == 0 && lane == 0 && it + 1 < ncols) {
                mbar_expect_tx(hfull + (hb ^ 1), hbytes);
                tma_load_1d(Hb + (size_t)(hb ^ 1) * P.hcap + 2, xcol + colstep + hA2, hbytes, hfull + (hb ^ 1));
            }
            if (MUL && P.gamma != 0.0 && tid == 32 % nthr) {
                // the part of this column's F that is not in the ring yet: pull it into L2 now
                const double *fc = fbase + (int64_t)it * fstep;
                for (int c = P.nf; c < NCH; ++c) {
                    const int64_t gs = ((int64_t)nh + (int64_t)c * G * m) & ~(int64_t)1;
                    const int nr = min(G, q - c * G);
                    if (nr > 0) l2_prefetch_bulk(fc + gs, (uint32_t)(((int64_t)nr * m + 2) & ~(int64_t)1) * 8u);
                }
            }
            if (it + 1 < ncols && iw < nissue)
                for (int c = iw; c < NCH; c += nissue) f2_issue_chunk(P, full, ring, lane, k0, ml, xcol + colstep, c, c);
        }
        F2_TICK(3)   // refill issue
        if (KSCAN) {
            if (warp == 0) {
                const int L = P.hat_L;
                double *hp = Hs + lane * L;
                const int nown = min(max(nhl - lane * L, 0), L);   // my hats that exist
                const double *cdT = kst + 10 * 32 + lane, *cuT = cdT + L * 32, *rdT = cuT + L * 32;
                double t = 0.0;
                for (int i = nown - 1; i >= 0; --i) t = fma(-cdT[i * 32], t, hp[i] * rdT[i * 32]);
#pragma unroll
                for (int st = 0; st < 5; ++st) t = fma(kst[st * 32 + lane], __shfl_down_sync(POP_FULL_MASK, t, 1 << st), t);
                double hin = __shfl_down_sync(POP_FULL_MASK, t, 1);
                if (lane == 31) hin = 0.0;
                for (int i = nown - 1; i >= 0; --i) {
                    hin = fma(-cdT[i * 32], hin, hp[i] * rdT[i * 32]);
                    hp[i] = hin;
                }
                t = 0.0;
                for (int i = 0; i < nown; ++i) t = fma(-cuT[i * 32], t, hp[i] * rdT[i * 32]);
#pragma unroll
                for (int st = 0; st < 5; ++st) t = fma(kst[(5 + st) * 32 + lane], __shfl_up_sync(POP_FULL_MASK, t, 1 << st), t);
                double xin = __shfl_up_sync(POP_FULL_MASK, t, 1);
                if (lane == 0) xin = 0.0;
                for (int i = 0; i < nown; ++i) {
                    xin = fma(-cuT[i * 32], xin, hp[i] * rdT[i * 32]);
                    hp[i] = xin;
                }
                if (lane == 0) {   // no neighbours: the hats next to the column's range are zero
                    Hs[-1] = 0.0;
                    Hs[nhl] = 0.0;
                    Hs[nhl + 1] = 0.0;
                }
            }
        } else if (warp < nsw) {
            double hl, t, Pm, at, aP, Gdn = 0.0, Gup = 0.0, Td = 0.0;
            double *ibc = inb + F2_INB_B + hb * 16;   // [rank][down, up]
            if (cs > 1) {
                // my segment's two zero-incoming maps -> every peer
                double Tu = 0.0;
                for (int w = 0; w < nw; ++w) {
                    Td += wagg[128 + w];
                    Tu += wagg[160 + w];
                }
                // lane r of warp 0 sends the pair to rank r: one instruction for all peers
                if (warp == 0 && lane < cs && lane != rank) st_async_v2f64(mapa_u32(ibc + 2 * rank, lane), Td, Tu, mapa_u32(hxB + hb, lane));
                if (tid == 0) mbar_expect_tx(hxB + hb, (uint32_t)(cs - 1) * 16u);
            }
            F2_TICK(11)   // dot products + send
            f2_scan_local<true>(Hs, hrd, hoffl, lo, hi, lane, warp, nsw, wagg, t, Pm, at, aP);
            F2_TICK(12)   // down: pass 1 + composition
            if (cs > 1) {
                mbar_wait(hxB + hb, hph);
                F2_TICK(13)   // wait for the peers' maps
                // value entering every rank's downward sweep (top down), and with it the exact upward maps below me
                const double *Pdn = inb + F2_INB_P, *Pup = Pdn + 8, *kap = Pdn + 16;
                double g = 0.0, mult = 1.0;
                for (int r = cs - 1; r >= 0; --r) {
                    if (r == rank) Gdn = g;
                    if (r < rank) {
                        Gup = fma(mult, fma(kap[r], g, ibc[2 * r + 1]), Gup);
                        mult *= Pup[r];
                    }
                    g = fma(Pdn[r], g, r == rank ? Td : ibc[2 * r]);
                }
            }
            f2_scan_finish<true>(Hs, hrd, hoffl, lo, hi, lane, warp, nsw, t, Pm, at, aP, Gdn, hl);
            F2_TICK(14)   // fold + down pass 2
            f2_scan_local<false>(Hs, hrd, hoffl, lo, hi, lane, warp, nsw, wagg + 64, t, Pm, at, aP);
            f2_scan_finish<false>(Hs, hrd, hoffl, lo, hi, lane, warp, nsw, t, Pm, at, aP, Gup, hl);
            // the neighbours' hats next to my segment: hat hA-1 entered the upward sweep; hat hB follows
            // from the value that entered the downward sweep and my last hat (both recurrences, one step)
            if (tid == 0) Hs[-1] = rank > 0 ? Gup : 0.0;
            if (nhl == 0 ? tid == 0 : (hi == nhl && hi > lo)) {
                Hs[nhl] = (hB < nh && nhl > 0) ? (Gdn - hoffl[nhl] * hl) * hrd[nhl] : 0.0;
                Hs[nhl + 1] = 0.0;
            }
        }
        F2_TICK(7)   // hat scans (scan warps only)
        __syncthreads();
        F2_TICK(8)   // barrier C
        // ---- correct the coupled bubbles (:475-477): hats k-1, k, k+1 of element k ----------------
        const double h0 = Hs[el - 1], h1 = Hs[el], h2 = Hs[el + 1];
#pragma unroll
        for (int i = 0; i < NBI; ++i) {
            const int b = NPAR * i + par;
            if (i < ROWS && b < nb) {
                double v = z[i];
#if F2_EARLY_COEF
                v = fma(-sc[i][0], h0, v);
                v = fma(-sc[i][1], h1, v);
                v = fma(-sc[i][2], h2, v);
#else
                const double *sb = sbk + (size_t)b * 3 * m;
                v = fma(-__ldg(sb), h0, v);
                v = fma(-__ldg(sb + m), h1, v);
                v = fma(-__ldg(sb + 2 * m), h2, v);
#endif
                z[i] = v;
            }
        }
        // ---- forward sweep  x_j = (z_j - W1(j-1) x_{j-1} - W2(j-2) x_{j-2}) rd_j  (:480-482) --------
        // plain solve: every x_j leaves for global memory as soon as it is final, so the stores drain under the
        // latency of the dependent chain instead of in one burst behind it
        double *xo = xcol + nh + k + (NPAR == 2 ? (int64_t)par * m : 0);
        const int64_t xs = (int64_t)NPAR * m;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const double2 cf = tabp[NPAR * i];
            double v = z[i];
            if (NPAR == 1) {
                if (ND >= 1 && !SKIP1 && i >= 1) v = fma(-tab1p[i - 1], z[i - 1], v);
                if (ND >= 2 && i >= 2) v = fma(-tabp[i - 2].y, z[i - 2], v);
            } else {
                if (ND >= 2 && i >= 1) v = fma(-tabp[NPAR * (i - 1)].y, z[i - 1], v);
            }
            z[i] = v * cf.x;
            if (!MUL) {
                const int jj = NPAR * i;                       // row j = jj + par
                if (active && (jj + NPAR - 1 <= JSAFE || jj + par < q)) F2_STORE(xo, z[i]);
                xo += xs;
            }
        }
        F2_TICK(9)   // correction + forward sweep (+ the bubble stores of the plain solve)
        if (!MUL) {
            for (int u = tid; u < nhl; u += nthr) xcol[hA + u] = Hs[u];
        } else {
            // ---- fused product  r = T x + gamma f,  T = Symmetric(upper data); cs == 1: all hats are mine ----
            // NPAR == 2: rows j = 2i + par; T never couples j and j+1 there, so a thread needs only its own parity.
            const bool useF = P.gamma != 0.0;
            const double *fcol = P.F + ((int64_t)c_first + (int64_t)it * ncl) * P.ldf;
            const double *t0 = tT + par, *t1 = tT + (JMAX + 2) + par, *t2 = tT + 2 * (JMAX + 2) + par;
            // the hat rows of the product read x of the coupled bubbles
#pragma unroll
            for (int i = 0; i < NBI; ++i) {
                const int b = NPAR * i + par;
                if (i < ROWS && b < P.nbT && active) xb[b * ncE + e] = z[i];
            }
            constexpr int BLK = G / NPAR;                     // thread rows per F chunk
            const int64_t xs = (int64_t)NPAR * m;
            double *xo = xcol + nh + k + (NPAR == 2 ? (int64_t)par * m : 0);
            const int foff = el + par0 + (NPAR == 2 ? par * m : 0);
            const int fstride = NPAR * m;
            const int64_t nfch = (int64_t)ncols * NCH;                 // F chunks of this cluster's sequence
#pragma unroll
            for (int cch = 0; cch < NCH; ++cch) {
                const int i0 = cch * BLK;
                double fq[BLK];
                if (useF) {
                    // refill the slot of the previous chunk once every warp has read it (deferred by one chunk)
                    if (cch > 0 && warp == ((cch - 1) % nw)) {
                        const int ps = fslot == 0 ? P.nf - 1 : fslot - 1;
                        const uint32_t pph = fslot == 0 ? fph ^ 1 : fph;
                        mbar_wait(fempty + ps, pph);
                        const int64_t fnext = (int64_t)it * NCH + (cch - 1) + P.nf;
                        if (lane == 0 && fnext < nfch) f2_issue_fchunk(P, ffull, fring, fbase + (fnext / NCH) * fstep, (int)(fnext % NCH), ps);
                    }
                    mbar_wait(ffull + fslot, fph);
                    const double *fp = fring + (size_t)fslot * P.slot_stride + foff;
#pragma unroll
                    for (int u = 0; u < BLK; ++u) {
                        fq[u] = *fp;
                        fp += fstride;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(fempty + fslot);
                    if (++fslot == P.nf) {
                        fslot = 0;
                        fph ^= 1;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < BLK; ++u) fq[u] = 0.0;
                }
#pragma unroll
                for (int u = 0; u < BLK; ++u) {
                    const int i = i0 + u;
                    const int jj = NPAR * i;          // row j = jj + par; table offsets are relative to par
                    double acc = t0[jj] * z[i];
                    if (NPAR == 1) {
                        if (!SKIP1) {
                            if (i + 1 < ROWS) acc = fma(t1[jj], z[i + 1], acc);
                            if (i >= 1) acc = fma(t1[jj - 1], z[i - 1], acc);
                        }
                        if (i + 2 < ROWS) acc = fma(t2[jj], z[i + 2], acc);
                        if (i >= 2) acc = fma(t2[jj - 2], z[i - 2], acc);
                    } else {
                        if (i + 1 < ROWS) acc = fma(t2[jj], z[i + 1], acc);
                        if (i >= 1) acc = fma(t2[jj - 2], z[i - 1], acc);
                    }
                    if (i < NBI && jj + par < P.nbT) {
                        const double *sb = P.Trow + (size_t)(jj + par) * 3 * m + k;
                        acc = fma(__ldg(sb), h0, acc);
                        acc = fma(__ldg(sb + m), h1, acc);
                        acc = fma(__ldg(sb + 2 * m), h2, acc);
                    }
                    if (active && (jj + NPAR - 1 <= JSAFE || jj + par < q)) F2_STORE(xo, fma(P.gamma, fq[u], acc));
                    xo += xs;
                }
            }
            __syncthreads();
            if (useF && tid == 0) {
                // every warp is past the product loop: the slot of the last chunk can be refilled without waiting
                const int ps = fslot == 0 ? P.nf - 1 : fslot - 1;
                const int64_t fnext = (int64_t)it * NCH + (NCH - 1) + P.nf;
                if (fnext < nfch) f2_issue_fchunk(P, ffull, fring, fbase + (fnext / NCH) * fstep, (int)(fnext % NCH), ps);
            }
            for (int i = tid; i < nh; i += nthr) {
                double acc = P.TAd[i] * Hs[i];
                if (i + 1 < nh) acc = fma(P.TAu[i], Hs[i + 1], acc);
                if (i > 0) acc = fma(P.TAu[i - 1], Hs[i - 1], acc);
                for (int b = 0; b < P.nbT; ++b)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int kk = i - t + 1;
                        if (kk >= 0 && kk < m) acc = fma(__ldg(P.Trow + ((size_t)b * 3 + t) * m + kk), xb[b * ncE + kk], acc);
                    }
                xcol[i] = useF ? fma(P.gamma, fcol[i], acc) : acc;
            }
        }
        F2_TICK(10)   // stores (or product + stores)
    }
    if (cs > 1) {
        cluster_arrive();
        cluster_wait();
    }
}

// one translation unit per row class: NPAR = 1 for every class, NPAR = 2 (q up to 128) for 40..64
template <int ROWS, int NPAR, int ND, bool SKIP1>
static f2_kernel_t f2_pick3(bool mul, bool cl) {
    if (mul) return cl ? nullptr : k_fused2<ROWS, NPAR, ND, SKIP1, true, false>;
    return cl ? k_fused2<ROWS, NPAR, ND, SKIP1, false, true> : k_fused2<ROWS, NPAR, ND, SKIP1, false, false>;
}
template <int ROWS>
static f2_kernel_t f2_pick(int npar, int variant, bool mul, bool cl) {
    if (npar == 2) return variant == 0 ? f2_pick3<ROWS, 2, 0, false>(mul, cl) : f2_pick3<ROWS, 2, 2, true>(mul, cl);
    if (variant == 0) return f2_pick3<ROWS, 1, 0, false>(mul, cl);
    if (variant == 1) return f2_pick3<ROWS, 1, 2, true>(mul, cl);
    return f2_pick3<ROWS, 1, 2, false>(mul, cl);
}
#define F2_DEFINE_CLASS(ROWS) \
    f2_kernel_t f2_kernel_r##ROWS(int npar, int variant, bool mul, bool cl) { return f2_pick<ROWS>(npar, variant, mul, cl); }
#endif  // __CUDACC__
