// Single-pass ReverseCholesky solve  X <- U^{-T} U^{-1} X  (optionally followed by the ADI
// product  R <- T X + gamma F) with every column resident on chip between the two sweeps:
// one HBM read and one HBM write per unknown (16 B; 24 B with the fused product).
//
// Replaces, for matrix right-hand sides, the column loop over
//   materialize!(::MatLdivVec{TriangularLayout{'U'}}) (src/arrowhead.jl:425-457) followed by
//   materialize!(::MatLdivVec{TriangularLayout{'L'}}) (src/arrowhead.jl:458-486)
// and, in the fused form, one half step of examples/adi_poisson.jl:54-55.
//
// Mapping (sm_100a):
//   * one thread-block cluster of CS CTAs owns one column at a time; CTA r holds the bubble
//     rows of elements [k0,k1) in shared memory (row pitch even, 16 B aligned rows);
//   * rows arrive by TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) of 16 B aligned
//     supersets, so Dirichlet columns (odd nh) need no peeling;
//   * lanes run over elements (consecutive doubles -> conflict-free LDS/STS), each thread walks
//     its element's degree chain; the backward sweep leaves z in shared memory, the forward
//     sweep streams x (or T x + gamma F) straight to global memory from registers;
//   * hats: per-CTA partial sums, exchanged through distributed shared memory when CS > 1,
//     then a segmented affine scan on one warp of rank 0.
#include <cooperative_groups.h>

#include <cstdlib>

#include "pop_internal.h"

namespace cg = cooperative_groups;

#define FULL 0xffffffffu
#define FUSED_MAX_ND 2

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct FusedParams {
    TriView up;       // U   (backward): W_d(j) couples j, j+d
    MulView T;        // optional product operator (symmetric view), valid when do_mul
    int do_mul;
    double gamma;
    const double *F;  // n x nrhs, leading dimension ldf
    int64_t ldf;
    double *X;
    int64_t ldx;
    int64_t nrhs;
    int cs;           // cluster size
    int mloc;         // elements per CTA (last rank may hold fewer)
    int pitch;        // doubles per shared-memory row (cs == 1: m, rows contiguous as in global memory)
    int uniform_tabs; // factor tables of the shared D block are staged in shared memory
};

// Shared-memory carve-up (doubles).  cs == 1: `img` is an exact image of the global column
// (hats, then rows of m), loaded by one run of bulk copies; H = img, slab = img + nh.
// cs > 1: H (rank 0) and the slab rows (pitch even, 16 B aligned) are separate.
struct Smem {
    double *H;      // hats of the column
    double *slab;   // bubble rows of this CTA's elements
    double *hloc;   // [mloc + 2]    hats k0-1 .. k1 seen by this CTA          (cs > 1)
    double *part;   // [mloc + 2]    partial sums for hats k0-1 .. k1          (cs > 1)
    double2 *tB;    // [q] (rd[j], W2[j])     backward tables of the shared D block
    double2 *tF;    // [q] (rd[j], W2[j-2])   forward tables
    double *t1B;    // [q] W1[j]
    double *t1F;    // [q] W1[j-1]
    double *tT;     // [3q] T.D(j,j), T.D(j,j+1), T.D(j,j+2)                    (fused product)
    double *hrd;    // [nh] reciprocal hat pivots
    double *hoff;   // [nh] hat super-diagonal of U (0 in the last slot)
    double *wagg;   // [128] warp aggregates of the block-wide hat scan
    uint64_t *bar;
};

__host__ __device__ inline size_t ev2(size_t x) { return (x + 1) & ~(size_t)1; }

__host__ __device__ inline size_t fused_layout(Smem *s, double *base, int cs, int q, int pitch, int nh, int m, int mloc, bool tabs, bool mul) {
    double *p = base;
    double *H, *slab;
    if (cs == 1) {
        H = p; slab = p + nh; p += ev2((size_t)nh + (size_t)q * m) + 2;
    } else {
        H = p; p += ev2(nh + 2);
        slab = p; p += (size_t)q * pitch;
    }
    double *hloc = p; p += ev2(mloc + 2);
    double *part = p; p += ev2(mloc + 2);
    double *tB = p; p += tabs ? 2 * (size_t)q : 0;
    double *tF = p; p += tabs ? 2 * (size_t)q : 0;
    double *t1B = p; p += tabs ? ev2(q) : 0;
    double *t1F = p; p += tabs ? ev2(q) : 0;
    double *tT = p; p += mul ? ev2(3 * (size_t)q) : 0;
    double *hrd = p; p += ev2(nh);
    double *hoff = p; p += ev2(nh);
    double *wagg = p; p += 128;
    double *bar = p; p += 2;
    if (s) {
        s->H = H; s->slab = slab; s->hloc = hloc; s->part = part;
        s->tB = reinterpret_cast<double2 *>(tB); s->tF = reinterpret_cast<double2 *>(tF);
        s->t1B = t1B; s->t1F = t1F; s->tT = tT; s->hrd = hrd; s->hoff = hoff; s->wagg = wagg;
        s->bar = reinterpret_cast<uint64_t *>(bar);
    }
    return (size_t)(p - base) * sizeof(double);
}

// coefficient access for the sweeps: shared-memory tables (one D block for all elements) ...
struct CoefTabs {
    const double2 *tB, *tF;
    const double *t1B, *t1F;
    __device__ __forceinline__ double2 bwd(int j) const { return tB[j]; }
    __device__ __forceinline__ double bwd1(int j) const { return t1B[j]; }
    __device__ __forceinline__ double2 fwd(int j) const { return tF[j]; }
    __device__ __forceinline__ double fwd1(int j) const { return t1F[j]; }
};
// ... or per-element blocks read from global memory (coalesced over elements, L1/L2 resident)
struct CoefElem {
    const double *rd, *w1, *w2;   // already offset by the element index; row stride m
    int m;
    __device__ __forceinline__ double2 bwd(int j) const { return make_double2(rd[(size_t)j * m], w2 ? w2[(size_t)j * m] : 0.0); }
    __device__ __forceinline__ double bwd1(int j) const { return w1 ? w1[(size_t)j * m] : 0.0; }
    __device__ __forceinline__ double2 fwd(int j) const { return make_double2(rd[(size_t)j * m], (w2 && j >= 2) ? w2[(size_t)(j - 2) * m] : 0.0); }
    __device__ __forceinline__ double fwd1(int j) const { return (w1 && j >= 1) ? w1[(size_t)(j - 1) * m] : 0.0; }
};

// z_j = (b_j - W1(j) z_{j+1} - W2(j) z_{j+2}) rd_j, j = q-1..0, in place in shared memory
// (src/arrowhead.jl:439-441).  Rows are processed in blocks of 4: all loads of a block are issued
// before its dependent chain, all stores after it.
template <int ND, bool SKIP1, class Coef>
__device__ __forceinline__ void sweep_bwd(double *p /* row q-1 */, const int pitch, const int q, const Coef cf) {
    double z1 = 0.0, z2 = 0.0;
    int j = q - 1;
    for (; j >= 3; j -= 4) {
        double b[4], w1[4];
        double2 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            b[u] = p[-u * pitch];
            c[u] = cf.bwd(j - u);
            w1[u] = (ND >= 1 && !SKIP1) ? cf.bwd1(j - u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double v = b[u];
            if (ND >= 1 && !SKIP1) v = fma(-w1[u], z1, v);
            if (ND >= 2) v = fma(-c[u].y, z2, v);
            v *= c[u].x;
            b[u] = v;
            z2 = z1;
            z1 = v;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) p[-u * pitch] = b[u];
        p -= 4 * pitch;
    }
    for (; j >= 0; --j) {
        double v = *p;
        const double2 c = cf.bwd(j);
        if (ND >= 1 && !SKIP1) v = fma(-cf.bwd1(j), z1, v);
        if (ND >= 2) v = fma(-c.y, z2, v);
        v *= c.x;
        *p = v;
        z2 = z1;
        z1 = v;
        p -= pitch;
    }
}

// x_j = (z_j - W1(j-1) x_{j-1} - W2(j-2) x_{j-2}) rd_j, j = 0..q-1, streamed to global memory
// (src/arrowhead.jl:480-482)
template <int ND, bool SKIP1, class Coef>
__device__ __forceinline__ void sweep_fwd(const double *p /* row 0 */, const int pitch, const int q, const Coef cf, double *xo, const int64_t xs) {
    double x1 = 0.0, x2 = 0.0;
    int j = 0;
    for (; j + 3 < q; j += 4) {
        double b[4], w1[4];
        double2 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            b[u] = p[u * pitch];
            c[u] = cf.fwd(j + u);
            w1[u] = (ND >= 1 && !SKIP1) ? cf.fwd1(j + u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double v = b[u];
            if (ND >= 1 && !SKIP1) v = fma(-w1[u], x1, v);
            if (ND >= 2) v = fma(-c[u].y, x2, v);
            v *= c[u].x;
            b[u] = v;
            x2 = x1;
            x1 = v;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) xo[(int64_t)u * xs] = b[u];
        p += 4 * pitch;
        xo += 4 * xs;
    }
    for (; j < q; ++j) {
        double v = *p;
        const double2 c = cf.fwd(j);
        if (ND >= 1 && !SKIP1) v = fma(-cf.fwd1(j), x1, v);
        if (ND >= 2) v = fma(-c.y, x2, v);
        v *= c.x;
        *xo = v;
        x2 = x1;
        x1 = v;
        p += pitch;
        xo += xs;
    }
}

template <int ND, bool SKIP1, bool MUL>
__global__ void __launch_bounds__(512, 1) k_fused_solve(const FusedParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = P.cs;
    const int rank = cs > 1 ? (int)cluster.block_rank() : 0;
    const TriView &tv = P.up;
    const int m = tv.m, nh = tv.nh, q = tv.q, mD = tv.mD, nb = tv.nb;
    const int k0 = rank * P.mloc;
    const int k1 = min(m, k0 + P.mloc);
    const int ml = k1 - k0;          // elements of this CTA (>= 1 by construction)
    const int pitch = P.pitch;
    const bool tabs = P.uniform_tabs != 0;
    Smem s;
    fused_layout(&s, reinterpret_cast<double *>(smem_raw), cs, q, pitch, nh, m, P.mloc, tabs, MUL);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    // cs > 1: parity of the global index of this CTA's first element in row j is odd0 ^ (j & modd);
    // the row was loaded as an aligned superset, element e sits at row offset e + parity.
    const int odd0 = cs > 1 ? ((nh + k0) & 1) : 0, modd = cs > 1 ? (m & 1) : 0;

    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
    }
    if (tabs) {
        for (int j = tid; j < q; j += nthr) {
            const double rd = tv.rdD[j];
            const double w1 = (ND >= 1 && tv.nd >= 1) ? tv.W[0][j] : 0.0;
            const double w2 = (ND >= 2 && tv.nd >= 2) ? tv.W[1][j] : 0.0;
            s.tB[j] = make_double2(rd, w2);
            s.t1B[j] = w1;
            s.tF[j] = make_double2(rd, (ND >= 2 && tv.nd >= 2 && j >= 2) ? tv.W[1][j - 2] : 0.0);
            s.t1F[j] = (ND >= 1 && tv.nd >= 1 && j >= 1) ? tv.W[0][j - 1] : 0.0;
        }
    }
    if (MUL) {
        const MulView &T = P.T;
        for (int j = tid; j < q; j += nthr) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                s.tT[d * q + j] = (d <= T.uD && j + d < q) ? T.Dp[POP_MAXD + d][(size_t)j * T.mD] : 0.0;
        }
    }
    for (int i = tid; i < nh; i += nthr) {
        s.hrd[i] = tv.rdA[i];
        s.hoff[i] = i < nh - 1 ? tv.offA[i] : 0.0;
    }
    if (cs > 1) cluster.sync(); else __syncthreads();

    double *Hr0 = cs > 1 ? cluster.map_shared_rank(s.H, 0) : s.H;   // hats of the column live on rank 0
    const CoefTabs ctab = {s.tB, s.tF, s.t1B, s.t1F};
    uint32_t phase = 0;
    const int64_t ncl = gridDim.x / cs;
    for (int64_t c = blockIdx.x / cs; c < P.nrhs; c += ncl) {
        double *xcol = P.X + c * P.ldx;
        // ---- load -------------------------------------------------------------------------
        if (cs == 1) {
            if (tid == 0) {
                fence_proxy_async();
                const uint32_t total = (uint32_t)ev2((size_t)nh + (size_t)q * m) * 8u;
                mbar_expect_tx(s.bar, total);
                for (uint32_t off = 0; off < total; off += 32768u) {
                    const uint32_t len = min(32768u, total - off);
                    tma_load_1d(reinterpret_cast<unsigned char *>(s.H) + off, reinterpret_cast<const unsigned char *>(xcol) + off, len, s.bar);
                }
            }
        } else if (warp == 0) {
            // aligned supersets of the row pieces, issued by the 32 lanes in parallel
            if (lane == 0) {
                fence_proxy_async();
                uint32_t total = 0;
                for (int j = 0; j < q; ++j) {
                    const int64_t g0 = (int64_t)nh + (int64_t)j * m + k0;
                    total += (uint32_t)((((g0 + ml + 1) & ~(int64_t)1) - (g0 & ~(int64_t)1)) * 8);
                }
                if (rank == 0) total += (uint32_t)ev2(nh) * 8u;
                mbar_expect_tx(s.bar, total);
                if (rank == 0) tma_load_1d(s.H, xcol, (uint32_t)ev2(nh) * 8u, s.bar);
            }
            __syncwarp();
            for (int j = lane; j < q; j += 32) {
                const int64_t g0 = (int64_t)nh + (int64_t)j * m + k0;
                const int64_t a0 = g0 & ~(int64_t)1, a1 = (g0 + ml + 1) & ~(int64_t)1;
                tma_load_1d(s.slab + (size_t)j * pitch, xcol + a0, (uint32_t)(a1 - a0) * 8u, s.bar);
            }
        }
        while (!mbar_try_wait(s.bar, phase)) {
        }
        phase ^= 1;

        // ---- backward sweep ---------------------------------------------------------------
        for (int e = tid; e < ml; e += nthr) {
            if (modd == 0) {
                double *p = s.slab + (size_t)(q - 1) * pitch + e + odd0;
                if (tabs) {
                    sweep_bwd<ND, SKIP1>(p, pitch, q, ctab);
                } else {
                    const int k = k0 + e;
                    const CoefElem ce = {tv.rdD + k, tv.nd >= 1 ? tv.W[0] + k : nullptr, tv.nd >= 2 ? tv.W[1] + k : nullptr, m};
                    sweep_bwd<ND, SKIP1>(p, pitch, q, ce);
                }
            } else {
                // odd m inside a cluster: the row parity alternates, plain indexed loop
                const int k = k0 + e, kD = mD == 1 ? 0 : k;
                double z1 = 0.0, z2 = 0.0;
                for (int j = q - 1; j >= 0; --j) {
                    double *pz = s.slab + (size_t)j * pitch + e + (odd0 ^ (j & 1));
                    double v = *pz;
                    if (ND >= 1 && tv.nd >= 1) v = fma(-tv.W[0][(size_t)j * mD + kD], z1, v);
                    if (ND >= 2 && tv.nd >= 2) v = fma(-tv.W[1][(size_t)j * mD + kD], z2, v);
                    v *= tv.rdD[(size_t)j * mD + kD];
                    *pz = v;
                    z2 = z1;
                    z1 = v;
                }
            }
        }
        // ---- hats ---------------------------------------------------------------------------
        if (cs == 1) {
            __syncthreads();
            // H[i] -= sum_b sum_t S[b][t][k] Z[b][k], k = i-t+1   (src/arrowhead.jl:449-451)
            for (int i = tid; i < nh; i += nthr) {
                double v = s.H[i];
                for (int b = 0; b < nb; ++b)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int k = i - t + 1;
                        if (k >= 0 && k < m) v -= tv.S[((size_t)b * 3 + t) * m + k] * s.slab[(size_t)b * pitch + k];
                    }
                s.H[i] = v;
            }
        } else {
            for (int e = tid; e < ml + 2; e += nthr) s.part[e] = 0.0;
            __syncthreads();
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                for (int e = tid; e < ml; e += nthr) {
                    const int k = k0 + e;
                    double acc = 0.0;
                    for (int b = 0; b < nb; ++b) acc += tv.S[((size_t)b * 3 + t) * m + k] * s.slab[(size_t)b * pitch + e + (odd0 ^ (b & modd))];
                    s.part[e + t] += acc;
                }
                __syncthreads();
            }
            cluster.sync();
            if (rank == 0) {
                for (int i = tid; i < nh; i += nthr) {
                    double v = s.H[i];
                    for (int r = 0; r < cs; ++r) {
                        const int rk0 = r * P.mloc;
                        if (rk0 >= m) break;
                        const int rml = min(m, rk0 + P.mloc) - rk0;
                        const int e = i - (rk0 - 1);
                        if (e >= 0 && e < rml + 2) {
                            const double *rp = r != 0 ? cluster.map_shared_rank(s.part, r) : s.part;
                            v -= rp[e];
                        }
                    }
                    s.H[i] = v;
                }
            }
        }
        if (rank == 0) {
            __syncthreads();
            // Block-wide segmented affine scan: thread t owns hats [t*sg, (t+1)*sg).
            //   down: h[i] = (H[i] - off[i] h[i+1]) rd[i]     (U solve,   src/arrowhead.jl:453)
            //   up  : h[i] = (H[i] - off[i-1] h[i-1]) rd[i]   (U' solve,  src/arrowhead.jl:472)
            // pass 1 reduces a segment to an affine map of its incoming value, warp shuffles and a
            // short loop over the warp aggregates compose the maps, pass 2 is the reference
            // recurrence itself with the exact incoming value.  Factor entries come from smem.
            const int sg = (nh + nthr - 1) / nthr;
            const int lo = min(tid * sg, nh), hi = min(lo + sg, nh);
            const int nw = nthr >> 5;
#pragma unroll
            for (int dir = 0; dir < 2; ++dir) {
                const bool bw = dir == 0;
                double t = 0.0, Pm = 1.0;
                for (int ii = lo; ii < hi; ++ii) {
                    const int i = bw ? (hi - 1 - (ii - lo)) : ii;
                    const double off = bw ? s.hoff[i] : (i > 0 ? s.hoff[i - 1] : 0.0);
                    const double rd = s.hrd[i];
                    t = (s.H[i] - off * t) * rd;
                    Pm = -off * rd * Pm;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t2 = bw ? __shfl_down_sync(FULL, t, o) : __shfl_up_sync(FULL, t, o);
                    const double P2 = bw ? __shfl_down_sync(FULL, Pm, o) : __shfl_up_sync(FULL, Pm, o);
                    const bool ok = bw ? (lane + o < 32) : (lane - o >= 0);
                    if (ok) {
                        t = t + Pm * t2;
                        Pm = Pm * P2;
                    }
                }
                // lane now holds the composite of lanes lane..31 (down) / 0..lane (up)
                if (lane == (bw ? 0 : 31)) {
                    s.wagg[dir * 64 + warp] = t;
                    s.wagg[dir * 64 + 32 + warp] = Pm;
                }
                __syncthreads();
                double G = 0.0;   // value entering this warp from the warps after (down) / before (up) it
                if (bw) {
                    for (int w2 = nw - 1; w2 > warp; --w2) G = s.wagg[w2] + s.wagg[32 + w2] * G;
                } else {
                    for (int w2 = 0; w2 < warp; ++w2) G = s.wagg[64 + w2] + s.wagg[96 + w2] * G;
                }
                const double tn = bw ? __shfl_down_sync(FULL, t, 1) : __shfl_up_sync(FULL, t, 1);
                const double Pn = bw ? __shfl_down_sync(FULL, Pm, 1) : __shfl_up_sync(FULL, Pm, 1);
                double hin = (bw ? (lane == 31) : (lane == 0)) ? G : tn + Pn * G;
                for (int ii = lo; ii < hi; ++ii) {
                    const int i = bw ? (hi - 1 - (ii - lo)) : ii;
                    const double off = bw ? s.hoff[i] : (i > 0 ? s.hoff[i - 1] : 0.0);
                    hin = (s.H[i] - off * hin) * s.hrd[i];
                    s.H[i] = hin;
                }
            }
        }
        if (cs > 1) cluster.sync(); else __syncthreads();
        // ---- correct the coupled bubbles (src/arrowhead.jl:475-477) --------------------------
        if (cs > 1) {
            for (int e = tid; e < ml + 2; e += nthr) {
                const int i = k0 - 1 + e;
                s.hloc[e] = (i >= 0 && i < nh) ? Hr0[i] : 0.0;
            }
            __syncthreads();
        }
        for (int e = tid; e < ml; e += nthr) {
            const int k = k0 + e;
            for (int b = 0; b < nb; ++b) {
                double *pz = s.slab + (size_t)b * pitch + e + (odd0 ^ (b & modd));
                double v = *pz;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int i = k + t - 1;
                    const double hv = cs > 1 ? s.hloc[e + t] : ((i >= 0 && i < nh) ? s.H[i] : 0.0);
                    v -= tv.S[((size_t)b * 3 + t) * m + k] * hv;
                }
                *pz = v;
            }
        }
        // each thread corrected only its own elements: the forward sweep below needs no barrier

        if (!MUL) {
            // hats out (rank 0), then forward sweep streaming x to global memory
            if (rank == 0)
                for (int i = tid; i < nh; i += nthr) xcol[i] = s.H[i];
            for (int e = tid; e < ml; e += nthr) {
                const int k = k0 + e;
                if (modd == 0) {
                    const double *p = s.slab + e + odd0;
                    if (tabs) {
                        sweep_fwd<ND, SKIP1>(p, pitch, q, ctab, xcol + nh + k, (int64_t)m);
                    } else {
                        const CoefElem ce = {tv.rdD + k, tv.nd >= 1 ? tv.W[0] + k : nullptr, tv.nd >= 2 ? tv.W[1] + k : nullptr, m};
                        sweep_fwd<ND, SKIP1>(p, pitch, q, ce, xcol + nh + k, (int64_t)m);
                    }
                } else {
                    const int kD = mD == 1 ? 0 : k;
                    double x1 = 0.0, x2 = 0.0;
                    for (int j = 0; j < q; ++j) {
                        double v = s.slab[(size_t)j * pitch + e + (odd0 ^ (j & 1))];
                        if (ND >= 1 && tv.nd >= 1 && j >= 1) v = fma(-tv.W[0][(size_t)(j - 1) * mD + kD], x1, v);
                        if (ND >= 2 && tv.nd >= 2 && j >= 2) v = fma(-tv.W[1][(size_t)(j - 2) * mD + kD], x2, v);
                        v *= tv.rdD[(size_t)j * mD + kD];
                        xcol[(int64_t)nh + (int64_t)j * m + k] = v;
                        x2 = x1;
                        x1 = v;
                    }
                }
            }
        } else {
            // ---- fused product (cs == 1):  r = T x + gamma f, T = Symmetric(upper data) ----------
            // The forward sweep keeps the window x[j-4..j]; row jo = j-2 of the product is emitted
            // once x[jo+2] is known:  r[jo] = sum_d T(jo,jo+d) x[jo+d] + coupling + gamma f[jo].
            const MulView &T = P.T;
            const bool useF = P.gamma != 0.0;
            const double *fcol = useF ? P.F + c * P.ldf : nullptr;
            const int tD = T.mD;
            for (int e = tid; e < ml; e += nthr) {
                const int k = k0 + e;
                const int kD = mD == 1 ? 0 : k, kT = tD == 1 ? 0 : k;
                const double *pz = s.slab + e;
                double *xo = xcol + nh + k;
                const double *fo = useF ? fcol + nh + k : nullptr;
                double xw[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // xw[i] = x[j - i]
                double fq[4], zq[4];
                for (int j0 = 0; j0 < q + 2; j0 += 4) {
                    // prefetch the 4 F entries this block of output rows (j0-2 .. j0+1) needs and the
                    // 4 rows of z it consumes (rows >= j0 are still untouched by this thread)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jo = j0 + u - 2;
                        fq[u] = (useF && jo >= 0 && jo < q) ? fo[(int64_t)jo * m] : 0.0;
                        zq[u] = (j0 + u < q) ? pz[(size_t)(j0 + u) * pitch] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + u;
                        double v = 0.0;
                        if (j < q) {
                            v = zq[u];
                            double2 cf;
                            double w1;
                            if (tabs) {
                                cf = s.tF[j];
                                w1 = s.t1F[j];
                            } else {
                                cf = make_double2(tv.rdD[(size_t)j * mD + kD], (tv.nd >= 2 && j >= 2) ? tv.W[1][(size_t)(j - 2) * mD + kD] : 0.0);
                                w1 = (tv.nd >= 1 && j >= 1) ? tv.W[0][(size_t)(j - 1) * mD + kD] : 0.0;
                            }
                            if (ND >= 1 && !SKIP1) v = fma(-w1, xw[0], v);
                            if (ND >= 2) v = fma(-cf.y, xw[1], v);
                            v *= cf.x;
                            if (j < T.nb_row) s.slab[(size_t)j * pitch + e] = v;   // the hat rows of the product read x_Z[b]
                        }
                        xw[4] = xw[3]; xw[3] = xw[2]; xw[2] = xw[1]; xw[1] = xw[0]; xw[0] = v;
                        const int jo = j - 2;
                        if (jo >= 0 && jo < q) {
                            double t0, t1u, t2u, t1l, t2l;   // T(jo,jo), T(jo,jo+1), T(jo,jo+2), T(jo-1,jo), T(jo-2,jo)
                            if (tD == 1) {
                                t0 = s.tT[jo]; t1u = s.tT[q + jo]; t2u = s.tT[2 * q + jo];
                                t1l = jo >= 1 ? s.tT[q + jo - 1] : 0.0;
                                t2l = jo >= 2 ? s.tT[2 * q + jo - 2] : 0.0;
                            } else {
                                t0 = T.Dp[POP_MAXD][(size_t)jo * tD + kT];
                                t1u = (T.uD >= 1 && jo + 1 < q) ? T.Dp[POP_MAXD + 1][(size_t)jo * tD + kT] : 0.0;
                                t2u = (T.uD >= 2 && jo + 2 < q) ? T.Dp[POP_MAXD + 2][(size_t)jo * tD + kT] : 0.0;
                                t1l = (T.uD >= 1 && jo >= 1) ? T.Dp[POP_MAXD + 1][(size_t)(jo - 1) * tD + kT] : 0.0;
                                t2l = (T.uD >= 2 && jo >= 2) ? T.Dp[POP_MAXD + 2][(size_t)(jo - 2) * tD + kT] : 0.0;
                            }
                            double acc = t0 * xw[2];
                            acc = fma(t1u, xw[1], acc);
                            acc = fma(t1l, xw[3], acc);
                            acc = fma(t2u, xw[0], acc);
                            acc = fma(t2l, xw[4], acc);
                            if (jo < T.nb_col) {
#pragma unroll
                                for (int t = 0; t < 3; ++t) {
                                    const int i = k + t - 1;
                                    if (i >= 0 && i < nh) acc = fma(T.Scol[((size_t)jo * 3 + t) * m + k], s.H[i], acc);
                                }
                            }
                            xo[(int64_t)jo * m] = useF ? fma(P.gamma, fq[u], acc) : acc;
                        }
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < nh; i += nthr) {
                double acc = T.Ad[i] * s.H[i];
                if (i + 1 < nh) acc += T.Aup[i] * s.H[i + 1];
                if (i > 0) acc += T.Alo[i - 1] * s.H[i - 1];
                for (int b = 0; b < T.nb_row; ++b) {
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int k = i - t + 1;
                        if (k >= 0 && k < m) acc += T.Srow[((size_t)b * 3 + t) * m + k] * s.slab[(size_t)b * pitch + k];
                    }
                }
                xcol[i] = useF ? acc + P.gamma * fcol[i] : acc;
            }
        }
        // the slab and H are rewritten by the next column's TMA loads
        if (cs > 1) cluster.sync(); else __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct FusedPlan {
    int cs, mloc, pitch, threads, ctas_per_sm;
    size_t smem;
    bool tabs;
    bool ok;
};

static FusedPlan plan_fused(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T) {
    FusedPlan pl = {};
    const pop_desc &d = U->d;
    if (d.uD > FUSED_MAX_ND || d.q < 1) return pl;
    if (T && T->d.uD > FUSED_MAX_ND) return pl;
    pl.tabs = U->mD == 1;
    const size_t limit = (size_t)ctx->max_smem_optin;
    int force_cs = 0;
    if (const char *e = getenv("POP_FUSED_CS")) force_cs = atoi(e);
    int best = 0;
    for (int cs = 1; cs <= 8; cs *= 2) {
        if (T && cs > 1) break;      // the fused product keeps the whole column in one CTA
        if (force_cs && cs != force_cs) continue;
        const int mloc = (d.m + cs - 1) / cs;
        if ((int64_t)(cs - 1) * mloc >= d.m) continue;   // an empty trailing rank
        const int pitch = cs == 1 ? d.m : ((mloc + 3) & ~1);
        const size_t smem = fused_layout(nullptr, nullptr, cs, d.q, pitch, d.nh, d.m, mloc, pl.tabs, T != nullptr) + 16;
        if (smem > limit) continue;
        // smallest cluster whose slab fits; a bigger cluster only if it is forced (measured: the
        // cluster barriers cost more than the extra CTA per SM gains)
        const int per_sm = (int)std::min<size_t>(8, (limit + 1024) / (smem + 1024));
        best = cs;
        pl.cs = cs; pl.mloc = mloc; pl.pitch = pitch; pl.smem = smem; pl.ctas_per_sm = per_sm;
        break;
    }
    if (!best) return pl;
    pl.threads = std::min(512, std::max(64, ((pl.mloc + 31) / 32) * 32));
    if (const char *e = getenv("POP_FUSED_THREADS")) {
        int t = atoi(e);
        if (t >= 32 && t <= 512 && t % 32 == 0) pl.threads = t;
    }
    // registers: 64K per SM
    pl.ctas_per_sm = std::max(1, std::min(pl.ctas_per_sm, 65536 / (pl.threads * 128)));
    pl.ok = true;
    return pl;
}

bool pop_fused_supported(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T) {
    if (const char *e = getenv("POP_DISABLE_FUSED"))
        if (atoi(e)) return false;
    return plan_fused(ctx, U, T).ok;
}

template <int ND, bool SKIP1, bool MUL>
static int launch_fused_t(pop_ctx *ctx, const FusedParams &P, const FusedPlan &pl) {
    auto kern = k_fused_solve<ND, SKIP1, MUL>;
    POP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    int64_t nclusters = std::min<int64_t>(P.nrhs, (int64_t)ctx->sm_count * pl.ctas_per_sm / pl.cs);
    if (nclusters < 1) nclusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * pl.cs));
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    POP_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    ctx->launches++;
    return POP_OK;
}

template <bool MUL>
static int launch_fused_nd(pop_ctx *ctx, const FusedParams &P, const FusedPlan &pl, int nd, bool skip1) {
    if (nd == 0) return launch_fused_t<0, false, MUL>(ctx, P, pl);
    if (nd == 1) return launch_fused_t<1, false, MUL>(ctx, P, pl);
    if (skip1) return launch_fused_t<2, true, MUL>(ctx, P, pl);
    return launch_fused_t<2, false, MUL>(ctx, P, pl);
}

int pop_launch_solve_fused(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *F, int64_t ldf,
                           double *X, int64_t ldx, int64_t nrhs) {
    FusedPlan pl = plan_fused(ctx, U, T);
    POP_REQUIRE(pl.ok, POP_ERR_DIM, "fused solve: shape not supported (n=%lld, uD=%d)", (long long)U->n(), U->d.uD);
    // TMA bulk copies need 16 B aligned column starts
    if ((ldx & 1) || ((uintptr_t)X & 15)) {
        // nothing is touched before the error: the caller (pop_solve_mul_axpy) routes unaligned panels to the staged path
        POP_REQUIRE(!T, POP_ERR_DIM, "fused solve+product needs an even leading dimension and a 16-byte aligned panel");
        TriView up, lo;
        pop_make_triview(U, POP_UPPER, 0, 0, &up);
        pop_make_triview(U, POP_UPPER, 1, 0, &lo);
        return pop_launch_solve_staged(ctx, up, lo, X, 1, ldx, nrhs);
    }
    FusedParams P = {};
    pop_make_triview(U, POP_UPPER, 0, 0, &P.up);
    P.do_mul = T != nullptr;
    if (T) pop_make_mulview(T, POP_SYM_UPPER, 0, &P.T);
    P.gamma = gamma;
    P.F = F; P.ldf = ldf;
    P.X = X; P.ldx = ldx; P.nrhs = nrhs;
    P.cs = pl.cs; P.mloc = pl.mloc; P.pitch = pl.pitch;
    P.uniform_tabs = pl.tabs ? 1 : 0;
    // the solve recurrence only needs the bands of U; the product handles its own bands (<= 2)
    const int nd = U->d.uD;
    const bool skip1 = nd == 2 && U->band1_zero;
    if (T) return launch_fused_nd<true>(ctx, P, pl, nd, skip1);
    return launch_fused_nd<false>(ctx, P, pl, nd, skip1);
}
