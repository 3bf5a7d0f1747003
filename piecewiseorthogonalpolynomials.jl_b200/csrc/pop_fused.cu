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
    int pitch;        // doubles per shared-memory row (even, >= mloc + 2)
    int uniform_tabs; // factor tables of the shared D block are staged in shared memory
};

// shared-memory carve-up (doubles unless noted)
struct Smem {
    double *slab;   // [q][pitch]
    double *H;      // [nh + 2]      (rank 0: hats of the column; all ranks when cs == 1)
    double *hloc;   // [mloc + 2]    hats k0-1 .. k1 seen by this CTA
    double *part;   // [mloc + 2]    partial sums for hats k0-1 .. k1
    double *tab;    // [(1 + nd) * q] rd, W1, W2 of the shared block
    double *hrd;    // [nh] reciprocal hat pivots
    double *hoff;   // [nh] hat super-diagonal of U (0 in the last slot)
    double *wagg;   // [128] warp aggregates of the block-wide hat scan (2 directions x (t, P) x 32)
    uint64_t *bar;
};

__device__ __forceinline__ Smem carve_smem(unsigned char *base, int q, int pitch, int nh, int mloc, int nd, bool tabs) {
    Smem s;
    double *p = reinterpret_cast<double *>(base);
    s.slab = p; p += (size_t)q * pitch;
    s.H = p; p += (nh + 3) & ~1;
    s.hloc = p; p += (mloc + 3) & ~1;
    s.part = p; p += (mloc + 3) & ~1;
    s.tab = p; p += tabs ? (((size_t)(1 + nd) * q + 1) & ~(size_t)1) : 0;
    s.hrd = p; p += (nh + 1) & ~1;
    s.hoff = p; p += (nh + 1) & ~1;
    s.wagg = p; p += 128;
    s.bar = reinterpret_cast<uint64_t *>(p);
    return s;
}

static size_t fused_smem_bytes(int q, int pitch, int nh, int mloc, int nd, bool tabs) {
    size_t d = (size_t)q * pitch + ((nh + 3) & ~1) + 2 * ((mloc + 3) & ~1) + 2 * ((nh + 1) & ~1) + 128;
    if (tabs) {
        size_t t = (size_t)(1 + nd) * q;
        d += t + (t & 1);
    }
    return d * sizeof(double) + 16;
}

template <int ND, bool MUL>
__global__ void __launch_bounds__(512) k_fused_solve(const FusedParams P) {
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
    Smem s = carve_smem(smem_raw, q, pitch, nh, P.mloc, ND, tabs);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    // parity of the global index of this CTA's first element in row j: odd0 ^ (j & modd)
    const int odd0 = (nh + k0) & 1, modd = m & 1;

    if (tid == 0) {
        mbar_init(s.bar, 1);
        fence_mbar_init();
    }
    if (tabs) {
        for (int j = tid; j < q; j += nthr) {
            s.tab[j] = tv.rdD[j];
#pragma unroll
            for (int d = 1; d <= ND; ++d) s.tab[d * q + j] = (d <= tv.nd) ? tv.W[d - 1][j] : 0.0;
        }
    }
    for (int i = tid; i < nh; i += nthr) {
        s.hrd[i] = tv.rdA[i];
        s.hoff[i] = i < nh - 1 ? tv.offA[i] : 0.0;
    }
    if (cs > 1) cluster.sync(); else __syncthreads();

    // hats of the whole column live on rank 0
    double *Hr0 = cs > 1 ? cluster.map_shared_rank(s.H, 0) : s.H;
    uint32_t phase = 0;
    const int64_t ncl = gridDim.x / cs;
    for (int64_t c = blockIdx.x / cs; c < P.nrhs; c += ncl) {
        double *xcol = P.X + c * P.ldx;
        // ---- load: aligned supersets of the row pieces (and the hats on rank 0) ------------
        if (tid == 0) {
            fence_proxy_async();
            uint32_t total = 0;
            for (int j = 0; j < q; ++j) {
                const int64_t g0 = (int64_t)nh + (int64_t)j * m + k0;
                const int64_t a0 = g0 & ~(int64_t)1, a1 = (g0 + ml + 1) & ~(int64_t)1;
                total += (uint32_t)(a1 - a0) * 8u;
            }
            const uint32_t hbytes = (uint32_t)((nh + 1) & ~1) * 8u;
            if (rank == 0) total += hbytes;
            mbar_expect_tx(s.bar, total);
            for (int j = 0; j < q; ++j) {
                const int64_t g0 = (int64_t)nh + (int64_t)j * m + k0;
                const int64_t a0 = g0 & ~(int64_t)1, a1 = (g0 + ml + 1) & ~(int64_t)1;
                tma_load_1d(s.slab + (size_t)j * pitch, xcol + a0, (uint32_t)(a1 - a0) * 8u, s.bar);
            }
            if (rank == 0) tma_load_1d(s.H, xcol, hbytes, s.bar);
        }
        while (!mbar_try_wait(s.bar, phase)) {
        }
        phase ^= 1;

        // ---- backward sweep  z_j = (b_j - sum_d W_d(j) z_{j+d}) rd_j  (src/arrowhead.jl:439-441)
        for (int e = tid; e < ml; e += nthr) {
            const int k = k0 + e;
            const int kD = mD == 1 ? 0 : k;
            double w[ND > 0 ? ND : 1];
#pragma unroll
            for (int d = 0; d < ND; ++d) w[d] = 0.0;
#pragma unroll 4
            for (int j = q - 1; j >= 0; --j) {
                const int odd = odd0 ^ (j & modd);
                double *pz = s.slab + (size_t)j * pitch + e + odd;
                double v = *pz;
#pragma unroll
                for (int d = 1; d <= ND; ++d) {
                    if (j + d < q) {
                        const double wd = tabs ? s.tab[d * q + j] : (d <= tv.nd ? tv.W[d - 1][(size_t)j * mD + kD] : 0.0);
                        v -= wd * w[d - 1];
                    }
                }
                v *= tabs ? s.tab[j] : tv.rdD[(size_t)j * mD + kD];
                *pz = v;
#pragma unroll
                for (int d = ND - 1; d >= 1; --d) w[d] = w[d - 1];
                if (ND > 0) w[0] = v;
            }
        }
        // ---- partial hat sums of this CTA: part[e'] for hat k0-1+e'  (src/arrowhead.jl:449-451)
        for (int e = tid; e < ml + 2; e += nthr) s.part[e] = 0.0;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            for (int e = tid; e < ml; e += nthr) {
                const int k = k0 + e;
                double acc = 0.0;
                for (int b = 0; b < nb; ++b) {
                    const int odd = odd0 ^ (b & modd);
                    acc += tv.S[((size_t)b * 3 + t) * m + k] * s.slab[(size_t)b * pitch + e + odd];
                }
                s.part[e + t] += acc;
            }
            __syncthreads();
        }
        if (cs > 1) cluster.sync();
        // ---- rank 0: assemble the hat right-hand side, both bidiagonal sweeps ----------------
        if (rank == 0) {
            for (int i = tid; i < nh; i += nthr) {
                double v = s.H[i];
                for (int r = 0; r < cs; ++r) {
                    const int rk0 = r * P.mloc;
                    if (rk0 >= m) break;
                    const int rml = min(m, rk0 + P.mloc) - rk0;
                    const int e = i - (rk0 - 1);
                    if (e >= 0 && e < rml + 2) {
                        const double *rp = (cs > 1 && r != 0) ? cluster.map_shared_rank(s.part, r) : s.part;
                        v -= rp[e];
                    }
                }
                s.H[i] = v;
            }
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
        // ---- every CTA: fetch its hats, correct the coupled bubbles (src/arrowhead.jl:475-477)
        for (int e = tid; e < ml + 2; e += nthr) {
            const int i = k0 - 1 + e;
            s.hloc[e] = (i >= 0 && i < nh) ? Hr0[i] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < ml; e += nthr) {
            const int k = k0 + e;
            for (int b = 0; b < nb; ++b) {
                const int odd = odd0 ^ (b & modd);
                double v = s.slab[(size_t)b * pitch + e + odd];
#pragma unroll
                for (int t = 0; t < 3; ++t) v -= tv.S[((size_t)b * 3 + t) * m + k] * s.hloc[e + t];
                s.slab[(size_t)b * pitch + e + odd] = v;
            }
        }
        // each thread corrected only its own elements: the forward sweep below needs no barrier

        if (!MUL) {
            // hats out (rank 0), then forward sweep streaming x to global memory
            if (rank == 0)
                for (int i = tid; i < nh; i += nthr) xcol[i] = s.H[i];
            for (int e = tid; e < ml; e += nthr) {
                const int k = k0 + e;
                const int kD = mD == 1 ? 0 : k;
                double w[ND > 0 ? ND : 1];
#pragma unroll
                for (int d = 0; d < ND; ++d) w[d] = 0.0;
                double *xo = xcol + nh + k;
#pragma unroll 4
                for (int j = 0; j < q; ++j) {
                    const int odd = odd0 ^ (j & modd);
                    double v = s.slab[(size_t)j * pitch + e + odd];
#pragma unroll
                    for (int d = 1; d <= ND; ++d) {
                        if (j - d >= 0) {
                            const double wd = tabs ? s.tab[d * q + j - d] : (d <= tv.nd ? tv.W[d - 1][(size_t)(j - d) * mD + kD] : 0.0);
                            v -= wd * w[d - 1];
                        }
                    }
                    v *= tabs ? s.tab[j] : tv.rdD[(size_t)j * mD + kD];
                    xo[(int64_t)j * m] = v;
#pragma unroll
                    for (int d = ND - 1; d >= 1; --d) w[d] = w[d - 1];
                    if (ND > 0) w[0] = v;
                }
            }
        } else {
            // ---- fused product (cs == 1):  r = T x + gamma f, T symmetric with bands <= ND -----
            const MulView &T = P.T;
            const int tD = T.mD;
            const bool useF = P.gamma != 0.0;
            const double *fcol = useF ? P.F + c * P.ldf : nullptr;
            // forward sweep keeping a window x[j-2ND .. j]; emits r[j-ND] once x[j] is known
            for (int e = tid; e < ml; e += nthr) {
                const int k = k0 + e;
                const int kD = mD == 1 ? 0 : k;
                const int kT = tD == 1 ? 0 : k;
                double xw[2 * ND + 1];   // xw[i] = x[j - i]
#pragma unroll
                for (int i = 0; i < 2 * ND + 1; ++i) xw[i] = 0.0;
                for (int j = 0; j < q + ND; ++j) {
                    double v = 0.0;
                    if (j < q) {
                        const int odd = odd0 ^ (j & modd);
                        v = s.slab[(size_t)j * pitch + e + odd];
#pragma unroll
                        for (int d = 1; d <= ND; ++d) {
                            if (j - d >= 0) {
                                const double wd = tabs ? s.tab[d * q + j - d] : (d <= tv.nd ? tv.W[d - 1][(size_t)(j - d) * mD + kD] : 0.0);
                                v -= wd * xw[d - 1];
                            }
                        }
                        v *= tabs ? s.tab[j] : tv.rdD[(size_t)j * mD + kD];
                        if (j < T.nb_col || j < T.nb_row) s.slab[(size_t)j * pitch + e + odd] = v;   // needed by the hat rows
                    }
#pragma unroll
                    for (int i = 2 * ND; i >= 1; --i) xw[i] = xw[i - 1];
                    xw[0] = v;
                    const int jo = j - ND;   // output row, its window is xw[ND - d] = x[jo + d]
                    if (jo >= 0) {
                        double acc = 0.0;
#pragma unroll
                        for (int d = -ND; d <= ND; ++d) {
                            const int jj = jo + d;
                            if (d >= -T.lD && d <= T.uD && jj >= 0 && jj < q) acc += T.Dp[d + POP_MAXD][(ptrdiff_t)jo * tD + kT] * xw[ND - d];
                        }
                        if (jo < T.nb_col) {
#pragma unroll
                            for (int t = 0; t < 3; ++t) acc += T.Scol[((size_t)jo * 3 + t) * m + k] * s.hloc[e + t];
                        }
                        const int64_t g = (int64_t)nh + (int64_t)jo * m + k;
                        xcol[g] = useF ? acc + P.gamma * fcol[g] : acc;
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < nh; i += nthr) {
                double acc = T.Ad[i] * s.H[i];
                if (i + 1 < nh) acc += T.Aup[i] * s.H[i + 1];
                if (i > 0) acc += T.Alo[i - 1] * s.H[i - 1];
                for (int b = 0; b < T.nb_row; ++b) {
                    const int odd = odd0 ^ (b & modd);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int k = i - t + 1;
                        if (k >= 0 && k < m) acc += T.Srow[((size_t)b * 3 + t) * m + k] * s.slab[(size_t)b * pitch + (k - k0) + odd];
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
    const int nd = d.uD;
    if (nd > FUSED_MAX_ND || d.q < 1) return pl;
    if (T) {
        if (T->d.uD > FUSED_MAX_ND) return pl;
    }
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
        const int pitch = (mloc + 3) & ~1;
        const size_t smem = fused_smem_bytes(d.q, pitch, d.nh, mloc, FUSED_MAX_ND, pl.tabs);
        if (smem > limit) continue;
        // prefer the smallest cluster that still leaves >= 2 CTAs per SM
        const int per_sm = (int)std::min<size_t>(8, (limit + 1024) / (smem + 1024));
        if (!best || (pl.ctas_per_sm < 2 && per_sm > pl.ctas_per_sm)) {
            best = cs;
            pl.cs = cs; pl.mloc = mloc; pl.pitch = pitch; pl.smem = smem; pl.ctas_per_sm = per_sm;
        }
        if (force_cs) break;
    }
    if (!best) return pl;
    pl.threads = std::min(512, std::max(64, ((pl.mloc + 31) / 32) * 32));
    pl.ok = true;
    return pl;
}

bool pop_fused_supported(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T) {
    if (const char *e = getenv("POP_DISABLE_FUSED"))
        if (atoi(e)) return false;
    return plan_fused(ctx, U, T).ok;
}

template <int ND, bool MUL>
static int launch_fused_t(pop_ctx *ctx, const FusedParams &P, const FusedPlan &pl) {
    auto kern = k_fused_solve<ND, MUL>;
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

int pop_launch_solve_fused(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *F, int64_t ldf,
                           double *X, int64_t ldx, int64_t nrhs) {
    FusedPlan pl = plan_fused(ctx, U, T);
    POP_REQUIRE(pl.ok, POP_ERR_DIM, "fused solve: shape not supported (n=%lld, uD=%d)", (long long)U->n(), U->d.uD);
    // TMA bulk copies need 16 B aligned column starts
    if ((ldx & 1) || ((uintptr_t)X & 15)) {
        TriView up, lo;
        pop_make_triview(U, POP_UPPER, 0, 0, &up);
        pop_make_triview(U, POP_UPPER, 1, 0, &lo);
        int rc = pop_launch_solve_staged(ctx, up, lo, X, 1, ldx, nrhs);
        if (rc || !T) return rc;
        pop_set_error("fused solve+product needs an even leading dimension and a 16-byte aligned panel");
        return POP_ERR_DIM;
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
    const int nd = std::max(U->d.uD, T ? T->d.uD : 0);
    if (T) {
        if (nd == 0) return launch_fused_t<0, true>(ctx, P, pl);
        if (nd == 1) return launch_fused_t<1, true>(ctx, P, pl);
        return launch_fused_t<2, true>(ctx, P, pl);
    }
    if (nd == 0) return launch_fused_t<0, false>(ctx, P, pl);
    if (nd == 1) return launch_fused_t<1, false>(ctx, P, pl);
    return launch_fused_t<2, false>(ctx, P, pl);
}
