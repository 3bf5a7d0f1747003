// One-pass row half step of the 2D ADI on ONE GPU:  rows of R <- T (S^{-1} r) + gamma f,  S = U U^T
// (X / F followed by X * A, examples/adi_poisson.jl:54), 24 B per entry instead of the 40 B of the
// two-pass row kernels in pop_row.cu.
//
// A block of RB consecutive rows of the n x n matrix (RB = 16: every access is a 128 B run) is owned by a
// thread-block CLUSTER of cs CTAs for the whole half step.  CTA c owns the mesh elements [c*mlc, (c+1)*mlc) of
// the row direction and their hats; thread (e, r) keeps the whole degree chain of element e in row r in
// REGISTERS between the backward sweep (src/arrowhead.jl:439-441), the hat solve (:449-453, :472-477) and
// the forward sweep (:480-482), exactly like the column kernel (pop_fused2.cuh) keeps a column.  The hat
// system of a row is distributed over the cluster: per row ONE exchange of two dot products (the zero-
// incoming segment maps of the two bidiagonal recurrences, see pop_fused2.cuh) through distributed shared
// memory, then the reference recurrences run on the CTA's own hats with the exact incoming values.
// F streams into shared memory by cp.async while the chains run; the next row block is pulled into L2.
// Instantiated per chain class in pop_row1_rNN.cu.
#pragma once
#include "pop_internal.h"
#include "pop_ptx.cuh"

#define R1_NBMAX 4

struct R1Params {
    // factor U (upper data, shared D block)
    const double *rd, *w1, *w2;    // [q]
    const double *S;               // [nb][3][m]
    const double *rdA, *offA;      // [nh], [nh-1]
    int nb;
    // product operator T = Symmetric(upper data); nbT < 0: plain solve
    const double *t0, *t1, *t2;
    const double *Ts;
    const double *TAd, *TAu;
    int nbT;
    double gamma;
    const double *F;               // null: no gamma*F term
    int64_t ldf;
    double *X;
    int64_t ld, nrows;
    int m, nh, q;
    int cs, mlc, hcap, nblk, stage_f;
    int contig;   // 1: cluster c owns the contiguous range of row blocks [c*per, (c+1)*per); 0: round robin
};

typedef void (*r1_kernel_t)(const R1Params);
// jmax = chain class (16, 32, 48, 64); npar = 2: two threads per (element, row), even / odd degrees (needs decoupled chains)
r1_kernel_t r1_kernel_r16(int rb, int npar, bool has1);
r1_kernel_t r1_kernel_r32(int rb, int npar, bool has1);
r1_kernel_t r1_kernel_r48(int rb, int npar, bool has1);
r1_kernel_t r1_kernel_r64(int rb, int npar, bool has1);
// largest block of a kernel whose threads keep `rows` chain entries in registers
__host__ __device__ constexpr int r1_tmax(int rows) { return rows <= 32 ? 512 : rows <= 48 ? 320 : 256; }

// shared-memory carve-up (doubles); the same function sizes the launch on the host
struct R1Smem {
    int tab, ttab, tw1, tt1, hrd, hoffl, hwd, hwu, pc, hb, ct, xb, inA, inD, inB, wagg, fs, total;
};
#define R1_BAR_BYTES 64   // 6 mbarriers at the start of shared memory
__host__ __device__ inline R1Smem r1_smem(int rows, int rb, int mlc, int hcap, int nthr, bool stage_f, int npar) {
    R1Smem s;
    int o = 0;
    s.tab = o; o += 2 * (rows + 2);
    s.ttab = o; o += 2 * (rows + 2);
    s.tw1 = o; o += rows + 2;
    s.tt1 = o; o += rows + 2;
    s.hrd = o; o += hcap + 2;
    s.hoffl = o; o += hcap + 2;
    s.hwd = o; o += hcap + 2;
    s.hwu = o; o += hcap + 2;
    s.pc = o; o += 24;
    s.hb = o; o += (hcap + 3) * rb;
    s.ct = o; o += 2 * 3 * mlc * rb;        // [parity][slot][element][row]
    s.xb = o; o += R1_NBMAX * mlc * rb;
    s.inA = o; o += 8 * rb;                 // [2 buffers][2 sides][2 parities][rb]
    s.inD = o; o += 8 * rb;
    o = (o + 1) & ~1;
    s.inB = o; o += 2 * 8 * rb * 2;         // [2 parities][8 ranks][rb][2]  (16 B pairs)
    s.wagg = o; o += 32 * rb;
    o = (o + 1) & ~1;
    s.fs = o; o += stage_f ? (rows / npar) * nthr : 0;   // [chain entry][thread]
    s.total = o;
    return s;
}

#ifdef __CUDACC__
__device__ __forceinline__ void cp_async_f64(double *dst_smem, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// p + bytes through an opaque instruction (see the load loop of k_row1)
template <typename T>
__device__ __forceinline__ T *ptr_step(T *p, int64_t bytes) {
    asm volatile("add.s64 %0, %0, %1;" : "+l"(p) : "l"(bytes));
    return p;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ROWS chain entries per thread; NPAR = 2: thread (par, e, r) owns the degrees j = 2i + par of element e in row r
// (the even and odd degree chains decouple when the first super-diagonal of every D block is zero, true for
// K + sigma M); NPAR = 1 handles the general two-band case.
template <int ROWS, int RB, int NPAR, bool HAS1>
__global__ void __launch_bounds__(r1_tmax(ROWS), 1) k_row1(const __grid_constant__ R1Params P) {
    static_assert(!(HAS1 && NPAR == 2), "split chains need a zero first super-diagonal");
    extern __shared__ __align__(16) unsigned char r1_smem_raw[];
    constexpr int JMAX = ROWS * NPAR;
    constexpr int JSAFE = JMAX - 16;   // degrees j < JSAFE exist for every q of the class (JMAX-16 < q <= JMAX)
    constexpr int NBI = (R1_NBMAX + NPAR - 1) / NPAR;
    uint64_t *bars = reinterpret_cast<uint64_t *>(r1_smem_raw);
    uint64_t *barA = bars, *barB = bars + 2, *barD = bars + 4;   // [2 parities] each
    double *sm = reinterpret_cast<double *>(r1_smem_raw + R1_BAR_BYTES);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    const int cs = P.cs;
    const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int m = P.m, nh = P.nh, q = P.q, mlc = P.mlc;
    const bool mul = P.nbT >= 0;
    const bool useF = mul && P.F != nullptr && P.gamma != 0.0;
    const R1Smem L = r1_smem(JMAX, RB, mlc, P.hcap, nthr, P.stage_f != 0, NPAR);
    double2 *tab = reinterpret_cast<double2 *>(sm + L.tab);     // (rd_j, w2_j), zero beyond q
    double2 *ttab = reinterpret_cast<double2 *>(sm + L.ttab);   // (t0_j, t2_j)
    double *tw1 = sm + L.tw1, *tt1 = sm + L.tt1;
    double *hrd = sm + L.hrd, *hoffl = sm + L.hoffl, *hwd = sm + L.hwd, *hwu = sm + L.hwu, *Pc = sm + L.pc;
    double *Hs = sm + L.hb + RB;   // Hs[u*RB + r] = hat hA+u of row r; u = -1 and u = nhl hold the neighbours' hats
    double *ct = sm + L.ct, *xb = sm + L.xb, *inA = sm + L.inA, *inD = sm + L.inD, *inB = sm + L.inB, *wagg = sm + L.wagg;
    double *fs = sm + L.fs;

    const int k0 = rank * mlc;
    const int hA = k0, hB = rank == cs - 1 ? nh : min(nh, k0 + mlc);
    const int nhl = max(hB - hA, 0);
    const int ncl = gridDim.x / cs;
    const int cl = blockIdx.x / cs;

    // ---- one-time set-up ---------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
    }
    for (int j = tid; j < JMAX + 2; j += nthr) {
        const bool in = j < q;
        tab[j] = make_double2(in ? P.rd[j] : 0.0, (in && P.w2 && j + 2 < q) ? P.w2[j] : 0.0);
        tw1[j] = (in && P.w1 && j + 1 < q) ? P.w1[j] : 0.0;
        ttab[j] = make_double2((in && mul) ? P.t0[j] : 0.0, (in && mul && P.t2 && j + 2 < q) ? P.t2[j] : 0.0);
        tt1[j] = (in && mul && P.t1 && j + 1 < q) ? P.t1[j] : 0.0;
    }
    for (int u = tid; u <= nhl + 1; u += nthr) {
        const int i = hA + u;
        hrd[u] = (u <= nhl && i < nh) ? P.rdA[i] : 0.0;
        hoffl[u] = (u <= nhl && i >= 1 && i < nh) ? P.offA[i - 1] : 0.0;   // couples local hats u-1 and u
    }
    __syncthreads();
    if (cs > 1) {
        if (tid == 0) {
            // zero-incoming segment maps as linear functionals of the hat right-hand sides (derivation: pop_fused2.cuh)
            //   a_u = -off(u,u+1) rd_u,  b_u = -off(u-1,u) rd_u
            const int n = nhl;
            double pa = 1.0;
            for (int k = 0; k < n; ++k) {
                hwd[k] = hrd[k] * pa;                       // rd_k prod_{u<k} a_u
                pa *= -hoffl[k + 1] * hrd[k];
            }
            double pb = 1.0;
            for (int k = n - 1; k >= 0; --k) {
                hwu[k] = hrd[k] * pb;                       // wup_k = rd_k prod_{u>k} b_u
                pb *= -hoffl[k] * hrd[k];
            }
            double A = 1.0, kap = 0.0;
            for (int k = n - 1; k >= 0; --k) {
                A *= -hoffl[k + 1] * hrd[k];                // prod_{u>=k} a_u
                kap = fma(hwu[k], A, kap);
            }
            double c = 0.0;
            for (int j = 0; j < n; ++j) {
                c = hwu[j] + (j > 0 ? -hoffl[j] * hrd[j - 1] : 0.0) * c;   // c_j = wup_j + a_{j-1} c_{j-1}
                hwu[j] = hrd[j] * c;
            }
            for (int r = 0; r < cs; ++r) {
                st_cluster_f64(mapa_u32(Pc + rank, r), pa);
                st_cluster_f64(mapa_u32(Pc + 8 + rank, r), pb);
                st_cluster_f64(mapa_u32(Pc + 16 + rank, r), kap);
            }
        }
        cluster_arrive();   // also: every CTA's mbarriers are initialised before a peer sends to them
        cluster_wait();
    }

    // thread <-> (element e of this CTA, row r of the block); lanes run over the rows first
    const int par = NPAR == 1 ? 0 : min(tid / (mlc * RB), NPAR - 1);
    const int tp = tid - par * mlc * RB;
    const int e = tp / RB, r = tp % RB;
    const bool act = tid < NPAR * mlc * RB;
    const int el = act ? e : mlc - 1;
    const int k = k0 + el;
    const double *sbk = P.S + k;
    const int64_t ld = P.ld, js = (int64_t)m * ld;
    const int64_t fjs = (int64_t)m * P.ldf;
    const int64_t jsb = js * 8 * NPAR, fjsb = fjs * 8 * NPAR;   // byte strides between the chain entries of a thread
    const double2 *tabp = tab + par, *ttabp = ttab + par;
    const bool edgeL = act && e == 0 && rank > 0, edgeR = act && e == mlc - 1 && rank < cs - 1;
    const uint32_t expA = cs > 1 ? (uint32_t)(8 * RB * NPAR * ((rank > 0) + (rank < cs - 1))) : 0u;
    const uint32_t expB = (uint32_t)(16 * RB * (cs - 1));
    const bool hatw = act && par == 0;                                // the hats are handled by the even-parity threads
    const bool hat0 = hatw && e < nhl, hat1 = hatw && e + mlc < nhl;   // this thread's hats: u = e and (last CTA of a C1 mesh) u = e + mlc

    // contiguous ranges keep the clusters' concurrent accesses 'per' row blocks apart (the columns of a row block
    // are a power-of-two stride apart: neighbouring blocks in flight would crowd the same memory channels)
    const int per = (P.nblk + ncl - 1) / ncl;
    const int blk0 = P.contig ? cl * per : cl, blk1 = P.contig ? min(P.nblk, blk0 + per) : P.nblk, bstep = P.contig ? 1 : ncl;
    int t = 0;
    for (int blk = blk0; blk < blk1; blk += bstep, ++t) {
        const int pb = t & 1;
        const uint32_t ph = (uint32_t)(t >> 1) & 1u;
        const int64_t row = (int64_t)blk * RB + r;
        const bool rowok = row < P.nrows;
        const bool rv = rowok && act;
        const int64_t rowc = rowok ? row : P.nrows - 1;   // loads of rows past the end are clamped, their results dropped
        const int64_t cb = (int64_t)(nh + k) * ld + (int64_t)par * js;
        // ---- P1: the row block -> registers (hats too); F -> shared memory (asynchronously) ----------------
        double z[ROWS];
        {
            // the pointer walks from degree to degree through an opaque add: otherwise the compiler hoists all ROWS
            // loop-invariant offsets j*js out of the block loop into registers the chain needs
            const double *p = P.X + rowc + cb + (int64_t)(ROWS - 1) * NPAR * js;
#pragma unroll
            for (int i = ROWS - 1; i >= 0; --i) {
                z[i] = (NPAR * i + NPAR - 1 < JSAFE || NPAR * i + par < q) ? *p : 0.0;
                p = ptr_step(p, -jsb);
            }
        }
        const double hreg0 = hat0 ? P.X[rowc + (int64_t)(hA + e) * ld] : 0.0;
        const double hreg1 = hat1 ? P.X[rowc + (int64_t)(hA + e + mlc) * ld] : 0.0;
        if (useF && P.stage_f) {
            const double *fp = P.F + rowc + (int64_t)(nh + k) * P.ldf + (int64_t)par * fjs;
            double *fd = fs + tid;
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                if (NPAR * i + NPAR - 1 < JSAFE || NPAR * i + par < q) cp_async_f64(fd + i * nthr, fp);
                fp = ptr_step(fp, fjsb);
            }
            cp_async_commit();
        }
        // ---- P2: backward sweep  z_j = (b_j - w1_j z_{j+1} - w2_j z_{j+2}) rd_j ; couplings to the hats ----
#pragma unroll
        for (int i = ROWS - 1; i >= 0; --i) {
            const double2 cf = tabp[NPAR * i];
            double v = z[i];
            if (NPAR == 1) {
                if (HAS1 && i + 1 < ROWS) v = fma(-tw1[i], z[i + 1], v);
                if (i + 2 < ROWS) v = fma(-cf.y, z[i + 2], v);
            } else {
                if (i + 1 < ROWS) v = fma(-cf.y, z[i + 1], v);
            }
            z[i] = v * cf.x;
        }
        {
            double c0 = 0.0, c1 = 0.0, c2 = 0.0;
#pragma unroll
            for (int i = 0; i < NBI; ++i) {
                const int b = NPAR * i + par;
                if (i < ROWS && b < P.nb) {
                    const double *sb = sbk + (size_t)b * 3 * m;
                    c0 = fma(__ldg(sb), z[i], c0);
                    c1 = fma(__ldg(sb + m), z[i], c1);
                    c2 = fma(__ldg(sb + 2 * m), z[i], c2);
                }
            }
            if (act) {
                ct[((par * 3 + 0) * mlc + e) * RB + r] = c0;
                ct[((par * 3 + 1) * mlc + e) * RB + r] = c1;
                ct[((par * 3 + 2) * mlc + e) * RB + r] = c2;
            }
            // my edge elements also couple to a neighbour's hat (round A)
            if (edgeL) st_async_f64(mapa_u32(inA + ((pb * 2 + 1) * 2 + par) * RB + r, rank - 1), c0, mapa_u32(barA + pb, rank - 1));
            if (edgeR) st_async_f64(mapa_u32(inA + ((pb * 2 + 0) * 2 + par) * RB + r, rank + 1), c2, mapa_u32(barA + pb, rank + 1));
            if (hat0) Hs[e * RB + r] = hreg0;
            if (hat1) Hs[(e + mlc) * RB + r] = hreg1;
        }
        if (expA && tid == 0) mbar_expect_tx(barA + pb, expA);
        __syncthreads();
        if (expA) mbar_wait(barA + pb, ph);
        // ---- P3: right-hand sides of my hats; the two zero-incoming segment maps as dot products -----------
        {
            double sd = 0.0, su = 0.0;
            if (hatw)
                for (int u = e; u < nhl; u += mlc) {
                    double v = 0.0;
#pragma unroll
                    for (int pp = 0; pp < NPAR; ++pp) {
                        if (u + 1 < mlc) v += ct[((pp * 3 + 0) * mlc + u + 1) * RB + r];
                        else if (u + 1 == mlc && rank < cs - 1) v += inA[((pb * 2 + 1) * 2 + pp) * RB + r];
                        if (u < mlc) v += ct[((pp * 3 + 1) * mlc + u) * RB + r];
                        if (u >= 1) {
                            if (u - 1 < mlc) v += ct[((pp * 3 + 2) * mlc + u - 1) * RB + r];
                        } else if (rank > 0) {
                            v += inA[((pb * 2 + 0) * 2 + pp) * RB + r];
                        }
                    }
                    const double h = Hs[u * RB + r] - v;
                    Hs[u * RB + r] = h;
                    sd = fma(hwd[u], h, sd);
                    su = fma(hwu[u], h, su);
                }
            if (cs > 1) {
#pragma unroll
                for (int o = RB; o < 32; o <<= 1) {
                    sd += __shfl_xor_sync(POP_FULL_MASK, sd, o);
                    su += __shfl_xor_sync(POP_FULL_MASK, su, o);
                }
                if (lane < RB) {
                    wagg[(warp * 2 + 0) * RB + lane] = sd;
                    wagg[(warp * 2 + 1) * RB + lane] = su;
                }
            }
        }
        __syncthreads();
        if (cs > 1) {
            // round B: my segment's pair of maps -> every peer (16 B per row and peer)
            for (int idx = tid; idx < RB * cs; idx += nthr) {
                const int p = idx / RB, rr = idx % RB;
                if (p != rank) {
                    double Td = 0.0, Tu = 0.0;
                    for (int w = 0; w < nw; ++w) {
                        Td += wagg[(w * 2 + 0) * RB + rr];
                        Tu += wagg[(w * 2 + 1) * RB + rr];
                    }
                    st_async_v2f64(mapa_u32(inB + ((pb * 8 + rank) * RB + rr) * 2, p), Td, Tu, mapa_u32(barB + pb, p));
                }
            }
            if (tid == 0) mbar_expect_tx(barB + pb, expB);
        }
        // the next row block on its way into L2 while the hats are solved
        if (act && blk + bstep < blk1 && row + (int64_t)bstep * RB < P.nrows) {
            const double *np = P.X + row + (int64_t)bstep * RB + (int64_t)(nh + k) * ld + (int64_t)(r * NPAR + par) * js;
            for (int j = r * NPAR + par; j < q; j += RB * NPAR) {
                prefetch_l2(np);
                np += (int64_t)RB * NPAR * js;
            }
        }
        // ---- P4: the hat recurrences of my segment, one thread per row (:453 down, :472 up) ----------------
        if (tid < RB) {
            double Gdn = 0.0, Gup = 0.0;
            if (cs > 1) {
                double Tdm = 0.0, Tum = 0.0;
                for (int w = 0; w < nw; ++w) {
                    Tdm += wagg[(w * 2 + 0) * RB + tid];
                    Tum += wagg[(w * 2 + 1) * RB + tid];
                }
                mbar_wait(barB + pb, ph);
                double g = 0.0, mult = 1.0;
                for (int p = cs - 1; p >= 0; --p) {
                    const double2 tp = *reinterpret_cast<const double2 *>(inB + ((pb * 8 + p) * RB + tid) * 2);
                    const double Tdp = p == rank ? Tdm : tp.x, Tup = p == rank ? Tum : tp.y;
                    if (p == rank) Gdn = g;
                    if (p < rank) {
                        Gup = fma(mult, fma(Pc[16 + p], g, Tup), Gup);
                        mult *= Pc[8 + p];
                    }
                    g = fma(Pc[p], g, Tdp);
                }
            }
            double *col = Hs + tid;
            double hin = Gdn;
            for (int b0 = nhl - 1; b0 >= 0; b0 -= 8) {
                double v[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) v[s] = b0 - s >= 0 ? col[(b0 - s) * RB] : 0.0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int u = b0 - s;
                    if (u >= 0) {
                        hin = (v[s] - hoffl[u + 1] * hin) * hrd[u];
                        col[u * RB] = hin;
                    }
                }
            }
            hin = Gup;
            for (int b0 = 0; b0 < nhl; b0 += 8) {
                double v[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) v[s] = b0 + s < nhl ? col[(b0 + s) * RB] : 0.0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int u = b0 + s;
                    if (u < nhl) {
                        hin = (v[s] - hoffl[u] * hin) * hrd[u];
                        col[u * RB] = hin;
                    }
                }
            }
            // the neighbours' hats next to my segment: hat hA-1 entered the upward sweep; hat hB follows from the
            // value that entered the downward sweep and my last hat
            col[-RB] = rank > 0 ? Gup : 0.0;
            col[nhl * RB] = (hB < nh && nhl > 0) ? (Gdn - hoffl[nhl] * hin) * hrd[nhl] : 0.0;
            col[(nhl + 1) * RB] = 0.0;
        }
        __syncthreads();
        // ---- P5: correct the coupled bubbles (:475-477), forward sweep (:480-482), product, stores -------------
        const double h0 = Hs[(el - 1) * RB + r], h1 = Hs[el * RB + r], h2 = Hs[(el + 1) * RB + r];
#pragma unroll
        for (int i = 0; i < NBI; ++i) {
            const int b = NPAR * i + par;
            if (i < ROWS && b < P.nb) {
                const double *sb = sbk + (size_t)b * 3 * m;
                double v = z[i];
                v = fma(-__ldg(sb), h0, v);
                v = fma(-__ldg(sb + m), h1, v);
                v = fma(-__ldg(sb + 2 * m), h2, v);
                z[i] = v;
            }
        }
        double *xo = P.X + row + cb;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            double v = z[i];
            if (NPAR == 1) {
                if (HAS1 && i >= 1) v = fma(-tw1[i - 1], z[i - 1], v);
                if (i >= 2) v = fma(-tab[i - 2].y, z[i - 2], v);
            } else {
                if (i >= 1) v = fma(-tabp[NPAR * (i - 1)].y, z[i - 1], v);
            }
            z[i] = v * tabp[NPAR * i].x;
            if (!mul) {
                if (rv && (NPAR * i + NPAR - 1 < JSAFE || NPAR * i + par < q)) *xo = z[i];
                xo = ptr_step(xo, jsb);
            }
        }
        if (!mul) {
            if (rowok) {
                if (hat0) P.X[row + (int64_t)(hA + e) * ld] = Hs[e * RB + r];
                if (hat1) P.X[row + (int64_t)(hA + e + mlc) * ld] = Hs[(e + mlc) * RB + r];
            }
        } else {
            // the hat rows of the product read x of the coupled bubbles; my edge elements feed a neighbour's hat (round D)
            {
                double eL = 0.0, eR = 0.0;
#pragma unroll
                for (int i = 0; i < NBI; ++i) {
                    const int b = NPAR * i + par;
                    if (i < ROWS && b < P.nbT) {
                        const double *tb = P.Ts + (size_t)b * 3 * m + k;
                        if (act) xb[(b * mlc + e) * RB + r] = z[i];
                        eL = fma(__ldg(tb), z[i], eL);
                        eR = fma(__ldg(tb + 2 * m), z[i], eR);
                    }
                }
                if (edgeL) st_async_f64(mapa_u32(inD + ((pb * 2 + 1) * 2 + par) * RB + r, rank - 1), eL, mapa_u32(barD + pb, rank - 1));
                if (edgeR) st_async_f64(mapa_u32(inD + ((pb * 2 + 0) * 2 + par) * RB + r, rank + 1), eR, mapa_u32(barD + pb, rank + 1));
                if (expA && tid == 0) mbar_expect_tx(barD + pb, expA);
            }
            if (useF && P.stage_f) cp_async_wait_all();
            const double *fp = P.F + rowc + (int64_t)(nh + k) * P.ldf + (int64_t)par * fjs;
            const double *fd = fs + tid;
#pragma unroll
            for (int i0 = 0; i0 < ROWS; i0 += 8) {
                double fq[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int i = i0 + s;
                    fq[s] = 0.0;
                    if (useF && (NPAR * i + NPAR - 1 < JSAFE || NPAR * i + par < q)) fq[s] = P.stage_f ? fd[i * nthr] : *fp;
                    if (!P.stage_f) fp = ptr_step(fp, fjsb);
                }
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int i = i0 + s;
                    const double2 tc = ttabp[NPAR * i];
                    double acc = tc.x * z[i];
                    if (NPAR == 1) {
                        if (HAS1) {
                            if (i + 1 < ROWS) acc = fma(tt1[i], z[i + 1], acc);
                            if (i >= 1) acc = fma(tt1[i - 1], z[i - 1], acc);
                        }
                        if (i + 2 < ROWS) acc = fma(tc.y, z[i + 2], acc);
                        if (i >= 2) acc = fma(ttab[i - 2].y, z[i - 2], acc);
                    } else {
                        if (i + 1 < ROWS) acc = fma(tc.y, z[i + 1], acc);
                        if (i >= 1) acc = fma(ttabp[NPAR * (i - 1)].y, z[i - 1], acc);
                    }
                    if (i < NBI && NPAR * i + par < P.nbT) {
                        const double *tb = P.Ts + (size_t)(NPAR * i + par) * 3 * m + k;
                        acc = fma(__ldg(tb), h0, acc);
                        acc = fma(__ldg(tb + m), h1, acc);
                        acc = fma(__ldg(tb + 2 * m), h2, acc);
                    }
                    if (rv && (NPAR * i + NPAR - 1 < JSAFE || NPAR * i + par < q)) *xo = fma(P.gamma, fq[s], acc);
                    xo = ptr_step(xo, jsb);
                }
            }
            __syncthreads();   // xb complete
            if (expA) mbar_wait(barD + pb, ph);
            // hat rows:  T_A h + sum_b T_b x_b (+ the neighbours' edge elements) + gamma f
            if (hatw && rowok)
                for (int u = e; u < nhl; u += mlc) {
                    const int i = hA + u;
                    double acc = P.TAd[i] * Hs[u * RB + r];
                    if (i + 1 < nh) acc = fma(P.TAu[i], Hs[(u + 1) * RB + r], acc);
                    if (i > 0) acc = fma(P.TAu[i - 1], Hs[(u - 1) * RB + r], acc);
                    for (int b = 0; b < P.nbT; ++b)
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            const int kl = u - s + 1;
                            if (kl >= 0 && kl < mlc) acc = fma(__ldg(P.Ts + ((size_t)b * 3 + s) * m + k0 + kl), xb[(b * mlc + kl) * RB + r], acc);
                        }
#pragma unroll
                    for (int pp = 0; pp < NPAR; ++pp) {
                        if (u == 0 && rank > 0) acc += inD[((pb * 2 + 0) * 2 + pp) * RB + r];
                        if (u == mlc - 1 && rank < cs - 1) acc += inD[((pb * 2 + 1) * 2 + pp) * RB + r];
                    }
                    P.X[row + (int64_t)i * ld] = useF ? fma(P.gamma, P.F[row + (int64_t)i * P.ldf], acc) : acc;
                }
        }
        __syncthreads();   // Hs, ct, xb are rewritten by the next block
    }
    if (cs > 1) {
        cluster_arrive();   // no CTA exits while a peer may still send into its shared memory
        cluster_wait();
    }
}

// one translation unit per chain class JMAX: one thread per chain with / without the first super-diagonal, and the
// split-chain variant
template <int JMAX>
static r1_kernel_t r1_pick(int rb, int npar, bool has1) {
    if (npar == 2) {
        if (rb == 16) return k_row1<JMAX / 2, 16, 2, false>;
        if (rb == 8) return k_row1<JMAX / 2, 8, 2, false>;
        return nullptr;
    }
    if (rb == 16) return has1 ? k_row1<JMAX, 16, 1, true> : k_row1<JMAX, 16, 1, false>;
    if (rb == 8) return has1 ? k_row1<JMAX, 8, 1, true> : k_row1<JMAX, 8, 1, false>;
    return nullptr;
}
#define R1_DEFINE_CLASS(JMAX) \
    r1_kernel_t r1_kernel_r##JMAX(int rb, int npar, bool has1) { return r1_pick<JMAX>(rb, npar, has1); }
#endif  // __CUDACC__
