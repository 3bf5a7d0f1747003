// Fused ReverseCholesky solve, register-resident ("v2"):  X <- U^{-T} U^{-1} X, optionally followed
// by the ADI product  R <- T X + gamma F, in ONE pass over HBM (16 B per unknown, 24 B with F).
//
// Replaces, for matrix right-hand sides, the column loop over
//   materialize!(::MatLdivVec{TriangularLayout{'U'}}) (src/arrowhead.jl:425-457) and
//   materialize!(::MatLdivVec{TriangularLayout{'L'}}) (src/arrowhead.jl:458-486),
// and with the product one half step of examples/adi_poisson.jl:54-55.
//
// Mapping (sm_100a):
//   * a cluster of CS CTAs owns one column at a time; CTA r owns elements [k0,k1).  One CONSUMER
//     thread per element (two when the even/odd degree chains are split, NPAR = 2) keeps the whole
//     degree chain of its element in REGISTERS between the backward and the forward sweep: the
//     register file is the on-chip home of the column (src/arrowhead.jl:439-441 and :480-482 run
//     back to back without touching memory).
//   * the columns stream into a shared-memory ring of row chunks by TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), in the order the backward sweep consumes them.
//     Once a column sits in registers (first block barrier of the column) warp 0 refills the
//     whole ring, so the loads of column c+1 (and the start of c+2) overlap the hat solve, the
//     forward sweep and the stores of column c.
//   * hats: every CTA reduces its elements' couplings to per-hat partial sums, sends them to its
//     peers with one shared->peer-shared bulk copy each (complete_tx on the peer's mbarrier: no
//     cluster barrier in the steady state) and then solves the bidiagonal hat systems redundantly
//     with a block-wide segmented affine scan (src/arrowhead.jl:449-453, :472-477).
//   * the forward sweep streams x (or T x + gamma F) to global memory straight from registers;
//     lanes run over elements, so every warp store is one run of consecutive doubles.
#include <cstdlib>

#include <mutex>
#include <unordered_map>

#include "pop_fused2.cuh"

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct F2Plan {
    bool ok;
    int rows, npar, variant;
    int threads, occ;
    size_t smem;
    f2_kernel_t kern;
    F2Params p;
};

static f2_kernel_t f2_kernel(int rows, int npar, int variant, bool mul, bool cl) {
    switch (rows) {
        case 8: return f2_kernel_r8(npar, variant, mul, cl);
        case 16: return f2_kernel_r16(npar, variant, mul, cl);
        case 24: return f2_kernel_r24(npar, variant, mul, cl);
        case 32: return f2_kernel_r32(npar, variant, mul, cl);
        case 40: return f2_kernel_r40(npar, variant, mul, cl);
        case 48: return f2_kernel_r48(npar, variant, mul, cl);
        case 56: return f2_kernel_r56(npar, variant, mul, cl);
        case 64: return f2_kernel_r64(npar, variant, mul, cl);
    }
    return nullptr;
}

static inline int ev2i(int x) { return (x + 1) & ~1; }

// registers per thread of a kernel, queried once (the planner runs on every launch)
static int f2_num_regs(const void *kern) {
    static std::mutex mu;
    static std::unordered_map<const void *, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kern);
    if (it != cache.end()) return it->second;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cache[kern] = fa.numRegs;
    return fa.numRegs;
}

static inline int hB_of_cs1(const pop_desc &d) { return d.nh; }   // a single CTA owns every hat of the column
static F2Plan plan_f2(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, bool gamma_nonzero = false) {
    F2Plan pl = {};
    const pop_desc &d = U->d;
    if (!U->is_factor || U->mD != 1 || d.uD > 2 || d.q < 1) return pl;
    if (std::min(d.nB, d.q) > F2_NBMAX) return pl;
    if (T && (T->mD != 1 || T->d.uD > 2 || std::min(T->d.nB, T->d.q) > F2_NBMAX)) return pl;
    const int nd = d.uD;
    const bool skip1 = nd == 2 && U->band1_zero;
    const int tu = T ? T->d.uD : 0;
    const bool t_even = !T || tu == 0 || (tu == 2 && T->band1_zero);   // the product never couples j and j+1
    if (nd == 0 && tu == 0) pl.variant = 0;
    else if ((nd == 0 || skip1) && t_even) pl.variant = 1;
    else pl.variant = 2;
    const int q = d.q;
    // two threads per element (even / odd degree chains) when the chains decouple: mandatory for
    // q > 64, optional below (POP_F2_NPAR=2: twice the warps per column, half the registers)
    const bool can_split = pl.variant <= 1;
    int want_npar = 0;
    if (const char *e = getenv("POP_F2_NPAR")) want_npar = atoi(e);
    if (q <= 64 && !(want_npar == 2 && can_split && q >= 2)) {
        pl.npar = 1;
        pl.rows = ((q + F2_CLASS - 1) / F2_CLASS) * F2_CLASS;
    } else if (q <= 128 && can_split) {
        pl.npar = 2;
        pl.rows = ((q + 2 * F2_CLASS - 1) / (2 * F2_CLASS)) * F2_CLASS;
    } else {
        return pl;
    }
    const int jmax = pl.rows * pl.npar;
    const int maxthreads = f2_tmax(pl.rows);
    int force_cs = 0, force_occ = 0, force_nf = 0;
    if (const char *e = getenv("POP_F2_CS")) force_cs = atoi(e);
    if (const char *e = getenv("POP_F2_OCC")) force_occ = atoi(e);
    if (const char *e = getenv("POP_F2_NF")) force_nf = atoi(e);
    const size_t limit = (size_t)ctx->max_smem_optin;
    struct { int numRegs; } fa;
    F2Plan first = {};
    for (int cs = 1; cs <= 8; cs *= 2) {
        if (T && cs > 1) break;
        if (first.ok && cs != 2 * first.p.cs) break;
        if (force_cs && cs != force_cs) continue;
        const int mloc = (d.m + cs - 1) / cs;
        if ((int64_t)(cs - 1) * mloc >= d.m) continue;
        const int ncE = ((mloc + 31) / 32) * 32;
        const int threads = ncE * pl.npar;
        if (threads > maxthreads) continue;
        if (cs > 1 && (d.m & 1)) continue;   // row parities would alternate inside a CTA's slab
        pl.kern = f2_kernel(pl.rows, pl.npar, pl.variant, T != nullptr, cs > 1);   // single-CTA columns: the kernel without the cluster code
        if (!pl.kern) continue;
        fa.numRegs = f2_num_regs((const void *)pl.kern);
        if (fa.numRegs <= 0) continue;
        F2Params &p = pl.p;
        p.cs = cs; p.mloc = mloc; p.ncE = ncE;
        p.contiguous = cs == 1 ? 1 : 0;
        p.rp = cs == 1 ? d.m : ev2i(mloc + 2);
        p.slot_stride = ev2i(F2_G * p.rp + 2);
        p.nch = jmax / F2_G;
        p.hcap = ev2i(mloc + 1) + 8;
        const size_t slot_bytes = (size_t)p.slot_stride * 8;
        // cs > 1: the set-up of the segment-map weights uses 5*hcap doubles of scratch at the start of the ring
        const size_t tail = 512 + (cs > 1 ? (size_t)std::max(0, 5 * p.hcap - p.nch * p.slot_stride) * 8 : 0);
        int o = 0;
        p.o_tab = o; o += 2 * (jmax + 2);
        p.o_tab1 = o; o += ev2i(jmax + 2);
        p.o_tT = o; o += T ? ev2i(3 * (jmax + 2)) : 0;
        p.o_hrd = o; o += p.hcap;
        p.o_hoff = o; o += p.hcap;
        p.o_H = o; o += 2 * p.hcap;
        p.o_ct = o; o += 3 * pl.npar * ncE;
        p.o_xb = o; o += T ? F2_NBMAX * ncE : 0;
        p.o_inb = o; o += F2_INB_SIZE;
        p.o_wagg = o; o += F2_WAGG;
        p.o_hw = o; o += cs > 1 ? 2 * p.hcap : 0;
        p.hat_L = (std::max(hB_of_cs1(d), 1) + 31) / 32;
        p.o_kst = o; o += (cs == 1 && !F2_OLD_SCAN) ? (10 + 3 * p.hat_L) * 32 : 0;
        o = (o + 15) & ~15;
        p.o_ring = o;
        const size_t fixed = F2_BAR_BYTES + (size_t)o * 8;
        const bool fring = T != nullptr && gamma_nonzero;
        // registers decide how many CTAs an SM can hold; shared memory is then split between them
        const int occ_regs = (int)(65536 / ((size_t)threads * ((fa.numRegs + 7) & ~7)));
        if (occ_regs < 1) continue;
        bool found = false;
        for (int occ = std::min(occ_regs, 4); occ >= 1 && !found; --occ) {
            if (force_occ && occ != force_occ) continue;
            const size_t budget = std::min(limit, (size_t)(233472 / occ) - 1024);
            // fused product: a second ring streams the F column in (at least 2 chunks, at most a column)
            const int nf_min = fring ? std::min(2, p.nch) : 0;
            if (budget < fixed + tail + slot_bytes * (size_t)(p.nch + nf_min)) continue;
            p.nslots = p.nch;   // the ring holds exactly one column (static chunk <-> slot mapping)
            p.nf = fring ? (int)std::min<size_t>(std::min(p.nch, F2_MAXFSLOTS), (budget - fixed - tail) / slot_bytes - p.nch) : 0;
            if (force_nf >= nf_min && force_nf <= p.nf) p.nf = force_nf;
            p.o_fring = p.o_ring + p.nch * p.slot_stride;
            pl.occ = occ;
            pl.threads = threads;
            pl.smem = fixed + tail + slot_bytes * (size_t)(p.nch + p.nf);
            found = true;
        }
        if (!found) continue;
        if (pl.ok) {
            // second candidate (twice the cluster size): keep it only if it doubles the CTAs per SM
            if (pl.occ >= 2) return pl;
            return first;
        }
        pl.ok = true;
        // a full-size block (512 threads x 128 registers, one CTA per SM) leaves an SM with a single phase
        // at a time; two half-size CTAs of a cluster twice as large interleave theirs (measured: cfg 2, +5 %)
        if (!force_cs && !T && pl.occ == 1 && threads > 256 && cs < 8 && !(d.m & 1)) {
            first = pl;
            continue;
        }
        break;
    }
    if (first.ok && !(pl.ok && pl.occ >= 2 && pl.p.cs == 2 * first.p.cs)) return first;
    return pl;
}

bool pop_fused2_supported(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, int64_t ldx, const double *X) {
    if (const char *e = getenv("POP_DISABLE_FUSED2"))
        if (atoi(e)) return false;
    if ((ldx & 1) || ((uintptr_t)X & 15)) return false;   // TMA bulk copies need 16 B aligned column starts
    return plan_f2(ctx, U, T, T != nullptr).ok;
}

int pop_launch_solve_fused2(pop_ctx *ctx, const pop_arrowhead *U, const pop_arrowhead *T, double gamma, const double *F, int64_t ldf,
                            double *X, int64_t ldx, int64_t nrhs) {
    F2Plan pl = plan_f2(ctx, U, T, T != nullptr && gamma != 0.0);
    POP_REQUIRE(pl.ok, POP_ERR_DIM, "fused solve v2: shape not supported (n=%lld, q=%d, uD=%d)", (long long)U->n(), U->d.q, U->d.uD);
    POP_REQUIRE(!(ldx & 1) && !((uintptr_t)X & 15), POP_ERR_DIM, "fused solve v2 needs an even leading dimension and a 16-byte aligned panel");
    if (nrhs == 0) return POP_OK;
    const pop_desc &d = U->d;
    F2Params &p = pl.p;
    const size_t plane = (size_t)d.q;   // mD == 1
    p.rdD = U->p.rdD;
    p.W1 = d.uD >= 1 ? U->p.D + (size_t)(d.lD + 1) * plane : nullptr;
    p.W2 = d.uD >= 2 ? U->p.D + (size_t)(d.lD + 2) * plane : nullptr;
    p.Sb = U->p.B;
    p.rdA = U->p.rdA;
    p.offA = U->p.Au;
    p.m = d.m; p.nh = d.nh; p.q = d.q; p.nb = std::min(d.nB, d.q);
    if (T) {
        const pop_desc &t = T->d;
        const size_t tpl = (size_t)t.q;
        p.Td0 = T->p.D + (size_t)t.lD * tpl;
        p.Td1 = t.uD >= 1 ? T->p.D + (size_t)(t.lD + 1) * tpl : nullptr;
        p.Td2 = t.uD >= 2 ? T->p.D + (size_t)(t.lD + 2) * tpl : nullptr;
        p.Trow = T->p.B;
        p.TAd = T->p.Ad;
        p.TAu = T->p.Au;
        p.nbT = std::min(t.nB, t.q);
    }
    POP_REQUIRE(!(T && gamma != 0.0) || (F && !(ldf & 1) && !((uintptr_t)F & 15)), POP_ERR_DIM,
                "fused solve+product needs an even leading dimension and a 16-byte aligned F panel");
    p.stagger_ns = 0;
    p.f_evict_first = 1;   // measured: 160.8 -> 155.4 us per ADI iteration at 1024 columns per GPU, neutral for large panels
    if (const char *e = getenv("POP_F_EVICT_FIRST")) p.f_evict_first = atoi(e);
    if (const char *e = getenv("POP_F2_STAGGER_NS")) p.stagger_ns = atoi(e);
    p.gamma = gamma;
    p.F = F ? F : X;
    p.ldf = F ? ldf : ldx;
    p.X = X; p.ldx = ldx; p.nrhs = nrhs;
    POP_CUDA(cudaFuncSetAttribute(pl.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    int64_t nclusters = std::min<int64_t>(nrhs, (int64_t)ctx->sm_count * pl.occ / p.cs);
    if (nclusters < 1) nclusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * p.cs));
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // set-up under the tail of the previous kernel (pdl_wait in the kernel)
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static int use_pdl = -1;
    if (use_pdl < 0) {
        const char *e = getenv("POP_PDL");
        use_pdl = e ? atoi(e) : 0;   // measured slower with it (profiles/r2_pdl.txt): off by default
    }
    cfg.numAttrs = 1;   // the occupancy query below wants the cluster attribute alone
    if (p.cs > 1) {
        // persistent clusters: never launch more than can be resident at once (GPC boundaries strand a few
        // SMs for cluster sizes that do not divide the GPC), a second wave would serialise whole columns
        int maxcl = 0;
        if (cudaOccupancyMaxActiveClusters(&maxcl, pl.kern, &cfg) == cudaSuccess && maxcl >= 1) {
            nclusters = std::min<int64_t>(nclusters, maxcl);
            cfg.gridDim = dim3((unsigned)(nclusters * p.cs));
        }
        cudaGetLastError();
    }
    if (const char *e = getenv("POP_F2_DEBUG")) {
        if (atoi(e)) {
            int occ_blocks = -1, occ_clusters = -1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, pl.kern, pl.threads, pl.smem);
            cudaOccupancyMaxActiveClusters(&occ_clusters, pl.kern, &cfg);
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, pl.kern);
            fprintf(stderr, "[pop_f2] rows=%d npar=%d variant=%d mul=%d cs=%d threads=%d regs=%d smem=%zu plan-occ=%d  occupancy: %d blocks/SM, %d clusters; grid=%u nf=%d\n",
                    pl.rows, pl.npar, pl.variant, T != nullptr, p.cs, pl.threads, fa.numRegs, pl.smem, pl.occ, occ_blocks, occ_clusters, cfg.gridDim.x, p.nf);
            cudaGetLastError();
        }
    }
#ifdef F2_PROF
    static unsigned long long *prof = nullptr;
    if (!prof) cudaMalloc((void **)&prof, 32 * 8);
    cudaMemsetAsync(prof, 0, 32 * 8, ctx->stream);
    p.prof = prof;
#endif
    cfg.numAttrs = use_pdl ? 2 : 1;
    POP_CUDA(cudaLaunchKernelEx(&cfg, pl.kern, p));
    ctx->launches++;
#ifdef F2_PROF
    if (getenv("POP_F2_PROF_PRINT")) {
        static const char *names[15] = {"load wait + column->regs", "backward sweep + couplings", "barrier A", "refill issue", "wait edges + hat load",
                                        "hat rhs", "barrier B", "hat scans: up", "barrier C", "correction + forward", "stores / product",
                                        "hat scans: dots + send", "hat scans: down pass 1", "hat scans: wait peers", "hat scans: fold + down pass 2"};
        unsigned long long h[32];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long tot[2] = {0, 0};
        for (int o = 0; o < 2; ++o)
            for (int i = 0; i < 15; ++i) tot[o] += h[o * 16 + i];
        fprintf(stderr, "[pop_f2 prof] cs=%d threads=%d grid=%u: cycles of block 0 (warp 0 | last warp), share of the column loop\n", p.cs, pl.threads, cfg.gridDim.x);
        for (int i = 0; i < 15; ++i)
            fprintf(stderr, "  %-30s %10llu %5.1f%% | %10llu %5.1f%%\n", names[i], h[i], 100.0 * h[i] / (double)tot[0], h[16 + i], 100.0 * h[16 + i] / (double)tot[1]);
    }
#endif
    return POP_OK;
}
