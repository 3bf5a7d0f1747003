// Inline-PTX helpers for sm_100a: mbarriers, TMA bulk copies (global -> shared, shared -> peer
// shared), proxy fences, named barriers and the cluster barrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define POP_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {   // arrive (count 1) + expect bytes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// acquire at cluster scope: the phase is completed by bytes another CTA of the cluster sent
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy (TMA, 1-D): 16 B aligned addresses, size a multiple of 16 B
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 eviction policies for streams with no reuse (evict_first) and for data the next kernel reads back (evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
// L2 prefetch of a contiguous global range (size a multiple of 16 B)
__device__ __forceinline__ void l2_prefetch_bulk(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared window) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// shared (this CTA) -> shared (peer CTA) bulk copy, completing bytes on the PEER's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster_addr, const void *src, uint32_t bytes, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster_addr)
                 : "memory");
}
// 8-byte store into a peer CTA's shared memory that completes 8 bytes on the PEER's mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t dst_cluster_addr, double v, uint32_t bar_cluster_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster_addr),
                 "l"(__double_as_longlong(v)), "r"(bar_cluster_addr)
                 : "memory");
}
// 16-byte variant (dst 16 B aligned): two doubles in one instruction, completing 16 bytes on the PEER's mbarrier
__device__ __forceinline__ void st_async_v2f64(uint32_t dst_cluster_addr, double a, double b, uint32_t bar_cluster_addr) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst_cluster_addr),
                 "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(bar_cluster_addr)
                 : "memory");
}
// plain store into a peer CTA's shared memory (made visible by the cluster barrier)
__device__ __forceinline__ void st_cluster_f64(uint32_t dst_cluster_addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst_cluster_addr), "d"(v) : "memory");
}
// streaming store that does not allocate a line in L1 (keeps the small L1 for the read-only coupling tables)
__device__ __forceinline__ void stg_na_f64(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// Programmatic dependent launch (no-ops unless the kernel was launched with cudaLaunchAttributeProgrammaticStreamSerialization):
// launch_dependents lets the NEXT kernel of the stream start its CTAs as soon as SM resources free up, wait blocks until the
// PREVIOUS kernel has completed and its writes are visible.  A kernel may run its set-up (tables of the factor into shared
// memory, barrier initialisation) in front of the wait, under the tail of its predecessor.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {   // arrive without waiting (the other side calls named_bar_sync)
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
