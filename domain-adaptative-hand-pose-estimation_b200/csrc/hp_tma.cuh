// hp_tma.cuh - sm_100a copy-engine and launch-dependency primitives shared by the staged kernels:
// mbarrier (arm with a byte count / wait on a phase), cp.async.bulk global -> shared, L2 eviction policy,
// griddepcontrol (programmatic dependent launch).  PTX names; in SASS: SYNCS.*, UBLKCP.S.G.
#pragma once
#include "hp_common.cuh"

namespace hp {

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// plain arrival (consumer side of a full/empty pair): release semantics at CTA scope
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// named barrier over the first `n_threads` threads' warps (consumer warps of a warp-specialised block)
template <int ID, int NTHREADS>
__device__ __forceinline__ void named_barrier() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
}
// the pair barrier of stage `st` (< 4) of a block whose warp pairs share one stage each: ids 1..4 as immediates, so
// that the kernel reserves five hardware barriers, not all sixteen (0 is __syncthreads)
__device__ __forceinline__ void pair_barrier(int st) {
    if (st == 0) named_barrier<1, 64>();
    else if (st == 1) named_barrier<2, 64>();
    else if (st == 2) named_barrier<3, 64>();
    else named_barrier<4, 64>();
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded: a copy that never lands (bad pointer) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}

// the same wait for a warp whose only job is to wait (a producer warp on an empty-barrier): sleep between the polls so
// that the spin loop does not take issue slots from the computing warps of its scheduler (ncu: 8 % of all executed
// warp-instructions of the dense disparity kernel were this loop)
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (ns) __nanosleep(ns);
        if (++spins > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace hp
