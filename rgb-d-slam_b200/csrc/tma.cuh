// Minimal TMA (cp.async.bulk.tensor) + mbarrier wrappers for sm_100a, inline PTX only.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace rs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
}

// 3-D tiled TMA load: box lands densely in shared memory, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(smem_dst)),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
            : "memory");
}

// L2 cache policy for data that is read exactly once (the depth image): evict it first, so that what the kernel WRITES
// (the per-cell records the next two kernels read back) stays resident in the 126 MB L2 instead of being pushed out by
// 315 MB of streamed input.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t policy)
{
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                    smem_u32(smem_dst)),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
            : "memory");
}

// Variants that take shared-memory addresses already converted with smem_u32 (hot loops convert once).
__device__ __forceinline__ void mbar_arrive_expect_tx_a(const uint32_t bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(const uint32_t bar, const uint32_t parity)
{
    asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint_a(const uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2,
                                                   const uint32_t bar, uint64_t policy)
{
    asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                    smem_dst),
            "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
            : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

}  // namespace rs
