// tma.cuh -- bulk asynchronous copies (cp.async.bulk, the 1-D form of the Tensor Memory Accelerator path: SASS UBLKCP) and the
// mbarrier handshakes that go with them, as thin inline-PTX wrappers for sm_100a.
//
// Pattern used by the kernels of this library (producer / consumer ring of shared-memory stages):
//   producer (one elected thread):  mbar_wait(empty[s], phase ^ 1);  mbar_arrive_expect_tx(full[s], bytes);
//                                   bulk_g2s(stage_s, global_src, bytes, full[s]);
//   consumers:                      mbar_wait(full[s], phase);  ... read the stage from shared memory ...
//                                   __syncwarp();  lane 0: mbar_arrive(empty[s]);
// Requirements of cp.async.bulk: source, destination and size are multiples of 16 bytes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dge
{

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrive_count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrive_count) : "memory");
}

// make the barrier initialisation visible to the async proxy (the bulk-copy engine) before the first copy is issued
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// try_wait suspends the thread in hardware for a bounded time; loop until the phase with this parity has completed
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// global -> shared bulk copy, completion signalled on `bar` with complete_tx(bytes)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// L2 eviction policy for data that is read exactly once (keeps the streamed records from evicting the barcode table)
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}

} // namespace dge
