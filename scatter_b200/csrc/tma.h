// mbarrier / bulk-copy (TMA) helpers shared by the ring kernels (sm_100a).
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = tma_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared, completion counted on `bar`; source and destination 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(tma_smem_u32(bar))
                 : "memory");
}
// same with the L2 evict-first policy: for streams that are read exactly once
__device__ __forceinline__ void tma_bulk_load_stream(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     tma_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(tma_smem_u32(bar)), "l"(0x12F0000000000000ull)
                 : "memory");
}
