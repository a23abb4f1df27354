// rb_tma.cuh -- the few PTX primitives the streaming kernels need: mbarrier + 1-D bulk async copies
// (cp.async.bulk, executed by the TMA unit; SASS UBLKCP).  sm_90+ PTX, used here on sm_100a.
#pragma once
#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifndef RB_DI
#define RB_DI __device__ __forceinline__
#endif

RB_DI uint32_t rb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

RB_DI void rb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rb_smem_addr(bar)), "r"(count) : "memory");
}
// Makes the barrier initialisation visible to the async proxy before the first bulk copy targets it.
RB_DI void rb_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// One arrival + the number of bytes the bulk copies of this phase will deliver.
RB_DI void rb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rb_smem_addr(bar)), "r"(bytes) : "memory");
}

RB_DI void rb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RB_DONE_%=;\n"
        "bra RB_WAIT_%=;\n"
        "RB_DONE_%=:\n"
        "}\n" ::"r"(rb_smem_addr(bar)), "r"(parity) : "memory");
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is reported to `bar`.
RB_DI void rb_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rb_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(rb_smem_addr(bar)) : "memory");
}
