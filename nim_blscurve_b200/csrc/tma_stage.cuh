// tma_stage.cuh — staging of the SignatureSet tile of a thread block through the TMA bulk-copy engine.
//
// The input of the path is an AoS array of 320-byte records (bls_batch_verifier.nim:34: pk 96 | msg 32 | sig 192).  A
// block that works on `count` consecutive sets owns count * 320 contiguous, 16-byte aligned bytes: ONE 1-D bulk copy
// (cp.async.bulk.shared::cluster.global, SASS UBLKCP) brings them into shared memory with full-line HBM/L2 requests and
// no register staging, completion signalled on an mbarrier; every thread then picks the fields of its own set out of
// shared memory (stride 320 B = 80 words: at most a two-way bank conflict on 16-byte accesses).  Without it each thread
// issues its own 32- to 96-byte strided loads (sector-granular, 30-90 % of every 128-byte line unused per request).
#pragma once
#include <stdint.h>

namespace bls {
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Called by EVERY thread of the block (contains __syncthreads): thread 0 arms the barrier and issues the bulk copy of
// `bytes` (multiple of 16) from gsrc (16-byte aligned) to smem_dst (16-byte aligned); all threads return once the bytes
// have landed.  `bar` is one 8-byte shared-memory word used for this single phase.
__device__ __forceinline__ void tma_stage_tile(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    const uint32_t b = smem_u32(bar), d = smem_u32(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b) : "memory");
    }
#endif
}

// The same for ONE FIELD of the records: thread t < count issues its own bulk copy of `row_bytes` (multiple of 16) from
// gsrc + t * g_stride to smem_dst + t * row_bytes, all completing on the one barrier that thread 0 armed with the total.
// For kernels that need 32 of the 320 bytes (the message): 4 KB of shared memory per block instead of the 40 KB tile,
// which would take L1 capacity away from kernels that live on their local-memory stacks.
__device__ __forceinline__ void tma_stage_rows(void *smem_dst, uint32_t row_bytes, const void *gsrc, size_t g_stride,
                                               uint32_t count, uint64_t *bar) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
    const uint32_t b = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(count * row_bytes) : "memory");
    }
    __syncthreads();
    if (threadIdx.x < count) {
        const uint32_t d = smem_u32((const uint8_t *)smem_dst + (size_t)threadIdx.x * row_bytes);
        const uint8_t *src = (const uint8_t *)gsrc + (size_t)threadIdx.x * g_stride;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(src), "r"(row_bytes), "r"(b) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b) : "memory");
    }
#endif
}

#endif
}  // namespace bls
