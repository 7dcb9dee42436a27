// kernels/async.cuh -- Blackwell asynchronous data movement: 1-D bulk copies (the TMA engine's non-tensor form,
// cp.async.bulk, SASS UBLKCP) global -> shared memory, completion signalled on an mbarrier with a transaction count.
// Part of the single translation unit simka_gpu.cu (included inside its anonymous namespace).
//
// Used where a CTA consumes a CONTIGUOUS run of global memory that it knows one step ahead (the records of the next
// counting bin, the segments of the next merge bin, the next tile of a sketch): one elected thread arms the barrier with the
// byte count and issues the copy, the copy engine fills shared memory while all warps work on the previous buffer -- no
// registers, no load instructions, no per-thread address arithmetic.
//
// Rules of the instruction (PTX ISA, cp.async.bulk): source and destination 16-byte aligned, size a multiple of 16 bytes.
// Callers round ranges outwards to 16 bytes and skip the head inside shared memory.
#pragma once

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barrier visible to the async proxy (the copy engine) before the first copy is issued
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes/reads of a shared buffer ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// blocks until the phase of the given parity has completed
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the barrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#else  // kernel-logic emulator (tests/emu): the copy happens at issue, the barrier counts phases, waiters yield their fiber

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t) { *bar = 0; }  // bits 0..31: completed phases, 32..63: pending bytes
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    *bar += (uint64_t)bytes << 32;
    if ((*bar >> 32) == 0) *bar = (*bar + 1) & 0xFFFFFFFFull;  // nothing to wait for
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { *bar = (*bar & ~0xFFFFFFFFull) | ((*bar + 1) & 0xFFFFFFFFull); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while ((((volatile uint64_t *)bar)[0] & 1u) == parity) emu::yield_to_scheduler();
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    if ((bytes & 15u) || ((uintptr_t)smem_dst & 15u) || ((uintptr_t)gmem_src & 15u)) {
        fprintf(stderr, "cuda_emu: bulk_g2s alignment violation (dst %p src %p bytes %u)\n", smem_dst, gmem_src, bytes);
        abort();
    }
    memcpy(smem_dst, gmem_src, bytes);
    uint64_t pend = (*bar >> 32) - bytes;
    *bar = (pend << 32) | (*bar & 0xFFFFFFFFull);
    if (pend == 0) *bar = (*bar + 1) & 0xFFFFFFFFull;
}

#endif
