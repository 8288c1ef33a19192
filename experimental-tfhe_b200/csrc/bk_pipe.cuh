// bk_pipe.cuh -- bootstrapping-key spectra: global (L2 resident) -> shared memory by TMA bulk copies, per lane group.
//
// Step i of a blind rotation uses BK_i (tfhe_blindRotate_FFT passes bkFFT+i, cb/lwe_functions.cpp:352), one spectrum polynomial
// BK_i[p][q] per multiply-accumulate.  A group has no registers left to hold key values in flight (two FP64 accumulators own
// 128 of its 255 registers), so an L2 round trip in front of every multiply-accumulate would be fully exposed.  Instead each
// polynomial is requested with one `cp.async.bulk` (SASS UBLKCP) a transform-half ahead of its use and lands in shared
// memory: BK_i[p][1] in a dedicated buffer at the start of forward transform p, BK_i[p][0] in the group's own transpose
// buffer as soon as the transpose of transform p is done with it.  Completion is signalled on per-group mbarriers
// (complete_tx); groups never wait for each other.
//
// (A CTA-wide ring shared by all 8 warps was tried first: 48 KB is less than one CMUX worth of key (64 KB), the fastest
// warp ran into the ring's end and the whole CTA convoyed at ~1/3 of the throughput -- profiles/r1_notes.md.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`  (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// generic-proxy accesses to a shared-memory buffer (the transpose) are ordered before a following bulk copy into it
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One landing buffer + its mbarrier; `uses` counts completed fills (phase parity).  Owned by one lane group.
struct BkSlot {
    uint64_t* bar;
    unsigned char* dst;
    uint32_t uses;
    // one lane, after the group has synchronised on its previous reads of dst
    __device__ __forceinline__ void request(const void* src, uint32_t bytes) const {
        fence_proxy_async_smem();
        mbar_expect_tx(bar, bytes);
        tma_load_1d(dst, src, bytes, bar);
    }
    // all lanes
    __device__ __forceinline__ void wait() {
        while (!mbar_try_wait(bar, uses & 1)) {}
        uses++;
    }
};

}  // namespace tfhe_b200
