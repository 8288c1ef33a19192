// bk_pipe.cuh -- mbarrier / TMA primitives and the bootstrapping-key pipeline into tensor memory (KeyPipe below).
//
// Step i of a blind rotation uses BK_i (tfhe_blindRotate_FFT passes bkFFT+i, cb/lwe_functions.cpp:352), one spectrum polynomial
// BK_i[p][q] per multiply-accumulate.
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

// ---------------------------------------------------------------------------------------------------------------
// KeyPipe: the bootstrapping key reaches the multiply-accumulates through TENSOR MEMORY, once per CTA.
//
// Registers can be filled through the load/store unit (128 B/clk/SM, shared with every transpose, twiddle and accumulator
// access of the kernel) or from tensor memory (tcgen05.ld, its own datapath).  Key values are 64 KB per CMUX per accumulator;
// loading them per warp through the LSU was 25 % of all LSU wavefronts and pinned 128 registers per thread for the prefetch.
// Here all warps of a CTA walk the key in lockstep, chunk by chunk (chunk = the two spectra BK_i[p][0..1], 16 KB for N=1024):
//     global --TMA bulk copy--> 16 KB shared staging --tcgen05.cp.32x128b.warpx4--> 128 TMEM columns, broadcast to the four
//     lane quarters --tcgen05.ld--> registers of every warp, 16 columns at a time inside the multiply-accumulate.
// tcgen05.cp with a no-swizzle descriptor (SBO = 128) maps staging row r (16 bytes at byte 16 r) to TMEM lane r and 4 columns
// (tools/tmem_cp_probe.cu, profiles/tmem_cp_probe_r1.txt), and our spectral layout keeps the 32 lanes of a slot contiguous,
// so one copy per slot moves a key polynomial with no reformatting.  Protocol (chunk k, phase parity k & 1):
//     full  : mbarrier armed by tcgen05.commit after the copies of chunk k      -> acquire() waits on it
//     done  : warps that finished reading chunk k
//     tma   : mbarrier (complete_tx) of the staging buffer
//     poll(): any warp with a moment to spare issues what is due -- the TMA of the next chunk once the staging buffer is free,
//             the copies of a landed chunk once every warp has released the previous one (claimed by compare-and-swap)
// Nobody ever waits for a warp that is merely slow: a chunk is needed one whole forward transform after the previous one
// was released.  A CMUX skipped because bara == 0 still acquires/releases its chunks so that the counts stay aligned.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_add_acq_rel(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
// shared-memory matrix descriptor, no swizzle: rows of 16 bytes, 8-row groups SBO bytes apart
__device__ __forceinline__ uint64_t tc_desc_rows16(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

struct KeyPipeShared {          // one per CTA, 64 bytes
    uint64_t tma_bar, full_bar;
    uint32_t done, tma_next, copy_next, active;
};
struct KeyPipe {
    KeyPipeShared* sh;
    unsigned char* stage;       // 16-byte rows, chunk_bytes
    const unsigned char* gkey;  // chunk 0 in global memory
    uint32_t tkey;              // TMEM address of the key columns in this warp's lane quarter
    uint32_t tkey0;             // same columns, lane field 0 (destination of the broadcasting copies)
    uint32_t chunk_bytes, nchunks;
    uint32_t k;                 // chunk this warp is at

    // one thread: staging -> tensor memory, then arm `full`.  The issuing thread is busy ~70 cycles per copy (2 200 per chunk,
    // profiles/tmem_cp_probe_r1.txt), which is why the duty goes to whoever has slack (poll) and never to the last warp out.
    __device__ __forceinline__ void copy_to_tmem() const {
        const uint32_t rows = chunk_bytes / 512;                    // one 32x128b copy per 512 bytes = 4 columns
        for (uint32_t r = 0; r < rows; r++) {
            const uint64_t desc = tc_desc_rows16(smem_u32(stage) + 512u * r, 128u);
            asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tkey0 + 4u * r), "l"(desc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sh->full_bar)) : "memory");
    }
    // lane 0 of any warp, at any time: move the key stream forward if something is due.
    //   TMA of chunk n  (n = tma_next)  is due when the copies of chunk n-1 have completed (staging buffer free)
    //   copy of chunk c (c = copy_next) is due when every warp has released chunk c-1 and the TMA of chunk c has landed
    __device__ __forceinline__ void poll() const {
        const uint32_t n = *(volatile uint32_t*)&sh->tma_next;
        if (n < nchunks && mbar_test_wait(&sh->full_bar, (n - 1) & 1)) {
            if (atomicCAS(&sh->tma_next, n, n + 1) == n) {
                mbar_expect_tx(&sh->tma_bar, chunk_bytes);
                tma_load_1d(stage, gkey + (size_t)n * chunk_bytes, chunk_bytes, &sh->tma_bar);
            }
        }
        const uint32_t c = *(volatile uint32_t*)&sh->copy_next;
        if (c < nchunks && *(volatile uint32_t*)&sh->done == sh->active && *(volatile uint32_t*)&sh->tma_next > c &&
            mbar_test_wait(&sh->tma_bar, c & 1)) {
            if (atomicCAS(&sh->copy_next, c, c + 1) == c) {
                sh->done = 0;                       // nobody releases chunk c before `full` of chunk c, armed below
                __threadfence_block();
                tc_fence_after();
                copy_to_tmem();
            }
        }
    }
    // thread 0 of the CTA, once, after the barriers are initialised (tma_next = copy_next = 1)
    __device__ __forceinline__ void prologue() const {
        mbar_expect_tx(&sh->tma_bar, chunk_bytes);
        tma_load_1d(stage, gkey, chunk_bytes, &sh->tma_bar);
        while (!mbar_try_wait(&sh->tma_bar, 0)) {}
        tc_fence_after();
        copy_to_tmem();
    }
    // all lanes of a warp: chunk k is readable in tensor memory on return
    __device__ __forceinline__ void acquire(const int lane) const {
        while (!mbar_test_wait(&sh->full_bar, k & 1)) {
            if (lane == 0) poll();
            __syncwarp();
        }
        if (lane == 0) poll();
        tc_fence_after();
    }
    // all lanes of a warp, after their last tcgen05.wait::ld on chunk k
    __device__ __forceinline__ void release(const int lane) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) smem_add_acq_rel(&sh->done, 1);
        k++;
    }
};

}  // namespace tfhe_b200
