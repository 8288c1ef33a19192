// tmem_cp_probe.cu -- what does tcgen05.cp.32x128b.warpx4 deliver?  (design probe, run on the B200)
// Fills shared memory with word w at byte 4w holding the value w, copies rows with tcgen05.cp under a no-swizzle matrix
// descriptor, and prints which shared-memory word every (TMEM lane, column) received.  Also times a 16 KB smem -> TMEM copy.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define TLD4(r, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr))

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    return d;                        // swizzle bits 61..63 = 0: no swizzle
}

__global__ void __launch_bounds__(128) probe(uint32_t* out, long long* cyc, int sbo, int lbo, int ncopies) {
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar;
    extern __shared__ __align__(1024) uint32_t sm[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += 128) sm[i] = i;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // zero the first 16 columns of every lane so untouched cells are visible
    {
        uint32_t z[4] = {0xdeadbeef, 0xdeadbeef, 0xdeadbeef, 0xdeadbeef};
        for (int c = 0; c < 4; c++)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};" ::"r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]),
                         "r"(tbase + (((uint32_t)w * 32u) << 16) + 4 * c) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t0 = clock64();
        for (int k = 0; k < ncopies; k++) {
            const uint64_t desc = make_desc(smem_u32(sm) + 512 * k, lbo, sbo);
            asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tbase + 4 * (k % 32)), "l"(desc) : "memory");
        }
        const long long ti = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        cyc[1] = ti - t0;                 // cycles the issuing thread spent on the copy instructions alone
        cyc[2] = clock64() - ti;          // ... and on the commit
    }
    // everyone waits for the copy
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    if (threadIdx.x == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 4; c++) {
        uint32_t r[4];
        TLD4(r, tbase + (((uint32_t)w * 32u) << 16) + 4 * c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; i++) out[(threadIdx.x) * 16 + 4 * c + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    uint32_t* d; long long* dc;
    cudaMalloc(&d, 128 * 16 * 4); cudaMalloc(&dc, 32);
    uint32_t h[128 * 16]; long long hc;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int cfg[][3] = {{128, 0, 1}, {128, 16, 1}, {256, 0, 1}, {128, 128, 4}};
    for (auto& c : cfg) {
        probe<<<1, 128, 32768 + 1024>>>(d, dc, c[0], c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo=%d lbo=%d: %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        printf("== SBO=%d LBO=%d copies=%d  (%lld cycles issue->visible)\n", c[0], c[1], c[2], hc);
        for (int lane : {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 96, 127}) {
            printf("  tmem lane %3d:", lane);
            for (int k = 0; k < 16; k++) { if (h[lane * 16 + k] == 0xdeadbeef) printf("    -"); else printf(" %4u", h[lane * 16 + k]); }
            printf("\n");
        }
    }
    // timing: 32 copies = 16 KB
    probe<<<1, 128, 32768 + 1024>>>(d, dc, 128, 0, 32);
    cudaDeviceSynchronize();
    long long h3[3];
    cudaMemcpy(h3, dc, 24, cudaMemcpyDeviceToHost);
    printf("32 x (32x128b.warpx4) = 16 KB smem -> TMEM: %lld cycles issue->visible; issuing thread busy %lld cycles on the copies, %lld on the commit\n", h3[0], h3[1], h3[2]);
    return 0;
}
