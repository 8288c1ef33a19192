// imma_probe.cu -- design probe for the tensor-core key switch (run on the B200):
//   tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 128, K = 32, u8 x u8 -> s32 in tensor memory, both operands K-major in shared
//   memory under no-swizzle matrix descriptors.  Checks which of (LBO, SBO) is the K-chunk stride and which the 8-row-group stride,
//   the accumulate flag, where D[m][n] lands in tensor memory, and times a stream of MMAs (4 per commit, like the four byte planes
//   of a key-switch step).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/imma_probe tools/imma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define TLD16(r, addr)                                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                      \
                 : "r"(addr) : "memory")

__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    return d;                        // layout type (bits 61..63) = 0: no swizzle
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6) = 2 (s32), a_format [7,10) = 0 (u8), b_format [10,13) = 0 (u8),
// a_major bit 15 = 0 (K), b_major bit 16 = 0 (K), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
__host__ __device__ inline uint32_t make_idesc(int M, int N) { return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ int g_timeout = 0;
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (long spin = 0; !ok; spin++) {
        if (spin > 20000000) { g_timeout = 1; return; }      // never hang the box
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

__host__ __device__ inline int a_val(int m, int k) { return (m * 7 + k * 3 + (m >> 3)) % 5; }
__host__ __device__ inline int b_val(int n, int k) { return (n * 5 + k * 11 + (n >> 2)) % 7; }

// layout mode 0: image [kchunk][rowgroup][8 rows][16 B]  -> K-chunk stride 2048, row-group stride 128
__global__ void __launch_bounds__(128) probe(int32_t* out, long long* cyc, int lbo, int sbo, int nrep) {
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t ring[4];
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm;
    unsigned char* B = sm + 4096;
    const int w = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 128 * 32; e += 128) {
        const int r = e / 32, k = e % 32;
        const int off = (k / 16) * 2048 + (r / 8) * 128 + (r % 8) * 16 + (k % 16);
        A[off] = (unsigned char)a_val(r, k);
        B[off] = (unsigned char)b_val(r, k);
    }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes of A/B visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    const uint64_t da = make_desc(smem_u32(A), lbo, sbo), db = make_desc(smem_u32(B), lbo, sbo);
    const uint32_t idesc = make_idesc(128, 128);
    uint32_t parity = 0;
    if (threadIdx.x == 0) {
        mma_i8(tbase, da, db, idesc, 0);            // D  = A B^T
        mma_i8(tbase, da, db, idesc, 1);            // D += A B^T
        mma_i8(tbase + 128, da, db, idesc, 0);      // second accumulator tile, columns 128..255
        commit(&bar);
    }
    wait_bar(&bar, parity); parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // lane = 32 w + lane holds row m = threadIdx.x; columns = n
    for (int c = 0; c < 256; c += 16) {
        uint32_t r[16];
        TLD16(r, tbase + (((uint32_t)w * 32u) << 16) + c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) out[threadIdx.x * 256 + c + i] = (int32_t)r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // throughput: nrep steps of 4 MMAs (the four byte planes of a key-switch step), one commit per step, at most 4 steps in flight
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t0 = clock64();
        for (int s = 0; s < nrep; s++) {            // one barrier per step in flight: a barrier must not run more than one phase ahead of its waiter
            if (s >= 4) wait_bar(&ring[s & 3], ((s >> 2) - 1) & 1);
            for (int p = 0; p < 4; p++) mma_i8(tbase + 128 * p, da, db, idesc, 1);
            commit(&ring[s & 3]);
        }
        for (int s = nrep - 4 < 0 ? 0 : nrep - 4; s < nrep; s++) wait_bar(&ring[s & 3], (s >> 2) & 1);
        cyc[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    int32_t* out; long long* cyc;
    cudaMalloc(&out, 128 * 256 * 4); cudaMalloc(&cyc, 8);
    int32_t* h = (int32_t*)malloc(128 * 256 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int cfg[2][2] = {{2048, 128}, {128, 2048}};       // (LBO, SBO)
    for (int c = 0; c < 2; c++) {
        cudaMemset(out, 0xff, 128 * 256 * 4);
        probe<<<1, 128, 16384>>>(out, cyc, cfg[c][0], cfg[c][1], 4096);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("LBO=%d SBO=%d: %s\n", cfg[c][0], cfg[c][1], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
        long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        int to = 0; cudaMemcpyFromSymbol(&to, g_timeout, 4); if (to) printf("  (a barrier wait timed out)\n");
        int bad2 = 0, bad1 = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 128; n++) {
                int ref = 0;
                for (int k = 0; k < 32; k++) ref += a_val(m, k) * b_val(n, k);
                if (h[m * 256 + n] != 2 * ref) bad2++;
                if (h[m * 256 + 128 + n] != ref) bad1++;
            }
        printf("LBO=%d SBO=%d: accumulated tile mismatches %d, plain tile mismatches %d of 16384; D[0][0..3] = %d %d %d %d (expect x2: %d..)  "
               "%.1f cycles per 4-MMA step\n", cfg[c][0], cfg[c][1], bad2, bad1, h[0], h[1], h[2], h[3],
               2 * [] { int r = 0; for (int k = 0; k < 32; k++) r += a_val(0, k) * b_val(0, k); return r; }(), (double)hc / 4096);
    }
    return 0;
}
