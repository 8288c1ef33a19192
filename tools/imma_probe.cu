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

__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    d |= (uint64_t)(layout & 7) << 61;   // 0: no swizzle, 6: SWIZZLE_32B
    return d;
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

// mode bit 0: N = 256 per MMA (two byte planes side by side as 256 B rows), 2 MMAs per step instead of 4
// mode bit 1: SWIZZLE_32B operand layout (rows of 32 contiguous bytes, 8-row groups 256 B apart, 16-byte halves swapped in rows 4-7)
// no swizzle: image [kchunk][rowgroup][8 rows][16 B]  -> K-chunk stride rows*16, row-group stride 128
__device__ __host__ inline int img_off(int r, int k, int rows, int swz) {
    if (!swz) return (k / 16) * (rows * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 16);
    const int chunk = (k / 16) ^ ((r >> 2) & 1);
    return (r / 8) * 256 + (r % 8) * 32 + chunk * 16 + (k % 16);
}
__global__ void __launch_bounds__(128) probe(int32_t* out, long long* cyc, int mode, int nrep) {
    __shared__ uint32_t tbase_s;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t ring[4];
    __shared__ __align__(8) uint64_t ring2[16];
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
    const int NB = (mode & 1) ? 256 : 128, swz = (mode >> 1) & 1;
    unsigned char* A = sm;
    unsigned char* B = sm + 4096;
    const int w = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < 128 * 32; e += 128) { const int r = e / 32, k = e % 32; A[img_off(r, k, 128, swz)] = (unsigned char)a_val(r, k); }
    for (int e = threadIdx.x; e < NB * 32; e += 128) { const int r = e / 32, k = e % 32; B[img_off(r, k, NB, swz)] = (unsigned char)b_val(r, k); }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring[i])) : "memory");
        for (int i = 0; i < 16; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring2[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes of A/B visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    const uint64_t da = swz ? make_desc(smem_u32(A), 16, 256, 6) : make_desc(smem_u32(A), 128 * 16, 128);
    const uint64_t db = swz ? make_desc(smem_u32(B), 16, 256, 6) : make_desc(smem_u32(B), NB * 16, 128);
    const uint32_t idesc = make_idesc(128, NB);
    uint32_t parity = 0;
    if (threadIdx.x == 0) {
        mma_i8(tbase, da, db, idesc, 0);            // D  = A B^T
        mma_i8(tbase, da, db, idesc, 1);            // D += A B^T
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
    // throughput: nrep steps covering 512 accumulator columns (4 MMAs of N = 128 or 2 of N = 256), one commit per step
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t0 = clock64();
        for (int s = 0; s < nrep; s++) {            // one barrier per step in flight: a barrier must not run more than one phase ahead of its waiter
            if (s >= 4) wait_bar(&ring[s & 3], ((s >> 2) - 1) & 1);
            for (int p = 0; p < 512 / NB; p++) mma_i8(tbase + NB * p, da, db, idesc, 1);
            commit(&ring[s & 3]);
        }
        for (int s = nrep - 4 < 0 ? 0 : nrep - 4; s < nrep; s++) wait_bar(&ring[s & 3], (s >> 2) & 1);
        cyc[0] = clock64() - t0;
        // the same MMAs, one commit per `cper` steps (mode bits 4-6), at most `nfl` commit groups in flight (mode bits 8-12):
        // separates the cost of the commit itself from the depth of the queue the tensor pipe needs
        const int cper = 1 << ((mode >> 4) & 7), nfl = (mode >> 8) & 31;
        if (nfl > 0) {
            const int ncommit = nrep / cper;
            const long long t1 = clock64();
            for (int c = 0; c < ncommit; c++) {
                if (c >= nfl) wait_bar(&ring2[c % nfl], (uint32_t)((c / nfl) - 1) & 1u);
                for (int q = 0; q < cper; q++)
                    for (int p = 0; p < 512 / NB; p++) mma_i8(tbase + NB * p, da, db, idesc, 1);
                commit(&ring2[c % nfl]);
            }
            for (int c = ncommit - nfl < 0 ? 0 : ncommit - nfl; c < ncommit; c++) wait_bar(&ring2[c % nfl], (uint32_t)(c / nfl) & 1u);
            cyc[1] = clock64() - t1;
        }
    }
    // several ISSUING threads (one per warp, mode bits 16-18 = how many), step s issued by warp s % nw, one commit per step on the
    // issuer's own barriers: tcgen05.commit covers the committing thread's MMAs only, so every slot is still released by the thread that
    // filled it; the integer accumulation does not care about the order in which the tensor core takes the two streams
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nw = (mode >> 16) & 7;
    if (nw > 0) {
        const long long t2 = clock64();
        if (w < nw && (threadIdx.x & 31) == 0) {
            // zero the accumulators first (one MMA each, accumulate = 0, issued by warp 0 only) so the final sums are checkable
            if (w == 0) { for (int p = 0; p < 512 / NB; p++) mma_i8(tbase + NB * p, da, db, idesc, 0); commit(&bar); }
        }
        if (w == 0 && (threadIdx.x & 31) == 0) { wait_bar(&bar, parity); parity ^= 1; }
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (w < nw && (threadIdx.x & 31) == 0) {
            uint64_t* myring = ring2 + 4 * w;          // 4 barriers per issuer
            int k = 0;
            for (int s = w; s < nrep; s += nw, k++) {
                if (k >= 4) wait_bar(&myring[k & 3], (uint32_t)((k >> 2) - 1) & 1u);
                for (int p = 0; p < 512 / NB; p++) mma_i8(tbase + NB * p, da, db, idesc, 1);
                commit(&myring[k & 3]);
            }
            for (int c = k - 4 < 0 ? 0 : k - 4; c < k; c++) wait_bar(&myring[c & 3], (uint32_t)(c >> 2) & 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) cyc[1] = clock64() - t2;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < 256; c += 16) {
            uint32_t r[16];
            TLD16(r, tbase + (((uint32_t)w * 32u) << 16) + c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; i++) out[threadIdx.x * 256 + c + i] = (int32_t)r[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    int32_t* out; long long* cyc;
    cudaMalloc(&out, 128 * 256 * 4); cudaMalloc(&cyc, 16);
    int32_t* h = (int32_t*)malloc(128 * 256 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    const char* names[4] = {"N=128 x4, no swizzle", "N=256 x2, no swizzle", "N=128 x4, SWIZZLE_32B", "N=256 x2, SWIZZLE_32B"};
    const int cfgs[][2] = {{1, 4}, {1, 8}, {2, 4}, {4, 2}, {8, 2}, {16, 4}, {-1, 1}, {-1, 2}, {-1, 3}, {-1, 4}};      // (steps per commit, commit groups in flight) or (-1, issuing warps)
    const int ncfg = sizeof(cfgs) / sizeof(cfgs[0]);
    for (int mi = 0; mi < 4 + ncfg; mi++) {
        int mode = mi;
        if (mi >= 4 && cfgs[mi - 4][0] > 0) { int lg = 0; while ((1 << lg) < cfgs[mi - 4][0]) lg++; mode = 1 | (lg << 4) | (cfgs[mi - 4][1] << 8); }
        if (mi >= 4 && cfgs[mi - 4][0] < 0) mode = 1 | (cfgs[mi - 4][1] << 16);
        cudaMemset(out, 0xff, 128 * 256 * 4);
        probe<<<1, 128, 32768>>>(out, cyc, mode, 4096);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", names[mode], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
        long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        int to = 0; cudaMemcpyFromSymbol(&to, g_timeout, 4); if (to) printf("  (a barrier wait timed out)\n");
        const int NB = (mode & 1) ? 256 : 128;
        int bad = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < NB; n++) {
                int ref = 0;
                for (int k = 0; k < 32; k++) ref += a_val(m, k) * b_val(n, k);
                if (h[m * 256 + n] != 2 * ref) bad++;
            }
        long long hc2 = 0; cudaMemcpy(&hc2, cyc + 1, 8, cudaMemcpyDeviceToHost);
        if (mi < 4) printf("%-24s mismatches %d of %d;  %.1f cycles per 512-column step\n", names[mode & 3], bad, 128 * NB, (double)hc / 4096);
        else if ((mode >> 16) & 7) {
            int bad2 = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 256; n++) {
                    int ref = 0;
                    for (int k = 0; k < 32; k++) ref += a_val(m, k) * b_val(n, k);
                    if (h[m * 256 + n] != 4097 * ref) bad2++;
                }
            printf("N=256 x2, %d issuing warps, one commit per step: %.1f cycles per 512-column step, %d mismatches of 32768 in the final sums\n",
                   (mode >> 16) & 7, (double)hc2 / 4096, bad2);
        }
        else printf("N=256 x2, one commit per %2d steps, %2d commits (%3d steps) in flight: %.1f cycles per 512-column step\n", 1 << ((mode >> 4) & 7),
                    (mode >> 8) & 31, ((mode >> 8) & 31) << ((mode >> 4) & 7), (double)hc2 / 4096);
    }
    return 0;
}
