// tmem_transpose_probe.cu -- the 16 x 32 transpose of the warp-resident FFT through TENSOR MEMORY instead of shared memory.
// (design probe; mapping derivation in profiles/r1_notes.md)  Elements e = 32 m + s (m = register 0..15, s = lane), 4 words each.
// Two round trips  st.32x32b -> ld.16x256b  move element bits (e8..e5) from registers to lanes and (e4..e1) from lanes to
// registers; afterwards lane t' = (e0 e6 e5 e8 e7) holds the 16 elements u = (e4 e3 e2 e1) of its node.  The inverse runs the
// mirrored  st.16x256b -> ld.32x32b  twice.  Prints the number of mismatches and the cycles of one forward + inverse pair.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define ST32_16(r, o, addr) asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" \
    ::"r"(r[o+0]),"r"(r[o+1]),"r"(r[o+2]),"r"(r[o+3]),"r"(r[o+4]),"r"(r[o+5]),"r"(r[o+6]),"r"(r[o+7]),"r"(r[o+8]),"r"(r[o+9]),"r"(r[o+10]),"r"(r[o+11]),"r"(r[o+12]),"r"(r[o+13]),"r"(r[o+14]),"r"(r[o+15]),"r"(addr) : "memory")
#define LD32_16(r, o, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
    : "=r"(r[o+0]),"=r"(r[o+1]),"=r"(r[o+2]),"=r"(r[o+3]),"=r"(r[o+4]),"=r"(r[o+5]),"=r"(r[o+6]),"=r"(r[o+7]),"=r"(r[o+8]),"=r"(r[o+9]),"=r"(r[o+10]),"=r"(r[o+11]),"=r"(r[o+12]),"=r"(r[o+13]),"=r"(r[o+14]),"=r"(r[o+15]) : "r"(addr) : "memory")
#define LD16_8(r, o, addr) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(r[o+0]),"=r"(r[o+1]),"=r"(r[o+2]),"=r"(r[o+3]),"=r"(r[o+4]),"=r"(r[o+5]),"=r"(r[o+6]),"=r"(r[o+7]),"=r"(r[o+8]),"=r"(r[o+9]),"=r"(r[o+10]),"=r"(r[o+11]),"=r"(r[o+12]),"=r"(r[o+13]),"=r"(r[o+14]),"=r"(r[o+15]),"=r"(r[o+16]),"=r"(r[o+17]),"=r"(r[o+18]),"=r"(r[o+19]),"=r"(r[o+20]),"=r"(r[o+21]),"=r"(r[o+22]),"=r"(r[o+23]),"=r"(r[o+24]),"=r"(r[o+25]),"=r"(r[o+26]),"=r"(r[o+27]),"=r"(r[o+28]),"=r"(r[o+29]),"=r"(r[o+30]),"=r"(r[o+31]) : "r"(addr) : "memory")
#define ST16_8(r, o, addr) asm volatile("tcgen05.st.sync.aligned.16x256b.x8.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" \
    ::"r"(r[o+0]),"r"(r[o+1]),"r"(r[o+2]),"r"(r[o+3]),"r"(r[o+4]),"r"(r[o+5]),"r"(r[o+6]),"r"(r[o+7]),"r"(r[o+8]),"r"(r[o+9]),"r"(r[o+10]),"r"(r[o+11]),"r"(r[o+12]),"r"(r[o+13]),"r"(r[o+14]),"r"(r[o+15]),"r"(r[o+16]),"r"(r[o+17]),"r"(r[o+18]),"r"(r[o+19]),"r"(r[o+20]),"r"(r[o+21]),"r"(r[o+22]),"r"(r[o+23]),"r"(r[o+24]),"r"(r[o+25]),"r"(r[o+26]),"r"(r[o+27]),"r"(r[o+28]),"r"(r[o+29]),"r"(r[o+30]),"r"(r[o+31]),"r"(addr) : "memory")
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// forward: W[4m + w] (lane s) -> L[32 h' + 4 kk' + 2 r1' + r0'] (lane t'), see header
__device__ __forceinline__ void transpose_fwd(uint32_t (&W)[64], const uint32_t ts) {
    uint32_t A[64];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 16; j++) A[16 * k + j] = W[4 * (4 * k + ((j >> 1) & 3)) + (j & 1) + 2 * ((j >> 3) & 1)];
#pragma unroll
    for (int k = 0; k < 4; k++) ST32_16(A, 16 * k, ts + 16 * k);
    wait_st(); __syncwarp();
    uint32_t L[64];
    LD16_8(L, 0, ts); LD16_8(L, 32, ts + (16u << 16));
    wait_ld(); __syncwarp();
    // second round: word L[32 h + 4 kk + 2 r1 + r0] -> column r0 + 2 kk1 + 4 kk2 + 8 kk0 + 16 r1 + 32 h
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int h = q >> 1, r1 = q & 1, r0 = j & 1, kk1 = (j >> 1) & 1, kk2 = (j >> 2) & 1, kk0 = (j >> 3) & 1;
            A[16 * q + j] = L[32 * h + 4 * (4 * kk2 + 2 * kk1 + kk0) + 2 * r1 + r0];
        }
#pragma unroll
    for (int q = 0; q < 4; q++) ST32_16(A, 16 * q, ts + 16 * q);
    wait_st(); __syncwarp();
    LD16_8(W, 0, ts); LD16_8(W, 32, ts + (16u << 16));
    wait_ld(); __syncwarp();
}
__device__ __forceinline__ void transpose_inv(uint32_t (&W)[64], const uint32_t ts) {
    ST16_8(W, 0, ts); ST16_8(W, 32, ts + (16u << 16));
    wait_st(); __syncwarp();
    uint32_t A[64], L[64];
#pragma unroll
    for (int q = 0; q < 4; q++) LD32_16(A, 16 * q, ts + 16 * q);
    wait_ld(); __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int h = q >> 1, r1 = q & 1, r0 = j & 1, kk1 = (j >> 1) & 1, kk2 = (j >> 2) & 1, kk0 = (j >> 3) & 1;
            L[32 * h + 4 * (4 * kk2 + 2 * kk1 + kk0) + 2 * r1 + r0] = A[16 * q + j];
        }
    ST16_8(L, 0, ts); ST16_8(L, 32, ts + (16u << 16));
    wait_st(); __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; k++) LD32_16(A, 16 * k, ts + 16 * k);
    wait_ld(); __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 16; j++) W[4 * (4 * k + ((j >> 1) & 3)) + (j & 1) + 2 * ((j >> 3) & 1)] = A[16 * k + j];
}

__global__ void __launch_bounds__(256) probe(int* bad_fwd, int* bad_inv, long long* cyc, int iters) {
    __shared__ uint32_t tbase_s;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t ts = tbase_s + (((uint32_t)(w & 3) * 32u) << 16) + (uint32_t)(w >> 2) * 64u;
    uint32_t W[64];
#pragma unroll
    for (int m = 0; m < 16; m++)
#pragma unroll
        for (int x = 0; x < 4; x++) W[4 * m + x] = (uint32_t)(((32 * m + lane) << 2) | x) + 1000000u * w;
    transpose_fwd(W, ts);
    // expected: lane t' = 16 p + 8 b1 + 4 b0 + 2 b3 + b2 holds element e = 32 b + p + 2 u, word x at 32 u1 + 16 u3 + 8 u2 + 4 x1 + 2 u0 + x0
    const int p = lane >> 4, b = (((lane >> 1) & 1) << 3) | ((lane & 1) << 2) | (((lane >> 3) & 1) << 1) | ((lane >> 2) & 1);
    int bad = 0;
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const int idx = 32 * ((u >> 1) & 1) + 16 * ((u >> 3) & 1) + 8 * ((u >> 2) & 1) + 4 * (x >> 1) + 2 * (u & 1) + (x & 1);
            bad += W[idx] != (uint32_t)(((32 * b + p + 2 * u) << 2) | x) + 1000000u * w;
        }
    if (bad) atomicAdd(bad_fwd, bad);
    transpose_inv(W, ts);
    bad = 0;
#pragma unroll
    for (int m = 0; m < 16; m++)
#pragma unroll
        for (int x = 0; x < 4; x++) bad += W[4 * m + x] != (uint32_t)(((32 * m + lane) << 2) | x) + 1000000u * w;
    if (bad) atomicAdd(bad_inv, bad);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) { transpose_fwd(W, ts); transpose_inv(W, ts); }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / iters;
    if (W[lane & 63] == 0xdeadbeef) cyc[1] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
}
int main() {
    int *d; long long* dc; cudaMalloc(&d, 8); cudaMalloc(&dc, 16); cudaMemset(d, 0, 8);
    probe<<<148, 256>>>(d, d + 1, dc, 200);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    int h[2]; long long hc[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    printf("forward mismatches %d, inverse mismatches %d; forward + inverse transposes of 8 KB per warp, 8 warps per SM: %lld cycles per pair\n", h[0], h[1], hc[0]);
    return 0;
}
