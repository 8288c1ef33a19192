// tmem_probe.cu -- can tensor memory serve as spill space for FP64 accumulators?  (design probe, run on the B200)
// Measures tcgen05.ld / tcgen05.st throughput from ordinary warps (no MMA involved) and whether they share a pipe with
// LDS or DFMA.  32x32b shape: lane i of warp w touches TMEM lane 32*(w%4)+i, N consecutive 32-bit columns.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define TLD16(r, addr)                                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
                 : "r"(addr))
#define TST16(r, addr)                                                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"       \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr) : "memory")

// MODE 0: ld only; 1: st only; 2: ld+fma+st (accumulator round trip); 3: ld+st + 16 LDS.128; 4: ld+st + 64 DFMA; 5: LDS only ref; 6: correctness
template <int MODE> __global__ void __launch_bounds__(256) k(double* out, int iters, int* err) {
    __shared__ uint32_t tbase_s;
    __shared__ double2 sm[1024];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_double2(i, 1.0);
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // this warp's window: lane quarter w%4, 128 columns starting at (w/4)*128
    const uint32_t taddr = tbase + (((uint32_t)(w & 3) * 32u) << 16) + (uint32_t)(w >> 2) * 128u;
    uint32_t r[16];
    for (int i = 0; i < 16; i++) r[i] = lane * 1000 + w * 100 + i;
    double acc = 0; double2 s = make_double2(0, 0);
    double f0 = lane, f1 = lane + 1, f2 = lane + 2, f3 = lane + 3;
    if (MODE == 6) {
        for (int c = 0; c < 8; c++) { for (int i = 0; i < 16; i++) r[i] = lane * 100000 + w * 1000 + c * 16 + i; TST16(r, taddr + c * 16); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        int bad = 0;
        for (int c = 0; c < 8; c++) {
            TLD16(r, taddr + c * 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; i++) bad += r[i] != (uint32_t)(lane * 100000 + w * 1000 + c * 16 + i);
        }
        if (bad) atomicAdd(err, bad);
    } else {
        TST16(r, taddr);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {      // 8 chunks of 16 columns = 128 columns = one pass over the accumulators
                if (MODE == 0 || MODE >= 2 && MODE != 5) { TLD16(r, taddr + c * 16); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
                if (MODE == 2) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        double d = __hiloint2double(r[i + 1], r[i]);
                        d = fma(d, 1.0000001, 1e-9);
                        r[i] = __double2loint(d); r[i + 1] = __double2hiint(d);
                    }
                }
                if (MODE == 3 || MODE == 5) {
#pragma unroll
                    for (int u = 0; u < 2; u++) { double2 v = sm[(threadIdx.x + 256 * u + it + c) & 1023]; s.x += v.x; s.y += v.y; }
                }
                if (MODE == 4) {
#pragma unroll
                    for (int u = 0; u < 2; u++) { f0 = fma(f0, 1.0000001, 1e-9); f1 = fma(f1, 1.0000001, 1e-9); f2 = fma(f2, 1.0000001, 1e-9); f3 = fma(f3, 1.0000001, 1e-9); }
                }
                if (MODE == 1 || MODE >= 2 && MODE != 5) { TST16(r, taddr + c * 16); }
            }
            if (MODE != 0 && MODE != 5) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        for (int i = 0; i < 16; i++) acc += r[i];
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc + s.x + s.y + f0 + f1 + f2 + f3;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE> void run(const char* name, double bytes_per_iter_per_thread) {
    double* out; int* err; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    const int iters = 2048;
    k<MODE><<<148, 256>>>(out, 8, err);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<148, 256>>>(out, iters, err); cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int herr; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    printf("%-40s %8.3f ms  %7.1f B/clk/SM  err=%d [%s]\n", name, ms, bytes_per_iter_per_thread * iters * 256 / (ms * 1e-3 * 1.965e9), herr, cudaGetErrorString(e));
    cudaFree(out); cudaFree(err);
}
int main() {
    run<6>("correctness (st then ld, 8 warps)", 0);
    run<0>("tcgen05.ld x16 only", 8 * 64);
    run<1>("tcgen05.st x16 only", 8 * 64);
    run<2>("ld + 8 DFMA + st per 16 cols", 8 * 128);
    run<5>("16 LDS.128 per pass (reference)", 8 * 32);
    run<3>("ld+st + 16 LDS.128 per pass", 8 * 128);
    run<4>("ld+st + 64 DFMA per pass", 8 * 128);
    return 0;
}
