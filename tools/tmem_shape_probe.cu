// tmem_shape_probe.cu -- which (TMEM lane, column) does each thread receive from the cross-lane tcgen05.ld shapes?  (design probe)
// Every thread of a warp stores its own lane's 16 columns with 32x32b (value = lane*256 + column); then the warp loads with
// 16x64b / 16x128b / 16x256b at lane offsets 0 and 16 and prints what each thread got.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    __shared__ uint32_t tbase_s;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase_s;
    if (w == 0) {
        uint32_t r[16];
        for (int i = 0; i < 16; i++) r[i] = lane * 256 + i;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                     ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                       "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(tb) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int half = 0; half < 2; half++) {
            const uint32_t ta = tb + ((uint32_t)(16 * half) << 16);
            uint32_t a0, a1, a2, a3;
            asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(a0) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[(0 * 2 + half) * 128 + lane * 4 + 0] = a0;
            asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[(1 * 2 + half) * 128 + lane * 4 + 0] = a0; out[(1 * 2 + half) * 128 + lane * 4 + 1] = a1;
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            out[(2 * 2 + half) * 128 + lane * 4 + 0] = a0; out[(2 * 2 + half) * 128 + lane * 4 + 1] = a1;
            out[(2 * 2 + half) * 128 + lane * 4 + 2] = a2; out[(2 * 2 + half) * 128 + lane * 4 + 3] = a3;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}
int main() {
    uint32_t* d; cudaMalloc(&d, 6 * 128 * 4); cudaMemset(d, 0xff, 6 * 128 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t h[6 * 128]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[3] = {"16x64b.x1 (1 reg)", "16x128b.x1 (2 regs)", "16x256b.x1 (4 regs)"};
    const int nreg[3] = {1, 2, 4};
    for (int s = 0; s < 3; s++)
        for (int half = 0; half < 2; half++) {
            printf("== %s, lane offset %d: thread -> (lane,col) per register\n", names[s], 16 * half);
            for (int t = 0; t < 32; t++) {
                printf("  t%2d:", t);
                for (int k = 0; k < nreg[s]; k++) { uint32_t v = h[(s * 2 + half) * 128 + t * 4 + k]; printf(" (%2u,%2u)", v >> 8, v & 255); }
                if (t % 4 == 3) printf("\n");
            }
        }
    return 0;
}
