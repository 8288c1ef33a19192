// tma_probe.cu -- latency / throughput of 1-D cp.async.bulk global->shared on B200 (design probe for bk_pipe.cuh)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../experimental-tfhe_b200/csrc/bk_pipe.cuh"
using namespace tfhe_b200;

// mode 0: one copy in flight at a time (latency); mode 1: DEPTH copies in flight (throughput)
template <int DEPTH> __global__ void probe(const unsigned char* src, size_t src_bytes, int chunk, int iters, int same, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bars[DEPTH];
    if (threadIdx.x == 0) { for (int i = 0; i < DEPTH; i++) mbar_init(&bars[i], 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned char* base = same ? src : src + ((size_t)blockIdx.x * 1048576) % (src_bytes - (size_t)chunk * iters);
    long long t0 = clock64();
    for (int i = 0; i < DEPTH && i < iters; i++) {
        mbar_expect_tx(&bars[i], chunk);
        tma_load_1d(sm + (size_t)i * chunk, base + (size_t)i * chunk, chunk, &bars[i]);
    }
    for (int i = 0; i < iters; i++) {
        const int s = i % DEPTH;
        while (!mbar_try_wait(&bars[s], (i / DEPTH) & 1)) {}
        if (i + DEPTH < iters) {
            mbar_expect_tx(&bars[s], chunk);
            tma_load_1d(sm + (size_t)s * chunk, base + (size_t)(i + DEPTH) * chunk, chunk, &bars[s]);
        }
    }
    out[blockIdx.x] = clock64() - t0;
}
template <int DEPTH> void run(const unsigned char* d, size_t bytes, int chunk, int grid, int same) {
    long long* out; cudaMalloc(&out, grid * 8);
    const int iters = 512;
    cudaFuncSetAttribute(probe<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * chunk);
    probe<DEPTH><<<grid, 32, DEPTH * chunk>>>(d, bytes, chunk, iters, same, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
    printf("chunk %5d B depth %d grid %3d %s: %8.0f cycles/copy (max over CTAs)  -> %.1f B/clk/SM  [%s]\n", chunk, DEPTH, grid,
           same ? "same addr " : "diff addrs", (double)mx / iters, (double)chunk * iters / mx, cudaGetErrorString(e));
    cudaFree(out);
}
int main() {
    size_t bytes = 256u << 20; unsigned char* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    for (int same = 1; same >= 0; same--)
        for (int grid : {1, 148}) {
            run<1>(d, bytes, 8192, grid, same);
            run<2>(d, bytes, 8192, grid, same);
            run<6>(d, bytes, 8192, grid, same);
            run<3>(d, bytes, 16384, grid, same);
        }
    return 0;
}
