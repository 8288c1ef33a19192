// misc_kernels.cu -- small elementwise pieces of the gate layer.
#include "engine.h"

namespace tfhe_b200 {

// boots* linear part [UPSTREAM, SURVEY Appendix C]: out = (0,cconst) + ka*a + kb*b on LWE(n) samples.
__global__ void lwe_lincomb_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                   int ka, int kb, int32_t cconst, int n, size_t total) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)ka * (uint32_t)a[e];
        if (b) x += (uint32_t)kb * (uint32_t)b[e];
        if ((int)(e % (size_t)(n + 1)) == n) x += (uint32_t)cconst;
        out[e] = (int32_t)x;
    }
}
cudaError_t launch_lwe_lincomb(int32_t* out, const int32_t* a, const int32_t* b, int ka, int kb, int32_t cconst,
                               int n, int count, cudaStream_t s) {
    const size_t total = (size_t)count * (n + 1);
    if (!total) return cudaSuccess;
    int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    lwe_lincomb_kernel<<<grid, 256, 0, s>>>(out, a, b, ka, kb, cconst, n, total);
    return cudaGetLastError();
}

// modSwitchFromTorus32 (cb/numeric_functions.cpp:54-60) / preModSwitch (cb/poc_CircuitBootstrapping.cpp:472-484), Msize = 2^k
__global__ void modswitch_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ in, int log2Msize, size_t total) {
    const uint64_t half = 1ull << (63 - log2Msize);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint64_t phase64 = ((uint64_t)(uint32_t)in[e] << 32) + half;
        out[e] = (int32_t)(phase64 >> (64 - log2Msize));
    }
}
cudaError_t launch_modswitch(int32_t* out, const int32_t* in, int log2Msize, size_t total, cudaStream_t s) {
    if (!total) return cudaSuccess;
    int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    modswitch_kernel<<<grid, 256, 0, s>>>(out, in, log2Msize, total);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
