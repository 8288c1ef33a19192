// misc_kernels.cu -- small elementwise pieces of the gate layer.
#include "engine.h"

namespace tfhe_b200 {

// boots* linear part [UPSTREAM, SURVEY Appendix C]: out = (0,cconst) + ka*a + kb*b on LWE(n) samples.
__global__ void lwe_lincomb_kernel(int32_t* out, const int32_t* a, const int32_t* b,
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

// CMUX tree helpers (vertical-packing LUT on TRGSW selectors): CMux(C, d1, d0) = C (x) (d1 - d0) + d0
__global__ void pair_combine_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ in, int len, size_t total, int mode) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t u = e / len; const int k = (int)(e % len);
        const uint32_t d0 = (uint32_t)in[(2 * u) * len + k];
        if (mode == 0) out[e] = (int32_t)((uint32_t)in[(2 * u + 1) * len + k] - d0);
        else out[e] = (int32_t)((uint32_t)out[e] + d0);
    }
}
cudaError_t launch_pair_combine(int32_t* out, const int32_t* in, int len, size_t units, int mode, cudaStream_t s) {
    const size_t total = units * (size_t)len;
    if (!total) return cudaSuccess;
    int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    pair_combine_kernel<<<grid, 256, 0, s>>>(out, in, len, total, mode);
    return cudaGetLastError();
}
// mode 0: out[c][i] = trivial TRLWE (0, table[2i+1] - table[2i]) ; mode 1: out[c][i].b += table[2i]     (i < pairs, every sample c)
__global__ void lut_table_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ table, int N, int pairs, size_t total, int mode) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % (2 * N)); const int i = (int)((e / (2 * N)) % pairs);
        if (k < N) { if (mode == 0) out[e] = 0; continue; }
        const uint32_t even = (uint32_t)table[(size_t)(2 * i) * N + (k - N)];
        if (mode == 0) out[e] = (int32_t)((uint32_t)table[(size_t)(2 * i + 1) * N + (k - N)] - even);
        else out[e] = (int32_t)((uint32_t)out[e] + even);
    }
}
cudaError_t launch_lut_table(int32_t* out, const int32_t* table, int N, int pairs, int count, int mode, cudaStream_t s) {
    const size_t total = (size_t)count * pairs * 2 * N;
    if (!total) return cudaSuccess;
    int grid = (int)((total + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    lut_table_kernel<<<grid, 256, 0, s>>>(out, table, N, pairs, total, mode);
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


// ---- diagnostics: measured FP64 and read-bandwidth ceilings for the roofline (bench.py)
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-9;
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;     // never true; keeps the chain alive
}
cudaError_t probe_fp64(double* tflops) {
    double* d; cudaError_t e = cudaMalloc(&d, 8); if (e != cudaSuccess) return e;
    const int grid = 148 * 8, iters = 1 << 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_probe_kernel<<<grid, 256>>>(d, 1024, 1.0);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        fp64_probe_kernel<<<grid, 256>>>(d, iters, 1.0);
        cudaEventRecord(e1); e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8 * (double)iters * 256 * grid / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return e;
}
__global__ void __launch_bounds__(256) read_probe_kernel(const int4* __restrict__ p, size_t n16, int passes, int* sink) {
    int acc = 0;
    for (int ps = 0; ps < passes; ps++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
            int4 v = __ldcg(p + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x7fffffff) *sink = acc;
}
cudaError_t probe_read(size_t bytes, int passes, double* gbs) {
    void* d; int* sink;
    cudaError_t e = cudaMalloc(&d, bytes); if (e != cudaSuccess) return e;
    cudaMalloc(&sink, 4); cudaMemset(d, 1, bytes);
    const size_t n16 = bytes / 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    read_probe_kernel<<<148 * 8, 256>>>((const int4*)d, n16, 1, sink);
    cudaEventRecord(e0);
    read_probe_kernel<<<148 * 8, 256>>>((const int4*)d, n16, passes, sink);
    cudaEventRecord(e1); e = cudaEventSynchronize(e1);
    float ms = 1; cudaEventElapsedTime(&ms, e0, e1);
    *gbs = (double)n16 * 16 * passes / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); cudaFree(sink);
    return e;
}

}  // namespace tfhe_b200
