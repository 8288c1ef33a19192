// hp_kernels.cu -- 128-bit fixed-point anticyclic FFT (hp/code.cpp) on sm_100a.
//
// Real96 (hp/code.cpp:25-75) = value * 2^64 in a wrapping 128-bit integer.  Every operation is exact integer
// arithmetic (adds wrap, products truncate to 64 fractional bits), so the device result is BIT-IDENTICAL to
// the reference for identical twiddle tables, whatever the butterfly schedule.
//
// One CTA per polynomial.  The N/2 complex points live in shared memory as four 64-bit planes
// (re_lo, re_hi, im_lo, im_hi: consecutive threads hit consecutive banks).  Each pass keeps a radix-4 group
// (two radix-2 levels of the reference's loop nest, hp/code.cpp:414-436 / :473-494) in registers, so the
// ten/eleven levels cost five/six shared-memory round trips.
#include "engine.h"

namespace tfhe_b200 {

typedef unsigned __int128 u128;
typedef __int128 i128;
struct c96 { u128 re, im; };

// intmul_best (hp/code.cpp:148-169): a * b truncated to 64 fractional bits, b a twiddle in [-1,1).
// Equals intmul_ref (:79-95) whenever a's integer part fits int32, which the reference asserts.
__device__ __forceinline__ u128 real96_mul(u128 a, u128 b) {
    const uint64_t alo = (uint64_t)a, blo = (uint64_t)b;
    const int64_t ahi = (int64_t)(uint64_t)(a >> 64);
    u128 w = (u128)__umul64hi(alo, blo);
    w += (u128)((i128)ahi) * (u128)blo;
    if ((int64_t)(uint64_t)(b >> 64) < 0) w -= a;
    return w;
}
// libstdc++ complex<T>::operator*= order: data on the left, twiddle on the right
__device__ __forceinline__ c96 cmul96(c96 a, c96 b) {
    c96 r;
    r.re = real96_mul(a.re, b.re) - real96_mul(a.im, b.im);
    r.im = real96_mul(a.re, b.im) + real96_mul(a.im, b.re);
    return r;
}
__device__ __forceinline__ c96 ld_tw(const uint64_t* __restrict__ tab, int i) {
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(tab) + 2 * i);
    const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2*>(tab) + 2 * i + 1);
    c96 r; r.re = ((u128)a.y << 64) | a.x; r.im = ((u128)b.y << 64) | b.x;
    return r;
}

struct Planes {
    uint64_t *rl, *rh, *il, *ih;
    __device__ __forceinline__ c96 get(int i) const {
        c96 r; r.re = ((u128)rh[i] << 64) | rl[i]; r.im = ((u128)ih[i] << 64) | il[i];
        return r;
    }
    __device__ __forceinline__ void put(int i, c96 v) const {
        rl[i] = (uint64_t)v.re; rh[i] = (uint64_t)(v.re >> 64); il[i] = (uint64_t)v.im; ih[i] = (uint64_t)(v.im >> 64);
    }
};
__device__ __forceinline__ c96 add96(c96 a, c96 b) { c96 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
__device__ __forceinline__ c96 sub96(c96 a, c96 b) { c96 r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }

constexpr int HP_THREADS = 256;

// iFFT (hp/code.cpp:391-443): P -> P(omega).  n = 2N, ns4 = N/2 points.
__global__ void __launch_bounds__(HP_THREADS) hp_ifft_kernel(tfhe_b200_cplx96* __restrict__ out, const int64_t* __restrict__ in,
                                                             const uint64_t* __restrict__ powomega, int N) {
    extern __shared__ __align__(16) uint64_t hp_smem[];
    const int ns4 = N / 2, n = 2 * N;
    Planes P{hp_smem, hp_smem + ns4, hp_smem + 2 * ns4, hp_smem + 3 * ns4};
    const int64_t* src = in + (size_t)blockIdx.x * N;
    // out[j] = (in[j] + i in[j+ns4]) * omega^j   (:407-408)
    for (int j = threadIdx.x; j < ns4; j += HP_THREADS) {
        c96 z; z.re = (u128)(i128)src[j]; z.im = (u128)(i128)src[j + ns4];
        P.put(j, cmul96(z, ld_tw(powomega, j)));
    }
    __syncthreads();
    // DIF levels nn = ns4 .. 2 (:414-436), two levels per pass
    int nn = ns4;
    while (nn >= 4) {
        const int h = nn >> 1, q = nn >> 2;            // halfnn of this level and of the next
        for (int g = threadIdx.x; g < ns4 / 4; g += HP_THREADS) {
            const int blk = (g / q) * nn, off = g % q;
            const int i0 = blk + off, i1 = i0 + q, i2 = i0 + h, i3 = i2 + q;
            c96 x0 = P.get(i0), x1 = P.get(i1), x2 = P.get(i2), x3 = P.get(i3);
            // level nn: pairs (i0,i2) with off, (i1,i3) with off+q
            const int m1 = 2 * (ns4 / h);
            c96 a0 = add96(x0, x2), a2 = cmul96(sub96(x0, x2), ld_tw(powomega, (m1 * off) % n));
            c96 a1 = add96(x1, x3), a3 = cmul96(sub96(x1, x3), ld_tw(powomega, (m1 * (off + q)) % n));
            // level nn/2: pairs (i0,i1) and (i2,i3), both with off
            const int m2 = 2 * (ns4 / q);
            const c96 w = ld_tw(powomega, (m2 * off) % n);
            P.put(i0, add96(a0, a1)); P.put(i1, cmul96(sub96(a0, a1), w));
            P.put(i2, add96(a2, a3)); P.put(i3, cmul96(sub96(a2, a3), w));
        }
        __syncthreads();
        nn >>= 2;
    }
    if (nn == 2) {     // odd number of levels: last level alone (halfnn = 1, twiddle index 0)
        const c96 w = ld_tw(powomega, 0);
        for (int g = threadIdx.x; g < ns4 / 2; g += HP_THREADS) {
            c96 t1 = P.get(2 * g), t2 = P.get(2 * g + 1);
            P.put(2 * g, add96(t1, t2)); P.put(2 * g + 1, cmul96(sub96(t1, t2), w));
        }
        __syncthreads();
    }
    tfhe_b200_cplx96* dst = out + (size_t)blockIdx.x * ns4;
    for (int j = threadIdx.x; j < ns4; j += HP_THREADS) {
        tfhe_b200_cplx96 o; o.re_lo = P.rl[j]; o.re_hi = P.rh[j]; o.im_lo = P.il[j]; o.im_hi = P.ih[j];
        dst[j] = o;
    }
}

// FFT (hp/code.cpp:446-512): P(omega) -> P
__global__ void __launch_bounds__(HP_THREADS) hp_fft_kernel(int64_t* __restrict__ out, const tfhe_b200_cplx96* __restrict__ in,
                                                            const uint64_t* __restrict__ powombar, int N) {
    extern __shared__ __align__(16) uint64_t hp_smem[];
    const int ns4 = N / 2, n = 2 * N;
    Planes P{hp_smem, hp_smem + ns4, hp_smem + 2 * ns4, hp_smem + 3 * ns4};
    const tfhe_b200_cplx96* src = in + (size_t)blockIdx.x * ns4;
    for (int j = threadIdx.x; j < ns4; j += HP_THREADS) {
        const tfhe_b200_cplx96 v = src[j];
        P.rl[j] = v.re_lo; P.rh[j] = v.re_hi; P.il[j] = v.im_lo; P.ih[j] = v.im_hi;
    }
    __syncthreads();
    int log_ns4 = 0; while ((1 << log_ns4) < ns4) log_ns4++;
    int nn = 2;
    if (log_ns4 & 1) {   // odd number of levels: first level alone (twiddle index 0)
        const c96 w = ld_tw(powombar, 0);
        for (int g = threadIdx.x; g < ns4 / 2; g += HP_THREADS) {
            c96 t1 = P.get(2 * g), t2 = cmul96(P.get(2 * g + 1), w);
            P.put(2 * g, add96(t1, t2)); P.put(2 * g + 1, sub96(t1, t2));
        }
        __syncthreads();
        nn = 4;
    }
    // DIT levels (:473-494), two per pass: level nn (halfnn = q) then level 2nn (halfnn = h)
    while (nn <= ns4 / 2) {
        const int q = nn >> 1, h = nn, nn2 = nn << 1;
        for (int g = threadIdx.x; g < ns4 / 4; g += HP_THREADS) {
            const int blk = (g / q) * nn2, off = g % q;
            const int i0 = blk + off, i1 = i0 + q, i2 = i0 + h, i3 = i2 + q;
            const int m1 = 2 * (ns4 / q);
            const c96 w = ld_tw(powombar, (m1 * off) % n);
            c96 x0 = P.get(i0), x1 = cmul96(P.get(i1), w), x2 = P.get(i2), x3 = cmul96(P.get(i3), w);
            c96 a0 = add96(x0, x1), a1 = sub96(x0, x1), a2 = add96(x2, x3), a3 = sub96(x2, x3);
            const int m2 = 2 * (ns4 / h);
            c96 b2 = cmul96(a2, ld_tw(powombar, (m2 * off) % n));
            c96 b3 = cmul96(a3, ld_tw(powombar, (m2 * (off + q)) % n));
            P.put(i0, add96(a0, b2)); P.put(i2, sub96(a0, b2));
            P.put(i1, add96(a1, b3)); P.put(i3, sub96(a1, b3));
        }
        __syncthreads();
        nn <<= 2;
    }
    // in[j] *= ombar^j ; out = v >> log2(N/2), low 64 bits (:500-504)
    int64_t* dst = out + (size_t)blockIdx.x * N;
    for (int j = threadIdx.x; j < ns4; j += HP_THREADS) {
        const c96 v = cmul96(P.get(j), ld_tw(powombar, j));
        dst[j] = (int64_t)(uint64_t)(v.re >> log_ns4);
        dst[j + ns4] = (int64_t)(uint64_t)(v.im >> log_ns4);
    }
}

cudaError_t hp_init() {
    cudaError_t e = cudaFuncSetAttribute(hp_ifft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 8);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(hp_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 8);
}
cudaError_t launch_hp_ifft(tfhe_b200_cplx96* out, const int64_t* in, const uint64_t* powomega, int N, int count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    hp_ifft_kernel<<<count, HP_THREADS, (size_t)4 * (N / 2) * 8, s>>>(out, in, powomega, N);
    return cudaGetLastError();
}
cudaError_t launch_hp_fft(int64_t* out, const tfhe_b200_cplx96* in, const uint64_t* powombar, int N, int count, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    hp_fft_kernel<<<count, HP_THREADS, (size_t)4 * (N / 2) * 8, s>>>(out, in, powombar, N);
    return cudaGetLastError();
}

// Arithmetic roofline of the 128-bit transforms: the rate of real96 products (the same real96_mul, eight independent chains per
// thread, operands in registers) that this GPU sustains -- measured in the bench run, next to the transforms it bounds.
__global__ void __launch_bounds__(256) real96_probe_kernel(uint64_t* out, int iters, uint64_t seed) {
    u128 a[8], b[2];
    for (int i = 0; i < 8; i++) a[i] = ((u128)(seed + threadIdx.x + i) << 64) | (0x9E3779B97F4A7C15ull * (threadIdx.x + i + 1));
    b[0] = (u128)(0xB5297A4D3F84D5B5ull ^ seed);                       // twiddles in [0,1): high word 0
    b[1] = ~(u128)0 << 64 | (0x68E31DA4B5297A4Dull + seed);            // and in [-1,0): high word -1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = real96_mul(a[i], b[i & 1]) + (u128)it;
    }
    uint64_t s = 0;
    for (int i = 0; i < 8; i++) s ^= (uint64_t)a[i] ^ (uint64_t)(a[i] >> 64);
    if (s == 0x1234567) out[0] = s;
}
cudaError_t probe_real96(double* gprod_per_s) {
    uint64_t* d; cudaError_t e = cudaMalloc(&d, 8); if (e != cudaSuccess) return e;
    const int grid = 148 * 8, iters = 1 << 13;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    real96_probe_kernel<<<grid, 256>>>(d, 256, 1);
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        real96_probe_kernel<<<grid, 256>>>(d, iters, 7 + rep);
        cudaEventRecord(e1); e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double g = 8.0 * (double)iters * 256 * grid / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *gprod_per_s = best;
    return e;
}

}  // namespace tfhe_b200
