// fft_device.cuh -- FP64 negacyclic ("half-complex") transform for sm_100a, device side.
//
// Replaces the reference's AVX2 kernels cb/spqlios/spqlios-ifft-fma.s:63-263 (coefficients -> spectrum,
// "ifft"/"reverse") and cb/spqlios/spqlios-fft-fma.s:79-274 (spectrum -> coefficients, "fft"/"direct").
// Same mathematics: fold N reals into M=N/2 complex z_j = c_j + i c_{j+M}, twist by
// w^j (w = e^{i pi/N}), complex DFT of size M.  Different algorithm: radix-8 Cooley-Tukey passes with
// 8 points per thread in registers and shared-memory exchanges between passes.
//
//   forward  (DIF):  registers hold x_{t + (M/8) r}  ->  registers hold spectrum slots 8t+s
//   backward (DIT):  the exact mirror; untwist folded into the last step.
//
// The spectral order is engine-private (digit-reversed); bk spectra are produced by the same forward
// code, so pointwise products line up slot by slot.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

typedef double2 cplx;

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
// acc += a*b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// multiply by (S i)
template <int S> __device__ __forceinline__ cplx mul_i(cplx a) {
    return S > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// y_s = sum_r x_r W8^{S r s}, W8 = e^{i pi/4}; in place, natural order in and out.
template <int S> __device__ __forceinline__ void bfly8(cplx (&v)[8]) {
    const double c = 0.70710678118654752440;
    cplx a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    cplx a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    cplx a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
    cplx a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
    // b1 *= W8^S ; b2 *= S i ; b3 *= W8^{3S}
    b1 = S > 0 ? make_double2((b1.x - b1.y) * c, (b1.x + b1.y) * c)
               : make_double2((b1.x + b1.y) * c, (b1.y - b1.x) * c);
    b2 = mul_i<S>(b2);
    b3 = S > 0 ? make_double2((-b3.x - b3.y) * c, (b3.x - b3.y) * c)
               : make_double2((b3.y - b3.x) * c, (-b3.x - b3.y) * c);
    cplx c0 = cadd(a0, a2), d0 = csub(a0, a2);
    cplx c1 = cadd(a1, a3), d1 = mul_i<S>(csub(a1, a3));
    v[0] = cadd(c0, c1); v[4] = csub(c0, c1);
    v[2] = cadd(d0, d1); v[6] = csub(d0, d1);
    cplx e0 = cadd(b0, b2), f0 = csub(b0, b2);
    cplx e1 = cadd(b1, b3), f1 = mul_i<S>(csub(b1, b3));
    v[1] = cadd(e0, e1); v[5] = csub(e0, e1);
    v[3] = cadd(f0, f1); v[7] = csub(f0, f1);
}

// ---------------------------------------------------------------------------------------------
// Plan for M complex points, 8 points per thread, T = M/8 threads per transform.
//   M = 512  : passes (L=512,R=8) (64,8) (8,8)
//   M = 1024 : passes (L=1024,R=8) (128,8) (16,8) (2,2)
// Table layout (cplx entries):  twist[M] | tw1[7][M/8] | tw2[7][M/64] | tw3[7][M/512]
//   twist[j] = e^{i pi j / N};  twP[s-1][j'] = e^{2 pi i j' s / L_P}
// ---------------------------------------------------------------------------------------------
template <int LOGM> struct FftPlan {
    static constexpr int M = 1 << LOGM;
    static constexpr int N = 2 * M;
    static constexpr int T = M / 8;                 // threads per transform
    static constexpr int BUF = M + M / 8;           // padded exchange buffer, cplx entries
    static constexpr int TW_TWIST = 0;
    static constexpr int TW1 = M;
    static constexpr int TW2 = TW1 + 7 * (M / 8);
    static constexpr int TW3 = TW2 + 7 * (M / 64);
    static constexpr int TW_TOTAL = TW3 + (LOGM == 10 ? 7 * (M / 512) : 0);
};

__device__ __forceinline__ int padidx(int p) { return p + (p >> 3); }

__device__ __forceinline__ void group_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

// forward: v[r] = twisted x_{t + T r}  ->  v[s] = spectrum slot 8t+s
template <int LOGM>
__device__ __forceinline__ void fft_forward(cplx (&v)[8], cplx* __restrict__ buf, const cplx* __restrict__ tw,
                                            int t, int bar_id) {
    typedef FftPlan<LOGM> P;
    constexpr int T = P::T;
    // pass 1: L = M, stride T, twiddle W_M^{t s}
    bfly8<1>(v);
#pragma unroll
    for (int s = 1; s < 8; s++) v[s] = cmul(v[s], tw[P::TW1 + (s - 1) * T + t]);
    group_sync(bar_id, T);                                   // WAR: previous transform's last loads
#pragma unroll
    for (int s = 0; s < 8; s++) buf[padidx(t + T * s)] = v[s];
    group_sync(bar_id, T);
    // pass 2: L = M/8, L' = M/64
    {
        constexpr int Lp = P::M / 64;
        const int b = t / Lp, j = t % Lp, base = b * (P::M / 8) + j;
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = buf[padidx(base + Lp * r)];
        bfly8<1>(v);
#pragma unroll
        for (int s = 1; s < 8; s++) v[s] = cmul(v[s], tw[P::TW2 + (s - 1) * Lp + j]);
#pragma unroll
        for (int s = 0; s < 8; s++) buf[padidx(base + Lp * s)] = v[s];
        group_sync(bar_id, T);
    }
    if (LOGM == 9) {
        // pass 3: L = 8, contiguous
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = buf[padidx(8 * t + r)];
        bfly8<1>(v);
    } else {
        // pass 3: L = 16, L' = 2
        {
            const int b = t >> 1, j = t & 1, base = b * 16 + j;
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = buf[padidx(base + 2 * r)];
            bfly8<1>(v);
#pragma unroll
            for (int s = 1; s < 8; s++) v[s] = cmul(v[s], tw[P::TW3 + (s - 1) * 2 + j]);
#pragma unroll
            for (int s = 0; s < 8; s++) buf[padidx(base + 2 * s)] = v[s];
            group_sync(bar_id, T);
        }
        // pass 4: L = 2, four radix-2 butterflies on 8 contiguous points
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = buf[padidx(8 * t + r)];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cplx a = v[2 * i], b = v[2 * i + 1];
            v[2 * i] = cadd(a, b); v[2 * i + 1] = csub(a, b);
        }
    }
}

// backward: v[s] = spectrum slot 8t+s  ->  v[r] = M * x_{t + T r} * (untwisted)   (no 1/M scaling)
template <int LOGM>
__device__ __forceinline__ void fft_backward(cplx (&v)[8], cplx* __restrict__ buf, const cplx* __restrict__ tw,
                                             int t, int bar_id) {
    typedef FftPlan<LOGM> P;
    constexpr int T = P::T;
    if (LOGM == 9) {
        bfly8<-1>(v);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cplx a = v[2 * i], b = v[2 * i + 1];
            v[2 * i] = cadd(a, b); v[2 * i + 1] = csub(a, b);
        }
    }
    group_sync(bar_id, T);                                   // WAR with the previous transform
#pragma unroll
    for (int r = 0; r < 8; r++) buf[padidx(8 * t + r)] = v[r];
    group_sync(bar_id, T);
    if (LOGM == 10) {
        const int b = t >> 1, j = t & 1, base = b * 16 + j;
#pragma unroll
        for (int s = 0; s < 8; s++) v[s] = buf[padidx(base + 2 * s)];
#pragma unroll
        for (int s = 1; s < 8; s++) v[s] = cmulc(v[s], tw[P::TW3 + (s - 1) * 2 + j]);
        bfly8<-1>(v);
#pragma unroll
        for (int r = 0; r < 8; r++) buf[padidx(base + 2 * r)] = v[r];
        group_sync(bar_id, T);
    }
    {
        constexpr int Lp = P::M / 64;
        const int b = t / Lp, j = t % Lp, base = b * (P::M / 8) + j;
#pragma unroll
        for (int s = 0; s < 8; s++) v[s] = buf[padidx(base + Lp * s)];
#pragma unroll
        for (int s = 1; s < 8; s++) v[s] = cmulc(v[s], tw[P::TW2 + (s - 1) * Lp + j]);
        bfly8<-1>(v);
#pragma unroll
        for (int r = 0; r < 8; r++) buf[padidx(base + Lp * r)] = v[r];
        group_sync(bar_id, T);
    }
#pragma unroll
    for (int s = 0; s < 8; s++) v[s] = buf[padidx(t + T * s)];
#pragma unroll
    for (int s = 1; s < 8; s++) v[s] = cmulc(v[s], tw[P::TW1 + (s - 1) * T + t]);
    bfly8<-1>(v);
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = cmulc(v[r], tw[P::TW_TWIST + t + T * r]);
}

// ---------------------------------------------------------------------------------------------
// double -> torus, truncation toward zero then wrap.
//   Torus32: int32_t(int64_t(x))                      cb/spqlios/fft_processor_spqlios.cpp:102
//   Torus64: significand shifted by the exponent      cb/spqlios/fft_processor_spqlios.cpp:131-142
// Done on the integer pipe from the raw bits so the FP64 pipe only sees butterflies.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t double_to_u64_trunc_wrap(double x) {
    const uint64_t bits = (uint64_t)__double_as_longlong(x);
    const uint64_t val = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
    const int trans = (int)((bits >> 52) & 0x7FF) - 1075;
    uint64_t val2;
    if (trans > 0) val2 = trans >= 64 ? 0ull : (val << trans);
    else           val2 = -trans >= 64 ? 0ull : (val >> -trans);
    return (bits >> 63) ? (0ull - val2) : val2;
}
__device__ __forceinline__ int32_t double_to_torus32(double x) { return (int32_t)(uint32_t)double_to_u64_trunc_wrap(x); }
__device__ __forceinline__ int64_t double_to_torus64(double x) { return (int64_t)double_to_u64_trunc_wrap(x); }

// (X^a - 1) * P at coefficient j, a in [0, 2N)   (cb/numeric_functions.cpp:304-323)
template <typename T, int N> __device__ __forceinline__ T rot_minus_one(const T* __restrict__ P, int j, int a) {
    const int idx = (j - a) & (2 * N - 1);
    const T r = P[idx & (N - 1)];
    return (T)(((idx & N) ? (T)0 - r : r) - P[j]);
}

}  // namespace tfhe_b200
