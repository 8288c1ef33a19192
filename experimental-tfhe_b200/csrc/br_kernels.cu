// br_kernels.cu -- fused blind rotation for sm_100a.
//
// One thread GROUP (T = N/16 threads: 64 for N=1024, 128 for N=2048) owns one TLWE accumulator for all n
// CMUX steps; the accumulator lives in shared memory for the whole kernel and is touched in HBM only
// at the start (input LWE sample) and at the end (extracted LWE sample).  Per step the group fuses
//   (X^a - 1) * ACC            tLweMulByXaiMinusOne         cb/tlwe_functions.cpp:209-213
//   gadget decomposition       tGswTorus32PolynomialDecompH cb/tgsw_functions.cpp:224-337
//                              tGswTorus64PolynomialDecompH cb/poc_CircuitBootstrapping.cpp:492-515
//   2l forward transforms      IntPolynomial_ifft           cb/tgsw_functions.cpp:438
//   2l x 2 spectral MACs       tLweFFTAddMulRTo             cb/tlwe_functions.cpp:318-325
//   2 backward transforms      tLweFromFFTConvert           cb/tlwe_functions.cpp:299-305
//   ACC += result              tLweAddTo                    cb/tlwe_functions.cpp:163-170
// i.e. tfhe_MuxRotate_FFT (cb/lwe_functions.cpp:328-333) inside the loop of tfhe_blindRotate_FFT
// (:337-361), plus modulus switch, test-vector rotation and sample extraction on either side
// (tfhe_bootstrap_woKS_FFT :399-430; circuitBootstrapWoKS cb/poc_CircuitBootstrapping.cpp:530-659).
//
// The reference's `if (barai==0) continue` (:350) is kept (group-uniform branch).
#include "engine.h"
#include "fft_device.cuh"

namespace tfhe_b200 {

template <typename Torus> struct TorusTraits;
template <> struct TorusTraits<int32_t> { typedef uint32_t U; static constexpr int W = 32; };
template <> struct TorusTraits<int64_t> { typedef uint64_t U; static constexpr int W = 64; };

// decomposition offset: Torus32 library form has no rounding bit (cb/tgsw_functions.cpp:30-36);
// the Torus64 PoC form carries one (cb/poc_CircuitBootstrapping.cpp:349-350).  SURVEY A.5.
__device__ __forceinline__ uint32_t decomp_offset(uint32_t, int l, int Bgbit) {
    uint32_t t = 0;
    for (int i = 0; i < l; i++) t += 1u << (32 - (i + 1) * Bgbit);
    return t * (uint32_t)((1 << Bgbit) / 2);
}
__device__ __forceinline__ uint64_t decomp_offset(uint64_t, int l, int Bgbit) {
    uint64_t t = 0;
    for (int i = 0; i <= l; i++) t |= 1ull << (63 - i * Bgbit);
    return t;
}
__device__ __forceinline__ int32_t to_torus(double x, int32_t) { return double_to_torus32(x); }
__device__ __forceinline__ int64_t to_torus(double x, int64_t) { return double_to_torus64(x); }

// modSwitchFromTorus32 (cb/numeric_functions.cpp:54-60) for Msize = 2^log2M
__device__ __forceinline__ int modswitch32(int32_t x, int log2Msize) {
    const uint64_t half = 1ull << (63 - log2Msize);           // interv/2, interv = 2^(64-log2Msize)
    const uint64_t phase64 = ((uint64_t)(uint32_t)x << 32) + half;
    return (int)(phase64 >> (64 - log2Msize));
}

// One CMUX: ACC <- ACC + BK_i (x) ((X^a - 1) ACC).   acc: shared [2][N].  bk: [2l][2][M] spectra (scaled 2/N).
template <int LOGM, typename Torus>
__device__ __forceinline__ void cmux_step(Torus* __restrict__ acc, const int a, const cplx* __restrict__ bk,
                                          const int l, const int Bgbit, cplx* __restrict__ buf,
                                          const cplx* __restrict__ tw, const int t, const int bar_id) {
    typedef FftPlan<LOGM> P;
    typedef typename TorusTraits<Torus>::U U;
    constexpr int M = P::M, N = P::N, T = P::T, W = TorusTraits<Torus>::W;
    const U offset = decomp_offset((U)0, l, Bgbit);
    const uint32_t mask = (1u << Bgbit) - 1u;
    const int half = 1 << (Bgbit - 1);

    cplx R0[8], R1[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { R0[e] = make_double2(0.0, 0.0); R1[e] = make_double2(0.0, 0.0); }

#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        U ure[8], uim[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int j = t + T * r;
            ure[r] = (U)rot_minus_one<Torus, N>(acc + q * N, j, a) + offset;
            uim[r] = (U)rot_minus_one<Torus, N>(acc + q * N, j + M, a) + offset;
        }
#pragma unroll 1
        for (int lev = 0; lev < l; lev++) {
            const int sh = W - (lev + 1) * Bgbit;
            cplx v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const double dre = (double)((int)((uint32_t)(ure[r] >> sh) & mask) - half);
                const double dim = (double)((int)((uint32_t)(uim[r] >> sh) & mask) - half);
                v[r] = cmul(make_double2(dre, dim), tw[P::TW_TWIST + t + T * r]);
            }
            fft_forward<LOGM>(v, buf, tw, t, bar_id);
            const cplx* __restrict__ bk0 = bk + (size_t)((q * l + lev) * 2) * M + t;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cplx b0 = __ldg(bk0 + e * T);
                const cplx b1 = __ldg(bk0 + M + e * T);
                cfma(R0[e], v[e], b0);
                cfma(R1[e], v[e], b1);
            }
        }
    }
    fft_backward<LOGM>(R0, buf, tw, t, bar_id);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = t + T * r;
        acc[j] = (Torus)((U)acc[j] + (U)to_torus(R0[r].x, (Torus)0));
        acc[j + M] = (Torus)((U)acc[j + M] + (U)to_torus(R0[r].y, (Torus)0));
    }
    fft_backward<LOGM>(R1, buf, tw, t, bar_id);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = t + T * r;
        acc[N + j] = (Torus)((U)acc[N + j] + (U)to_torus(R1[r].x, (Torus)0));
        acc[N + j + M] = (Torus)((U)acc[N + j + M] + (U)to_torus(R1[r].y, (Torus)0));
    }
    group_sync(bar_id, T);      // accumulator writes visible before the next step's rotated reads
}

template <int LOGM, typename Torus, int GROUPS> struct BRSmem {
    typedef FftPlan<LOGM> P;
    static constexpr int NPAD = 1024;   // room for bara (n <= 1024)
    static constexpr size_t TW_BYTES = sizeof(cplx) * ((P::TW_TOTAL + 1) & ~1);
    static constexpr size_t GROUP_BYTES = sizeof(cplx) * P::BUF + sizeof(Torus) * 2 * P::N + sizeof(int32_t) * NPAD;
    static constexpr size_t TOTAL = TW_BYTES + GROUPS * GROUP_BYTES;
};

template <int LOGM, typename Torus, int GROUPS, int MINB>
__global__ void __launch_bounds__(GROUPS * FftPlan<LOGM>::T, MINB) blind_rotate_kernel(const BRArgs A) {
    typedef FftPlan<LOGM> P;
    typedef typename TorusTraits<Torus>::U U;
    typedef BRSmem<LOGM, Torus, GROUPS> S;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += GROUPS * T) tw[i] = A.tw[i];
    __syncthreads();

    const int g = threadIdx.x / T, t = threadIdx.x % T;
    const int bar_id = 1 + g;
    unsigned char* gbase = smem_raw + S::TW_BYTES + (size_t)g * S::GROUP_BYTES;
    cplx* buf = reinterpret_cast<cplx*>(gbase);
    Torus* acc = reinterpret_cast<Torus*>(gbase + sizeof(cplx) * P::BUF);
    int32_t* bara = reinterpret_cast<int32_t*>(gbase + sizeof(cplx) * P::BUF + sizeof(Torus) * 2 * N);

    // unit = (sample, test-vector index); n_mu > 1 only on the circuit-bootstrap path
    const int n_mu = A.n_mu > 0 ? A.n_mu : 1;
    const long unit = (long)blockIdx.x * GROUPS + g;
    if (unit >= (long)A.count * n_mu) return;          // whole group leaves; named barriers are per group
    const int ct = (int)(unit / n_mu), w = (int)(unit % n_mu);
    const int n = A.n;

    // ---- rotation amounts and the initial accumulator
    int barb = 0;
    Torus mu = (Torus)A.mu;
    if (A.mode == BR_LWE) {
        if (sizeof(Torus) == 4) {
            // tfhe_bootstrap_woKS_FFT :416-419 on x = (0,cconst) + ka*xa + kb*xb  (boots* linear part)
            const int32_t* xa = A.xa + (size_t)ct * (n + 1);
            const int32_t* xb = A.xb ? A.xb + (size_t)ct * (n + 1) : nullptr;
            for (int i = t; i <= n; i += T) {
                uint32_t x = (uint32_t)A.ka * (uint32_t)xa[i];
                if (xb) x += (uint32_t)A.kb * (uint32_t)xb[i];
                if (i == n) x += (uint32_t)A.cconst;
                bara[i] = modswitch32((int32_t)x, LOGM + 2);
            }
        } else {
            // circuitBootstrapWoKS: abar[n0+1] already mod-switched (cb/poc_CircuitBootstrapping.cpp:542,581)
            const int32_t* ab = A.bara + (size_t)ct * (n + 1);
            for (int i = t; i <= n; i += T) bara[i] = ab[i];
            if (A.mu_bgbit > 0) mu = (Torus)(1ull << (64 - (w + 1) * A.mu_bgbit));   // mu_w, poc:846
        }
        group_sync(bar_id, T);
        barb = bara[n];
    } else {
        const int32_t* ab = A.bara + (size_t)ct * n;
        for (int i = t; i < n; i += T) bara[i] = ab[i];
        if (A.mode == BR_TESTVEC) barb = A.barb[ct];
        group_sync(bar_id, T);
    }

    if (A.mode == BR_ACCUM) {
        const Torus* src = reinterpret_cast<const Torus*>(A.accum) + (size_t)ct * 2 * N;
        for (int j = t; j < 2 * N; j += T) acc[j] = src[j];
    } else {
        // testvectbis = X^(2N - barb) * v  (cb/lwe_functions.cpp:385; defect D3 of the PoC corrected the same way)
        const int rot = (2 * N - barb) & (2 * N - 1);
        const Torus* v = reinterpret_cast<const Torus*>(A.v);
        for (int j = t; j < N; j += T) {
            const int idx = (j - rot) & (2 * N - 1);
            const int k0 = idx & (N - 1);
            Torus val;
            if (A.mode == BR_TESTVEC) val = v[k0];
            else if (sizeof(Torus) == 4) val = mu;                              // [mu,...,mu]  :422
            else val = (k0 < M) ? (Torus)(0 - (U)(mu / 2)) : (Torus)(mu / 2);   // -mu/2 | +mu/2  poc:552-553
            acc[j] = 0;                                                          // tLweNoiselessTrivial
            acc[N + j] = (idx & N) ? (Torus)(0 - (U)val) : val;
        }
    }
    group_sync(bar_id, T);

    // ---- n CMUX steps (tfhe_blindRotate_FFT :348-354)
    const size_t bk_stride = (size_t)2 * A.l * 2 * M;
    for (int i = 0; i < n; i++) {
        const int a = bara[i];
        if (a == 0) continue;
        cmux_step<LOGM, Torus>(acc, a, A.bkfft + (size_t)i * bk_stride, A.l, A.Bgbit, buf, tw, t, bar_id);
    }

    // ---- epilogue
    if (A.mode == BR_ACCUM) {
        Torus* dst = reinterpret_cast<Torus*>(A.accum) + (size_t)ct * 2 * N;
        for (int j = t; j < 2 * N; j += T) dst[j] = acc[j];
    } else {
        // tLweExtractLweSampleIndex(index 0) cb/tlwe_functions.cpp:351-363 ; PoC adds mu/2 to b (:648)
        Torus* out = reinterpret_cast<Torus*>(A.out) + (size_t)unit * A.out_stride;
        for (int j = t; j < N; j += T) out[j] = (j == 0) ? acc[0] : (Torus)(0 - (U)acc[N - j]);
        if (t == 0) {
            U b = (U)acc[N];
            if (sizeof(Torus) == 8 && A.mode == BR_LWE) b += (U)(mu / 2);
            out[N] = (Torus)b;
        }
    }
}

static bool g_inited = false;
constexpr int G32 = 4, G64 = 2;

cudaError_t blind_rotate_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<9, int32_t, G32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)BRSmem<9, int32_t, G32>::TOTAL);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(blind_rotate_kernel<10, int64_t, G64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)BRSmem<10, int64_t, G64>::TOTAL);
    if (e != cudaSuccess) return e;
    g_inited = true;
    return cudaSuccess;
}

cudaError_t launch_blind_rotate32(const BRArgs& a, cudaStream_t s) {
    if (!g_inited) { cudaError_t e = blind_rotate_init(); if (e != cudaSuccess) return e; }
    if (a.count <= 0) return cudaSuccess;
    const int grid = (a.count + G32 - 1) / G32;
    blind_rotate_kernel<9, int32_t, G32, 1><<<grid, G32 * FftPlan<9>::T, BRSmem<9, int32_t, G32>::TOTAL, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_blind_rotate64(const BRArgs& a, cudaStream_t s) {
    if (!g_inited) { cudaError_t e = blind_rotate_init(); if (e != cudaSuccess) return e; }
    if (a.count <= 0) return cudaSuccess;
    const long units = (long)a.count * (a.n_mu > 0 ? a.n_mu : 1);
    const int grid = (int)((units + G64 - 1) / G64);
    blind_rotate_kernel<10, int64_t, G64, 1><<<grid, G64 * FftPlan<10>::T, BRSmem<10, int64_t, G64>::TOTAL, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Standalone transforms (one group of T threads per polynomial, 4 polynomials per CTA)
// ---------------------------------------------------------------------------------------------
template <int LOGM, typename Torus>
__global__ void __launch_bounds__(4 * FftPlan<LOGM>::T) poly_to_spectrum_kernel(cplx* __restrict__ out, const Torus* __restrict__ in,
                                                                                const cplx* __restrict__ twg, int count, double scale) {
    typedef FftPlan<LOGM> P;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += 4 * T) tw[i] = twg[i];
    __syncthreads();
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    cplx* buf = tw + ((P::TW_TOTAL + 1) & ~1) + g * P::BUF;
    const long poly = (long)blockIdx.x * 4 + g;
    if (poly >= count) return;
    const Torus* src = in + (size_t)poly * N;
    cplx v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = t + T * r;
        // execute_reverse_int: exact int->double (:33-46); execute_reverse_torus64: C cast, round to 53 bits (:166-170)
        v[r] = cmul(make_double2((double)src[j], (double)src[j + M]), tw[P::TW_TWIST + j]);
    }
    fft_forward<LOGM>(v, buf, tw, t, 1 + g);
    cplx* dst = out + (size_t)poly * M + t;
#pragma unroll
    for (int e = 0; e < 8; e++) dst[e * T] = make_double2(v[e].x * scale, v[e].y * scale);
}

template <int LOGM, typename Torus>
__global__ void __launch_bounds__(4 * FftPlan<LOGM>::T) spectrum_to_torus_kernel(Torus* __restrict__ out, const cplx* __restrict__ in,
                                                                                 const cplx* __restrict__ twg, int count, double scale) {
    typedef FftPlan<LOGM> P;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += 4 * T) tw[i] = twg[i];
    __syncthreads();
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    cplx* buf = tw + ((P::TW_TOTAL + 1) & ~1) + g * P::BUF;
    const long poly = (long)blockIdx.x * 4 + g;
    if (poly >= count) return;
    const cplx* src = in + (size_t)poly * M + t;
    cplx v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { cplx x = src[e * T]; v[e] = make_double2(x.x * scale, x.y * scale); }   // 2/N pre-scale (:78-100)
    fft_backward<LOGM>(v, buf, tw, t, 1 + g);
    Torus* dst = out + (size_t)poly * N;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int j = t + T * r;
        dst[j] = to_torus(v[r].x, (Torus)0);
        dst[j + M] = to_torus(v[r].y, (Torus)0);
    }
}

template <int LOGM> static size_t tr_smem() { return sizeof(cplx) * (((FftPlan<LOGM>::TW_TOTAL + 1) & ~1) + 4 * FftPlan<LOGM>::BUF); }

template <int LOGM, typename Torus>
static cudaError_t launch_p2s(cplx* out, const Torus* in, const cplx* tw, int count, double scale, cudaStream_t s) {
    auto kern = poly_to_spectrum_kernel<LOGM, Torus>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem<LOGM>());
    if (e != cudaSuccess) return e;
    if (count <= 0) return cudaSuccess;
    kern<<<(count + 3) / 4, 4 * FftPlan<LOGM>::T, tr_smem<LOGM>(), s>>>(out, in, tw, count, scale);
    return cudaGetLastError();
}
template <int LOGM, typename Torus>
static cudaError_t launch_s2t(Torus* out, const cplx* in, const cplx* tw, int count, double scale, cudaStream_t s) {
    auto kern = spectrum_to_torus_kernel<LOGM, Torus>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem<LOGM>());
    if (e != cudaSuccess) return e;
    if (count <= 0) return cudaSuccess;
    kern<<<(count + 3) / 4, 4 * FftPlan<LOGM>::T, tr_smem<LOGM>(), s>>>(out, in, tw, count, scale);
    return cudaGetLastError();
}

cudaError_t launch_poly_to_spectrum32(cplx* out, const int32_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_p2s<9, int32_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_p2s<10, int32_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_poly_to_spectrum64(cplx* out, const int64_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_p2s<9, int64_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_p2s<10, int64_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_spectrum_to_torus32(int32_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_s2t<9, int32_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_s2t<10, int32_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_spectrum_to_torus64(int64_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_s2t<9, int64_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_s2t<10, int64_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}

// res += a (.) b, complex, slot by slot (LagrangeHalfCPolynomialAddMulASM, cb/spqlios/lagrangehalfc_impl_fma.s:78-135)
__global__ void spectrum_addmul_kernel(cplx* __restrict__ res, const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx r = res[i];
        cfma(r, a[i], b[i]);
        res[i] = r;
    }
}
cudaError_t launch_spectrum_addmul(cplx* res, const cplx* a, const cplx* b, size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    spectrum_addmul_kernel<<<grid, 256, 0, s>>>(res, a, b, n);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
