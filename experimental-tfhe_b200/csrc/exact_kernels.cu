// exact_kernels.cu -- the Torus64 blind rotation with EXACT external products (exact_ntt.cuh: Goldilocks NTT, two 32-bit limbs per key
// coefficient).  Same loop as the FP64 kernel (cb/poc_CircuitBootstrapping.cpp:580-642 with D1-D3 corrected; br_kernels.cu) with
// tGswFFTExternMulToTLwe replaced by the reference's exact build (`fake FFT', :285-316 -> Karatsuba cb/poc_karatsuba.cpp:135-206):
// every accumulator is bit-identical to orc_tGsw64ExternMulToTLwe_exact applied step by step.  It answers "FP64 keeps 53 of ~85
// product bits"  and serves as the noise-free yardstick for the FP64 path; it is the slower of the two by design.
//
// One CTA of N/2 threads owns one accumulator.  Shared memory: ACC int64[2][N] | digit polynomial / its spectrum u64[N] |
// four spectral accumulators (q, limb) u64[4][N] | psi tables 2 x u64[N]  = 144 KB at N = 2048.
// Per CMUX: 2l forward transforms (one butterfly per thread per stage), 2l x 4 slot-wise multiply-accumulates against the
// NTT-domain key (read coalesced from L2), 4 inverse transforms run side by side (4 butterflies per thread per stage), recombination
// lift(lo) + (lift(hi) << 32) into ACC.
#include "engine.h"
#include "exact_ntt.cuh"
#include "tree_fft.cuh"

namespace tfhe_b200 {

// digit offset of the Torus64 decomposition WITH rounding bit (cb/poc_CircuitBootstrapping.cpp:349-350)
__device__ __forceinline__ uint64_t exact_decomp_offset(int l, int Bgbit) {
    uint64_t t = 0;
    for (int i = 0; i <= l; i++) t |= 1ull << (63 - i * Bgbit);
    return t;
}

template <int LOGN> struct ExactSmem {
    static constexpr int N = 1 << LOGN;
    static constexpr size_t ACC = sizeof(int64_t) * 2 * N;
    static constexpr size_t WORK = sizeof(uint64_t) * N;
    static constexpr size_t SPEC = sizeof(uint64_t) * 4 * N;
    static constexpr size_t TAB = sizeof(uint64_t) * 2 * N;
    static constexpr size_t TOTAL = ACC + WORK + SPEC + TAB;
};

// coefficient-domain key polynomials (Torus64) -> NTT domain, two limbs each: out[poly][limb][N]
template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 1)) exact_key_kernel(uint64_t* __restrict__ out, const int64_t* __restrict__ in,
                                                                    const uint64_t* __restrict__ psi_rev) {
    constexpr int N = 1 << LOGN, H = N / 2;
    __shared__ uint64_t a[N];
    const int tid = threadIdx.x;
    const size_t poly = blockIdx.x >> 1; const int limb = blockIdx.x & 1;
    const int64_t* src = in + poly * N;
    for (int j = tid; j < N; j += H) { const uint64_t t = (uint64_t)src[j]; a[j] = limb ? (t >> 32) : (t & GL_EPS); }
    __syncthreads();
    for (int m = 1, t = H; m < N; m *= 2, t /= 2) { gl_fwd_butterfly(a, psi_rev, m, t, tid); __syncthreads(); }
    uint64_t* dst = out + (poly * 2 + limb) * N;
    for (int j = tid; j < N; j += H) dst[j] = a[j];
}

template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 1), 1) exact_blind_rotate_kernel(const ExactArgs A) {
    typedef ExactSmem<LOGN> S;
    constexpr int N = 1 << LOGN, H = N / 2, M = N / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int64_t* acc = reinterpret_cast<int64_t*>(smem_raw);
    uint64_t* work = reinterpret_cast<uint64_t*>(smem_raw + S::ACC);
    uint64_t* spec = reinterpret_cast<uint64_t*>(smem_raw + S::ACC + S::WORK);          // [q][limb][N]
    uint64_t* psi_rev = reinterpret_cast<uint64_t*>(smem_raw + S::ACC + S::WORK + S::SPEC);
    uint64_t* psi_inv_rev = psi_rev + N;
    const int tid = threadIdx.x;
    const int n_mu = A.n_mu > 0 ? A.n_mu : 1;
    const long unit = blockIdx.x;
    const int ct = (int)(unit / n_mu), w = (int)(unit % n_mu);
    const int n = A.n, l = A.l, Bgbit = A.Bgbit;
    for (int j = tid; j < N; j += H) { psi_rev[j] = A.psi_rev[j]; psi_inv_rev[j] = A.psi_inv_rev[j]; }

    // ---- initial accumulator (same conventions as blind_rotate_kernel<10,int64_t>)
    int64_t mu = A.mu;
    if (A.mode == BR_LWE && A.mu_bgbit > 0) mu = (int64_t)(1ull << (64 - (w + 1) * A.mu_bgbit));
    const int32_t* bara = A.bara + (size_t)ct * (A.mode == BR_LWE ? n + 1 : n);
    if (A.mode == BR_ACCUM) {
        const int64_t* src = A.accum + (size_t)ct * 2 * N;
        for (int j = tid; j < 2 * N; j += H) acc[j] = src[j];
    } else {
        const int barb = bara[n];
        const int rot = (2 * N - barb) & (2 * N - 1);                  // X^(2N - barb) v  (cb/lwe_functions.cpp:385; D3)
        for (int j = tid; j < N; j += H) {
            const int idx = (j - rot) & (2 * N - 1);
            const int k0 = idx & (N - 1);
            const int64_t val = (k0 < M) ? (int64_t)(0 - (uint64_t)(mu / 2)) : (mu / 2);      // poc:552-553
            acc[j] = 0;
            acc[N + j] = (idx & N) ? (int64_t)(0 - (uint64_t)val) : val;
        }
    }
    __syncthreads();

    const uint64_t offset = exact_decomp_offset(l, Bgbit);
    const uint32_t mask = (1u << Bgbit) - 1u;
    const int half = 1 << (Bgbit - 1);
    const size_t key_stride = (size_t)2 * l * 2 * 2 * N;               // [2l][2 q][2 limbs][N] per step
    for (int i = 0; i < n; i++) {
        const int a = bara[i];
        if (a == 0) continue;                                          // cb/lwe_functions.cpp:350
        const uint64_t* __restrict__ key = A.key + (size_t)i * key_stride;
        for (int p = 0; p < 2 * l; p++) {
            const int q = p >= l, lev = p - q * l;
            const int sh = 64 - (lev + 1) * Bgbit;
            const int64_t* aq = acc + q * N;
            for (int j = tid; j < N; j += H) {
                const uint64_t u = (uint64_t)rot_minus_one<int64_t, N>(aq, j, a) + offset;
                work[j] = gl_from_i64((int64_t)((int)((uint32_t)(u >> sh) & mask) - half));       // poc:492-515
            }
            __syncthreads();
            for (int m = 1, t = H; m < N; m *= 2, t /= 2) { gl_fwd_butterfly(work, psi_rev, m, t, tid); __syncthreads(); }
            const uint64_t* __restrict__ kp = key + (size_t)p * 4 * N;
            for (int j = tid; j < N; j += H) {
                const uint64_t d = work[j];
#pragma unroll
                for (int x = 0; x < 4; x++) {                          // x = q' * 2 + limb
                    const uint64_t prod = gl_mul(d, __ldg(kp + (size_t)x * N + j));
                    spec[x * N + j] = p == 0 ? prod : gl_add(spec[x * N + j], prod);
                }
            }
            __syncthreads();
        }
        for (int h = H, t = 1; h >= 1; h /= 2, t *= 2) {
#pragma unroll
            for (int x = 0; x < 4; x++) gl_inv_butterfly(spec + x * N, psi_inv_rev, h, t, tid);
            __syncthreads();
        }
        for (int j = tid; j < N; j += H) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int64_t lo = gl_lift(gl_mul(spec[(2 * q) * N + j], A.n_inv));
                const int64_t hi = gl_lift(gl_mul(spec[(2 * q + 1) * N + j], A.n_inv));
                acc[q * N + j] = (int64_t)((uint64_t)acc[q * N + j] + (uint64_t)lo + ((uint64_t)hi << 32));
            }
        }
        __syncthreads();
    }

    if (A.mode == BR_ACCUM) {
        int64_t* dst = A.accum + (size_t)ct * 2 * N;
        for (int j = tid; j < 2 * N; j += H) dst[j] = acc[j];
    } else {
        int64_t* out = A.out + (size_t)unit * A.out_stride;            // tLweExtractLweSampleIndex(0); PoC adds mu/2 to b (:648)
        for (int j = tid; j < N; j += H) out[j] = (j == 0) ? acc[0] : (int64_t)(0 - (uint64_t)acc[N - j]);
        if (tid == 0) out[N] = (int64_t)((uint64_t)acc[N] + (uint64_t)(mu / 2));
    }
}

cudaError_t launch_exact_key(uint64_t* out, const int64_t* in, const uint64_t* psi_rev, int N, size_t npoly, cudaStream_t s) {
    if (N != 2048) return cudaErrorInvalidValue;
    if (npoly == 0) return cudaSuccess;
    exact_key_kernel<11><<<(unsigned)(npoly * 2), 1024, 0, s>>>(out, in, psi_rev);
    return cudaGetLastError();
}
cudaError_t launch_exact_blind_rotate(const ExactArgs& a, cudaStream_t s) {
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(exact_blind_rotate_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ExactSmem<11>::TOTAL);
        if (e != cudaSuccess) return e;
        attr_done.done();
    }
    const long units = (long)a.count * (a.n_mu > 0 ? a.n_mu : 1);
    if (units <= 0) return cudaSuccess;
    exact_blind_rotate_kernel<11><<<(unsigned)units, 1024, ExactSmem<11>::TOTAL, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
