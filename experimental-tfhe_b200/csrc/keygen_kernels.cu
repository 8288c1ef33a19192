// keygen_kernels.cu -- cloud-key generation on the device.  The reference builds its keys inside
// Globals::Globals (cb/poc_CircuitBootstrapping.cpp:342-423) on one core in ~100 s, most of it the 327,840 TLWE encryptions of the
// private key-switching key (:406-419).  Here every key row is one CTA:
//   LWE rows   (lweCreateKeySwitchKey_fromArray cb/lwe_functions.cpp:113-133; preKS poc:372-383):  a uniform, b = m + e + <a, s>
//   TLWE rows  (tLweSymEncryptZero cb/tlwe_functions.cpp:75-90, tGswSymEncrypt :122-132,174-177; poc:191-227,388-391,406-419):
//              a uniform, b = e + a * K mod X^N + 1 (exact wrap-around arithmetic, K binary: a sum over the set key bits),
//              then the gadget / key-switch message on coefficient 0 of polynomial z.
// Randomness: Philox4x32-10, counter = (element, row, stream), key = seed: every element of every row has its own counter, so the
// result does not depend on the launch geometry.  Gaussians by Box-Muller in FP64, scaled to the torus like the reference's
// gaussian32 / gaussian64 (cb/generic_utils.h:175-189).  Outputs are in the RAW host layouts of include/tfhe_b200.h, so the same
// device-side transform / repack as for host-supplied keys follows.
#include "engine.h"

namespace tfhe_b200 {

struct Philox { uint32_t c[4]; };
__device__ __forceinline__ Philox philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox p; p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
    return p;
}
// standard normal from two 32-bit words
__device__ __forceinline__ double box_muller(uint32_t x, uint32_t y) {
    const double u1 = ((double)x + 1.0) * (1.0 / 4294967296.0);        // (0, 1]
    const double u2 = (double)y * (1.0 / 4294967296.0);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
template <typename Torus> __device__ __forceinline__ Torus gaussian_torus(double stdev, uint32_t x, uint32_t y);
template <> __device__ __forceinline__ int32_t gaussian_torus<int32_t>(double stdev, uint32_t x, uint32_t y) {
    return (int32_t)(uint32_t)(int64_t)(stdev * box_muller(x, y) * 4294967296.0);                 // generic_utils.h:175-181
}
template <> __device__ __forceinline__ int64_t gaussian_torus<int64_t>(double stdev, uint32_t x, uint32_t y) {
    return (int64_t)(uint64_t)__double2ll_rz(stdev * box_muller(x, y) * 18446744073709551616.0);   // :183-189
}

// ---- LWE key-switching key rows: out[row][n_out+1], row = (i * t + j) * base + d, message s_in[i] * d * 2^(32-(j+1) basebit)
__global__ void __launch_bounds__(128) lwe_ks_keygen_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ s_in, const int32_t* __restrict__ s_out,
                                                            int n_out, int t, int basebit, double stdev, uint64_t seed, uint32_t stream) {
    const uint32_t row = blockIdx.x;
    const int base = 1 << basebit;
    const int d = row % base, j = (row / base) % t, i = row / (base * t);
    int32_t* dst = out + (size_t)row * (n_out + 1);
    uint32_t dot = 0;
    for (int e = threadIdx.x; e < n_out; e += 128) {
        const uint32_t a = philox4x32((uint32_t)e, row, stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)).c[0];
        dst[e] = (int32_t)a;
        dot += a * (uint32_t)s_out[e];
    }
    __shared__ uint32_t red[128];
    red[threadIdx.x] = dot;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) {
        const Philox g = philox4x32(0xFFFFFFFFu, row, stream, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t mess = ((uint32_t)(s_in[i] * d)) * (1u << (32 - (j + 1) * basebit));
        dst[n_out] = (int32_t)(mess + (uint32_t)gaussian_torus<int32_t>(stdev, g.c[0], g.c[1]) + red[0]);
    }
}

// ---- TLWE rows: out[row][2][N] = (a, b = e + a * K), then out[row][z][0] += message(row)
//   GADGET rows (bootstrapping key, row = (i * 2l + bloc * l + j)): z = bloc, message = s[i] * 2^(W - (j+1) Bgbit)
//   PRIVKS rows (row = ((z * rows_i + i) * t + j) * base + d):      z = z,    message = s[i] * d * 2^(32 - (j+1) basebit)
enum { TLWE_GADGET = 0, TLWE_PRIVKS = 1 };
struct TlweGenArgs {
    void* out; const int32_t* key_bits_idx; int key_weight;       // indices of the set bits of the binary TLWE key K
    const int32_t* s;                                             // the key whose bits / coefficients are being encrypted
    int kind, l, Bgbit, rows_i, t, basebit;
    double stdev; uint64_t seed; uint32_t stream; size_t row0;
};
template <typename Torus, int N>
__global__ void __launch_bounds__(N / 2) tlwe_keygen_kernel(const TlweGenArgs A) {
    typedef typename std::conditional<sizeof(Torus) == 4, uint32_t, uint64_t>::type U;
    __shared__ U a[N];
    __shared__ int32_t kidx[N];
    const size_t row = A.row0 + blockIdx.x;
    const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    for (int e = threadIdx.x; e < N; e += N / 2) {
        const Philox p = philox4x32((uint32_t)e, rlo, A.stream ^ (rhi << 8), 0u, k0, k1);
        a[e] = sizeof(Torus) == 4 ? (U)p.c[0] : (U)(((uint64_t)p.c[1] << 32) | p.c[0]);
    }
    for (int e = threadIdx.x; e < A.key_weight; e += N / 2) kidx[e] = A.key_bits_idx[e];
    __syncthreads();
    Torus* dst = reinterpret_cast<Torus*>(A.out) + (size_t)blockIdx.x * 2 * N;
    for (int j = threadIdx.x; j < N; j += N / 2) {
        const Philox g = philox4x32((uint32_t)j, rlo, A.stream ^ (rhi << 8), 1u, k0, k1);
        U b = (U)gaussian_torus<Torus>(A.stdev, g.c[0], g.c[1]);
        for (int e = 0; e < A.key_weight; e++) {                   // (a * K)[j] = sum over set bits i of +-a[(j - i) mod N]
            const int idx = j - kidx[e];
            b += idx >= 0 ? a[idx] : (U)0 - a[idx + N];
        }
        dst[j] = (Torus)a[j];
        dst[N + j] = (Torus)b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int z; U mess;
        if (A.kind == TLWE_GADGET) {
            const int p = (int)(row % (size_t)(2 * A.l)); const size_t i = row / (size_t)(2 * A.l);
            z = p / A.l; const int j = p % A.l;
            mess = (U)(int64_t)A.s[i] * ((U)1 << (sizeof(Torus) * 8 - (j + 1) * A.Bgbit));
        } else {
            const int base = 1 << A.basebit;
            const int d = (int)(row % base), j = (int)((row / base) % A.t);
            const size_t i = (row / ((size_t)base * A.t)) % (size_t)A.rows_i;
            z = (int)(row / ((size_t)base * A.t * A.rows_i));
            mess = (U)((uint32_t)((uint32_t)A.s[i] << (32 - (j + 1) * A.basebit)) * (uint32_t)d);
        }
        dst[z * N] = (Torus)((U)dst[z * N] + mess);
    }
}

cudaError_t launch_lwe_ks_keygen(int32_t* out, const int32_t* s_in, const int32_t* s_out, int rows_in, int n_out, int t, int basebit, double stdev,
                                 uint64_t seed, uint32_t stream, cudaStream_t st) {
    const unsigned rows = (unsigned)rows_in * t * (1u << basebit);
    if (!rows) return cudaSuccess;
    lwe_ks_keygen_kernel<<<rows, 128, 0, st>>>(out, s_in, s_out, n_out, t, basebit, stdev, seed, stream);
    return cudaGetLastError();
}
cudaError_t launch_tlwe_gadget_keygen32(int32_t* out, const int32_t* kidx, int kweight, const int32_t* s, int n, int l, int Bgbit, double stdev,
                                        uint64_t seed, uint32_t stream, cudaStream_t st) {
    TlweGenArgs A{out, kidx, kweight, s, TLWE_GADGET, l, Bgbit, 0, 0, 0, stdev, seed, stream, 0};
    const unsigned rows = (unsigned)n * 2 * l;
    if (!rows) return cudaSuccess;
    tlwe_keygen_kernel<int32_t, 1024><<<rows, 512, 0, st>>>(A);
    return cudaGetLastError();
}
cudaError_t launch_tlwe_gadget_keygen64(int64_t* out, const int32_t* kidx, int kweight, const int32_t* s, int n, int l, int Bgbit, double stdev,
                                        uint64_t seed, uint32_t stream, cudaStream_t st) {
    TlweGenArgs A{out, kidx, kweight, s, TLWE_GADGET, l, Bgbit, 0, 0, 0, stdev, seed, stream, 0};
    const unsigned rows = (unsigned)n * 2 * l;
    if (!rows) return cudaSuccess;
    tlwe_keygen_kernel<int64_t, 2048><<<rows, 1024, 0, st>>>(A);
    return cudaGetLastError();
}
// rows [row0, row0 + nrows) of the private key-switching key [2][rows_i][t][base] x TLWE32(N = 1024)
cudaError_t launch_tlwe_privks_keygen(int32_t* out, const int32_t* kidx, int kweight, const int32_t* s, int rows_i, int t, int basebit, double stdev,
                                      uint64_t seed, uint32_t stream, size_t row0, size_t nrows, cudaStream_t st) {
    TlweGenArgs A{out, kidx, kweight, s, TLWE_PRIVKS, 0, 0, rows_i, t, basebit, stdev, seed, stream, row0};
    if (!nrows) return cudaSuccess;
    tlwe_keygen_kernel<int32_t, 1024><<<(unsigned)nrows, 512, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
