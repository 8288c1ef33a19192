// exact_ntt.cuh -- exact negacyclic products for the Torus64 path (the reference's answers to "FP64 loses ~28 bits
// at Torus64" are the exact `fake FFT' build, cb/poc_CircuitBootstrapping.cpp:285-316 -> Karatsuba cb/poc_karatsuba.cpp:135-206, and
// the 128-bit FFT hp/code.cpp:391-512).
//
// Method: number-theoretic transform over the Goldilocks prime p = 2^64 - 2^32 + 1.  An external product is
//     res_q = sum_p digit_p (*) BK[p][q]   mod X^N + 1, mod 2^64,   |digit| <= Bg/2.
// A torus coefficient t only matters mod 2^64, so it is split into two 32-bit limbs t = t_lo + 2^32 t_hi; each limb product, summed over
// the 2l digit polynomials, is an integer below 2l * N * Bg/2 * 2^32 <= 2^57 in magnitude (N = 2048, l = 6, Bg = 2^10), far inside
// (-p/2, p/2): computed mod p, lifted to its centred representative, it IS the integer, and
//     res = lift(sum_lo) + (lift(sum_hi) << 32)   mod 2^64
// is bit-identical to the schoolbook product (orc_tGsw64ExternMulToTLwe_exact).  No floating point anywhere on this path.
//
// Transform: the usual in-place negacyclic pair -- forward Cooley-Tukey with the powers of psi (a primitive 2N-th root of unity) in
// bit-reversed order, natural order in, bit-reversed out; inverse Gentleman-Sande back to natural order, scaled by N^-1.  Products are
// taken slot by slot in the bit-reversed domain.
//
// Every function here is __host__ __device__: tests/cpp/ntt_host_check.cpp compiles this header with g++ and checks the field
// arithmetic and the transform against a schoolbook negacyclic product without a GPU.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

namespace tfhe_b200 {

static const uint64_t GL_P = 0xFFFFFFFF00000001ull;       // 2^64 - 2^32 + 1
static const uint64_t GL_EPS = 0xFFFFFFFFull;             // 2^32 - 1 = 2^64 mod p
static const uint64_t GL_GEN = 7;                         // generator of the multiplicative group

// canonical representatives in [0, p) throughout
GL_HD uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a) s += GL_EPS;                               // wrapped past 2^64: + (2^64 mod p); cannot wrap again (a, b < p)
    return s >= GL_P ? s - GL_P : s;
}
GL_HD uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
GL_HD uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

GL_HD void gl_mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b; hi = __umul64hi(a, b);
#else
    const unsigned __int128 w = (unsigned __int128)a * b;
    lo = (uint64_t)w; hi = (uint64_t)(w >> 64);
#endif
}
// (hi, lo) mod p with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p):  lo - hi_hi + hi_lo * (2^32 - 1)
GL_HD uint64_t gl_reduce128(uint64_t hi, uint64_t lo) {
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;                         // borrowed 2^64 = p + eps: take eps back
    const uint64_t t1 = hi_lo * GL_EPS;                   // < 2^64
    uint64_t t2 = t0 + t1;
    if (t2 < t1) t2 += GL_EPS;
    return t2 >= GL_P ? t2 - GL_P : t2;
}
GL_HD uint64_t gl_mul(uint64_t a, uint64_t b) {
    uint64_t hi, lo;
    gl_mul64(a, b, hi, lo);
    return gl_reduce128(hi, lo);
}
GL_HD uint64_t gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
    return r;
}
GL_HD uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }
// signed integer -> field element and back (centred lift)
GL_HD uint64_t gl_from_i64(int64_t x) { return x >= 0 ? (uint64_t)x : GL_P - (uint64_t)(-x); }        // |x| < p
GL_HD int64_t gl_lift(uint64_t a) { return a > (GL_P >> 1) ? -(int64_t)(GL_P - a) : (int64_t)a; }      // |result| < 2^63

// ---- tables, N a power of two <= 2^31: psi_rev[k] = psi^bitrev(k), psi_inv_rev[k] = psi^-bitrev(k), k < N
static inline unsigned gl_bitrev(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static inline void gl_make_tables(int logN, uint64_t* psi_rev, uint64_t* psi_inv_rev, uint64_t* n_inv) {
    const int N = 1 << logN;
    const uint64_t psi = gl_pow(GL_GEN, (GL_P - 1) >> (logN + 1));          // primitive 2N-th root of unity
    const uint64_t psi_inv = gl_inv(psi);
    for (int k = 0; k < N; k++) {
        const unsigned e = gl_bitrev((unsigned)k, logN);
        psi_rev[k] = gl_pow(psi, e);
        psi_inv_rev[k] = gl_pow(psi_inv, e);
    }
    *n_inv = gl_inv((uint64_t)N);
}

// ---- one butterfly stage, butterfly index b in [0, N/2).  Serial code loops b; the kernels give one b to each thread.
// forward stage with m blocks (m = 1, 2, 4, ..., N/2), t = N / (2m)
GL_HD void gl_fwd_butterfly(uint64_t* a, const uint64_t* psi_rev, int m, int t, int b) {
    const int i = b / t, j = 2 * i * t + (b - i * t);
    const uint64_t S = psi_rev[m + i];
    const uint64_t U = a[j], V = gl_mul(a[j + t], S);
    a[j] = gl_add(U, V); a[j + t] = gl_sub(U, V);
}
// inverse stage with h = m/2 blocks (m = N, N/2, ..., 2), t = N / m
GL_HD void gl_inv_butterfly(uint64_t* a, const uint64_t* psi_inv_rev, int h, int t, int b) {
    const int i = b / t, j = 2 * i * t + (b - i * t);
    const uint64_t S = psi_inv_rev[h + i];
    const uint64_t U = a[j], V = a[j + t];
    a[j] = gl_add(U, V); a[j + t] = gl_mul(gl_sub(U, V), S);
}
static inline void gl_ntt_forward_serial(uint64_t* a, const uint64_t* psi_rev, int N) {
    for (int m = 1, t = N / 2; m < N; m *= 2, t /= 2)
        for (int b = 0; b < N / 2; b++) gl_fwd_butterfly(a, psi_rev, m, t, b);
}
static inline void gl_ntt_inverse_serial(uint64_t* a, const uint64_t* psi_inv_rev, uint64_t n_inv, int N) {
    for (int h = N / 2, t = 1; h >= 1; h /= 2, t *= 2)
        for (int b = 0; b < N / 2; b++) gl_inv_butterfly(a, psi_inv_rev, h, t, b);
    for (int j = 0; j < N; j++) a[j] = gl_mul(a[j], n_inv);
}

}  // namespace tfhe_b200
