/*
 * hpfft_oracle.c -- CPU restatement of the 128-bit fixed-point anticyclic FFT of
 * /root/reference/high-precision-anticyclic-fft/src/code.cpp ("hp/").
 * TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).
 *
 * Real96 (hp/code.cpp:25-75): value = v / 2^64 with v a wrapping 128-bit integer.
 * Parity: the reference builds its twiddles with NTL RR (absent from this image and from
 * /root/reference); this file regenerates them as round(cos|sin(2 pi i/n) * 2^64) in IEEE binary128
 * (libquadmath).  That table is the definition the GPU kernel is compared against, bit for bit.
 */
#include "tfhe_oracle.h"
#include <quadmath.h>
#include <stdlib.h>

typedef __int128 i128;
typedef orc_u128 u128;

/* hp/code.cpp:148-169 (intmul_best, the variant the reference's Makefile selects: `#define intmul intmul_best`,
 * :21-23): a * b truncated to 64 fractional bits; b is a twiddle in [-1,1) (high word 0 or -1).  It equals
 * intmul_ref (:79-95) whenever a's integer part fits int32, which the reference asserts (:88-89,163-168). */
orc_u128 orc_real96_mul(orc_u128 a, orc_u128 b) {
    const uint64_t alo = (uint64_t)a, blo = (uint64_t)b;
    const int64_t ahi = (int64_t)(uint64_t)(a >> 64);
    u128 w = ((u128)alo * (u128)blo) >> 64;          /* _mulx_u64 high half  (:155) */
    w += (u128)(i128)ahi * (u128)blo;                /* signed ahi * blo      (:156-158) */
    if ((int64_t)(uint64_t)(b >> 64) < 0) w -= a;    /* b negative            (:159) */
    return w;
}

/* std::complex<Real96> product as libstdc++ evaluates operator*= :
 * re = a.re*b.re - a.im*b.im ; im = a.re*b.im + a.im*b.re  (data on the left, twiddle on the right) */
static orc_cplx96 cmul(orc_cplx96 a, orc_cplx96 b) {
    orc_cplx96 r;
    r.re = orc_real96_mul(a.re, b.re) - orc_real96_mul(a.im, b.im);
    r.im = orc_real96_mul(a.re, b.im) + orc_real96_mul(a.im, b.re);
    return r;
}

static u128 round_to_fix64(__float128 x) {
    /* round(x * 2^64) as a wrapping 128-bit integer (hp/code.cpp:249-254) */
    __float128 s = roundq(ldexpq(x, 64));
    int neg = s < 0;
    if (neg) s = -s;
    __float128 hi = floorq(ldexpq(s, -64));
    __float128 lo = s - ldexpq(hi, 64);
    u128 v = ((u128)(uint64_t)hi << 64) | (u128)(uint64_t)lo;
    return neg ? (u128)0 - v : v;
}
/* hp/code.cpp:246-261 */
static u128 accurate_cos(int i, int n) {
    i = ((i % n) + n) % n;
    if (i == 0) return (u128)UINT64_MAX;                      /* 1.0 is stored as 2^64-1 (:248) */
    return round_to_fix64(cosq(2 * M_PIq * i / n));
}
/* hp/code.cpp:263-277 */
static u128 accurate_sin(int i, int n) {
    i = ((i % n) + n) % n;
    if (i == n / 4) return (u128)UINT64_MAX;                  /* :265 */
    return round_to_fix64(sinq(2 * M_PIq * i / n));
}
/* hp/code.cpp:378-382 */
void orc_hp_precomp_iFFT(orc_cplx96* powomega, int n) {
    for (int i = 0; i < n; i++) { powomega[i].re = accurate_cos(i, n); powomega[i].im = accurate_sin(i, n); }
}
/* hp/code.cpp:384-388 */
void orc_hp_precomp_FFT(orc_cplx96* powombar, int n) {
    for (int i = 0; i < n; i++) { powombar[i].re = accurate_cos(i, n); powombar[i].im = accurate_sin((n - i) % n, n); }
}

/* hp/code.cpp:184-189 */
static u128 t64tor96(Torus64 v) { return (u128)(i128)v; }

/* hp/code.cpp:391-443 : P -> P(omega), n = 2N */
void orc_hp_iFFT(orc_cplx96* out, const Torus64* in, int n, const orc_cplx96* powomega) {
    const int ns4 = n / 4;
    for (int j = 0; j < ns4; j++) {
        orc_cplx96 z = { t64tor96(in[j]), t64tor96(in[j + ns4]) };
        out[j] = cmul(z, powomega[j]);
    }
    for (int nn = ns4; nn >= 2; nn /= 2) {
        int halfnn = nn / 2;
        for (int block = 0; block < ns4; block += nn)
            for (int off = 0; off < halfnn; off++) {
                orc_cplx96 t1 = out[block + off], t2 = out[block + off + halfnn];
                orc_cplx96 s = { t1.re + t2.re, t1.im + t2.im };
                orc_cplx96 d = { t1.re - t2.re, t1.im - t2.im };
                out[block + off] = s;
                out[block + off + halfnn] = cmul(d, powomega[(2 * (ns4 / halfnn) * off) % n]);
            }
    }
}

/* hp/code.cpp:446-512 : P(omega) -> P ; the final shift is log2(N/2) (">>10" for N=2048, :502-503) */
void orc_hp_FFT(Torus64* out, orc_cplx96* in, int n, const orc_cplx96* powombar) {
    const int ns4 = n / 4;
    int shift = 0;
    while ((1 << shift) < ns4) shift++;
    for (int nn = 2; nn <= ns4; nn *= 2) {
        int halfnn = nn / 2;
        for (int block = 0; block < ns4; block += nn)
            for (int off = 0; off < halfnn; off++) {
                orc_cplx96 t1 = in[block + off];
                orc_cplx96 t2 = cmul(in[block + off + halfnn], powombar[(2 * (ns4 / halfnn) * off) % n]);
                in[block + off].re = t1.re + t2.re; in[block + off].im = t1.im + t2.im;
                in[block + off + halfnn].re = t1.re - t2.re; in[block + off + halfnn].im = t1.im - t2.im;
            }
    }
    for (int j = 0; j < ns4; j++) {
        in[j] = cmul(in[j], powombar[j]);
        out[j] = (Torus64)(uint64_t)(in[j].re >> shift);
        out[j + ns4] = (Torus64)(uint64_t)(in[j].im >> shift);
    }
}
