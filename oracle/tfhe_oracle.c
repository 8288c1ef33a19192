/*
 * tfhe_oracle.c -- CPU restatement of the experimental-tfhe bootstrapping hot path.
 * TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).  Independent restatement: nothing here is
 * copied from the reference; each function names the reference lines whose arithmetic it follows.
 */
#include "tfhe_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define K1 1 /* cb/poc_types.h:10  (#define k 1) */

static void* xmalloc(size_t s) {
    void* p = malloc(s ? s : 1);
    if (!p) { fprintf(stderr, "tfhe_oracle: out of memory (%zu bytes)\n", s); abort(); }
    return p;
}

/* ================================================================== RNG */
/* The reference draws from a default-seeded std::default_random_engine (cb/generic_utils.h:153-167),
 * which is libstdc++ specific; the oracle owns a portable splitmix64 stream instead. */
void orc_rng_seed(orc_rng* r, uint64_t seed) { r->s = seed; r->has_spare = 0; r->spare = 0; }
uint64_t orc_rng_u64(orc_rng* r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
int orc_rng_bit(orc_rng* r) { return (int)(orc_rng_u64(r) >> 63); }
double orc_rng_normal(orc_rng* r) {
    if (r->has_spare) { r->has_spare = 0; return r->spare; }
    double u1 = ((double)(orc_rng_u64(r) >> 11) + 1.0) * (1.0 / 9007199254740992.0); /* (0,1] */
    double u2 = (double)(orc_rng_u64(r) >> 11) * (1.0 / 9007199254740992.0);          /* [0,1) */
    double rad = sqrt(-2.0 * log(u1));
    r->spare = rad * sin(6.283185307179586476925 * u2);
    r->has_spare = 1;
    return rad * cos(6.283185307179586476925 * u2);
}
/* cb/generic_utils.h:176-181: truncate stdev*N(0,1)*2^32 to an integer, add the centre */
Torus32 orc_gaussian32(orc_rng* r, Torus32 center, double stdev) {
    double val = stdev * orc_rng_normal(r) * 4294967296.0;
    return (Torus32)((uint32_t)(int64_t)val + (uint32_t)center);
}
/* cb/generic_utils.h:183-189 */
Torus64 orc_gaussian64(orc_rng* r, Torus64 center, double stdev) {
    double val = stdev * orc_rng_normal(r) * 18446744073709551616.0;
    return (Torus64)((uint64_t)(int64_t)val + (uint64_t)center);
}

/* ================================================================== portable half-complex FFT */
/* tables per ring degree N: unit roots e^{2 pi i m / (N/2)} and the twist e^{i pi j / N} */
typedef struct { int N; double *wre, *wim, *tre, *tim; } fft_tab;
static fft_tab g_tabs[8];
static int g_ntabs = 0;

static const fft_tab* get_tab(int N) {
    const fft_tab* found = NULL;
    for (int i = 0; i < g_ntabs; i++) if (g_tabs[i].N == N) found = &g_tabs[i];
    if (found) return found;
#ifdef _OPENMP
#pragma omp critical(orc_fft_tab)
#endif
    {
        for (int i = 0; i < g_ntabs; i++) if (g_tabs[i].N == N) found = &g_tabs[i];
        if (!found) {
            if (g_ntabs >= 8) { fprintf(stderr, "tfhe_oracle: too many FFT sizes\n"); abort(); }
            fft_tab t; t.N = N;
            int M = N / 2;
            t.wre = xmalloc(sizeof(double) * M); t.wim = xmalloc(sizeof(double) * M);
            t.tre = xmalloc(sizeof(double) * M); t.tim = xmalloc(sizeof(double) * M);
            const long double PI = 3.14159265358979323846264338327950288L;
            for (int m = 0; m < M; m++) {
                t.wre[m] = (double)cosl(2.0L * PI * m / M); t.wim[m] = (double)sinl(2.0L * PI * m / M);
                t.tre[m] = (double)cosl(PI * m / N);        t.tim[m] = (double)sinl(PI * m / N);
            }
            g_tabs[g_ntabs] = t;
            found = &g_tabs[g_ntabs];
            __sync_synchronize();
            g_ntabs++;
        }
    }
    return found;
}

/* "ifft" of the reference = coefficients -> evaluations.  z_j = (a_j + i a_{j+N/2}) w^j, then DIF radix-2,
 * no bit reversal, no scaling.  Restates cb/spqlios/spqlios-fft-impl.cpp:469-641 (ifft_model). */
void orc_ifft_raw(int N, double* d) {
    const fft_tab* t = get_tab(N);
    int M = N / 2;
    double* re = d; double* im = d + M;
    for (int j = 0; j < M; j++) {
        double a = re[j], b = im[j];
        re[j] = a * t->tre[j] - b * t->tim[j];
        im[j] = a * t->tim[j] + b * t->tre[j];
    }
    for (int nn = M; nn >= 2; nn >>= 1) {
        int h = nn >> 1, step = M / nn;
        for (int blk = 0; blk < M; blk += nn)
            for (int off = 0; off < h; off++) {
                int i0 = blk + off, i1 = i0 + h;
                double ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
                re[i0] = ar + br; im[i0] = ai + bi;
                double dr = ar - br, di = ai - bi;
                double wr = t->wre[off * step], wi = t->wim[off * step];
                re[i1] = dr * wr - di * wi;
                im[i1] = dr * wi + di * wr;
            }
    }
}
/* "fft" of the reference = evaluations -> coefficients, unscaled: fft(ifft(x)) = (N/2) x.
 * Restates cb/spqlios/spqlios-fft-impl.cpp:204-397 (fft_model). */
void orc_fft_raw(int N, double* d) {
    const fft_tab* t = get_tab(N);
    int M = N / 2;
    double* re = d; double* im = d + M;
    for (int nn = 2; nn <= M; nn <<= 1) {
        int h = nn >> 1, step = M / nn;
        for (int blk = 0; blk < M; blk += nn)
            for (int off = 0; off < h; off++) {
                int i0 = blk + off, i1 = i0 + h;
                double wr = t->wre[off * step], wi = -t->wim[off * step];
                double br = re[i1] * wr - im[i1] * wi;
                double bi = re[i1] * wi + im[i1] * wr;
                double ar = re[i0], ai = im[i0];
                re[i0] = ar + br; im[i0] = ai + bi;
                re[i1] = ar - br; im[i1] = ai - bi;
            }
    }
    for (int j = 0; j < M; j++) {
        double a = re[j], b = im[j];
        re[j] = a * t->tre[j] + b * t->tim[j];
        im[j] = -a * t->tim[j] + b * t->tre[j];
    }
}

/* cb/spqlios/fft_processor_spqlios.cpp:102 */
Torus32 orc_double_to_torus32(double x) { return (Torus32)(int64_t)x; }
/* cb/spqlios/fft_processor_spqlios.cpp:131-142; shift counts >= 64 (UB there) yield 0 */
Torus64 orc_double_to_torus64(double x) {
    uint64_t bits; memcpy(&bits, &x, 8);
    uint64_t val = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
    int expo = (int)((bits >> 52) & 0x7FF);
    int trans = expo - 1075;
    uint64_t val2;
    if (trans > 0) val2 = trans >= 64 ? 0 : (val << trans);
    else           val2 = -trans >= 64 ? 0 : (val >> -trans);
    return (bits >> 63) ? (Torus64)(0 - val2) : (Torus64)val2;
}

static void p_ifft_int(int N, double* res, const int32_t* a) {
    for (int i = 0; i < N; i++) res[i] = (double)a[i];
    orc_ifft_raw(N, res);
}
static void p_ifft_torus64(int N, double* res, const int64_t* a) {
    for (int i = 0; i < N; i++) res[i] = (double)a[i];
    orc_ifft_raw(N, res);
}
static void p_fft_torus32(int N, int32_t* res, const double* a) {
    double buf[N];
    const double s = 2.0 / N;
    for (int i = 0; i < N; i++) buf[i] = a[i] * s;
    orc_fft_raw(N, buf);
    for (int i = 0; i < N; i++) res[i] = orc_double_to_torus32(buf[i]);
}
static void p_fft_torus64(int N, int64_t* res, const double* a) {
    double buf[N];
    const double s = 2.0 / N;
    for (int i = 0; i < N; i++) buf[i] = a[i] * s;
    orc_fft_raw(N, buf);
    for (int i = 0; i < N; i++) res[i] = orc_double_to_torus64(buf[i]);
}
/* cb/poc_CircuitBootstrapping.cpp:263-270 (the scalar form of the asm) */
static void p_addmul(int N, double* res, const double* a, const double* b) {
    int M = N / 2;
    for (int i = 0; i < M; i++) {
        double ra = a[i], ia = a[M + i], rb = b[i], ib = b[M + i];
        res[i] += ra * rb - ia * ib;
        res[M + i] += ra * ib + ia * rb;
    }
}
static const orc_fft_backend g_portable = { p_ifft_int, p_ifft_torus64, p_fft_torus32, p_fft_torus64, p_addmul };
static const orc_fft_backend* g_backend = &g_portable;
void orc_set_fft_backend(const orc_fft_backend* b) { g_backend = b ? b : &g_portable; }
const orc_fft_backend* orc_get_fft_backend(void) { return g_backend; }

/* ================================================================== exact products */
/* par/poc_karatsuba.cpp:10-21 (naive product mod X^N+1, wrap-around) */
void orc_torus32PolynomialMultAddNaive(Torus32* result, const int32_t* p1, const Torus32* p2, int N) {
    for (int i = 0; i < N; i++) {
        uint32_t ri = 0;
        for (int j = 0; j <= i; j++) ri += (uint32_t)p1[j] * (uint32_t)p2[i - j];
        for (int j = i + 1; j < N; j++) ri -= (uint32_t)p1[j] * (uint32_t)p2[N + i - j];
        result[i] = (Torus32)((uint32_t)result[i] + ri);
    }
}
void orc_torus64PolynomialMultAddNaive(Torus64* result, const int32_t* p1, const Torus64* p2, int N) {
    for (int i = 0; i < N; i++) {
        uint64_t ri = 0;
        for (int j = 0; j <= i; j++) ri += (uint64_t)(int64_t)p1[j] * (uint64_t)p2[i - j];
        for (int j = i + 1; j < N; j++) ri -= (uint64_t)(int64_t)p1[j] * (uint64_t)p2[N + i - j];
        result[i] = (Torus64)((uint64_t)result[i] + ri);
    }
}
/* Exact product by a {0,1} key through the double FFT: |coef| < 2^(1+32+log2 N) fits 53 bits, so
 * round-to-nearest of the inverse transform is the exact integer.  Used by key generation only,
 * where the reference calls Karatsuba (cb/poc_CircuitBootstrapping.cpp:151,200). */
void orc_torus32PolynomialMultAddBinKey(Torus32* result, const double* keyFFT, const Torus32* p2, int N) {
    double a[N], acc[N];
    for (int i = 0; i < N; i++) a[i] = (double)p2[i];
    orc_ifft_raw(N, a);
    memset(acc, 0, sizeof(double) * N);
    p_addmul(N, acc, keyFFT, a);
    orc_fft_raw(N, acc);
    const double s = 2.0 / N;
    for (int i = 0; i < N; i++) result[i] = (Torus32)((uint32_t)result[i] + (uint32_t)(int64_t)llrint(acc[i] * s));
}
void orc_torus64PolynomialMultAddBinKey(Torus64* result, const double* keyFFT, const Torus64* p2, int N) {
    double lo[N], hi[N], acc[N];
    for (int i = 0; i < N; i++) {
        uint64_t v = (uint64_t)p2[i];
        lo[i] = (double)(uint32_t)v; hi[i] = (double)(uint32_t)(v >> 32);
    }
    const double s = 2.0 / N;
    orc_ifft_raw(N, lo); orc_ifft_raw(N, hi);
    memset(acc, 0, sizeof(double) * N);
    p_addmul(N, acc, keyFFT, lo);
    orc_fft_raw(N, acc);
    for (int i = 0; i < N; i++) result[i] = (Torus64)((uint64_t)result[i] + (uint64_t)(int64_t)llrint(acc[i] * s));
    memset(acc, 0, sizeof(double) * N);
    p_addmul(N, acc, keyFFT, hi);
    orc_fft_raw(N, acc);
    for (int i = 0; i < N; i++) result[i] = (Torus64)((uint64_t)result[i] + ((uint64_t)(int64_t)llrint(acc[i] * s) << 32));
}
static double* key_to_fft(const int32_t* key, int N) {
    double* f = xmalloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) f[i] = (double)key[i];
    orc_ifft_raw(N, f);
    return f;
}

/* ================================================================== numeric */
/* cb/numeric_functions.cpp:54-60 */
int orc_modSwitchFromTorus32(Torus32 phase, int Msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t half_interval = interv / 2;
    uint64_t phase64 = ((uint64_t)(uint32_t)phase << 32) + half_interval;
    return (int)(phase64 / interv);
}
/* cb/numeric_functions.cpp:62-67 */
Torus32 orc_modSwitchToTorus32(int mu, int Msize) {
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t phase64 = (uint64_t)(int64_t)mu * interv;
    return (Torus32)(phase64 >> 32);
}
/* cb/numeric_functions.cpp:304-323 */
void orc_torusPolynomialMulByXaiMinusOne(Torus32* out, int a, const Torus32* in, int N) {
    const uint32_t* x = (const uint32_t*)in; uint32_t* y = (uint32_t*)out;
    if (a < N) {
        for (int i = 0; i < a; i++) y[i] = 0u - x[i - a + N] - x[i];
        for (int i = a; i < N; i++) y[i] = x[i - a] - x[i];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) y[i] = x[i - aa + N] - x[i];
        for (int i = aa; i < N; i++) y[i] = 0u - x[i - aa] - x[i];
    }
}
/* cb/numeric_functions.cpp:327-347 */
void orc_torusPolynomialMulByXai(Torus32* out, int a, const Torus32* in, int N) {
    const uint32_t* x = (const uint32_t*)in; uint32_t* y = (uint32_t*)out;
    if (a < N) {
        for (int i = 0; i < a; i++) y[i] = 0u - x[i - a + N];
        for (int i = a; i < N; i++) y[i] = x[i - a];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) y[i] = x[i - aa + N];
        for (int i = aa; i < N; i++) y[i] = 0u - x[i - aa];
    }
}
/* 64-bit twins: the corrected form of cb/poc_CircuitBootstrapping.cpp:591-598 (defect D2) */
void orc_torus64PolynomialMulByXaiMinusOne(Torus64* out, int a, const Torus64* in, int N) {
    const uint64_t* x = (const uint64_t*)in; uint64_t* y = (uint64_t*)out;
    if (a < N) {
        for (int i = 0; i < a; i++) y[i] = 0ull - x[i - a + N] - x[i];
        for (int i = a; i < N; i++) y[i] = x[i - a] - x[i];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) y[i] = x[i - aa + N] - x[i];
        for (int i = aa; i < N; i++) y[i] = 0ull - x[i - aa] - x[i];
    }
}
void orc_torus64PolynomialMulByXai(Torus64* out, int a, const Torus64* in, int N) {
    const uint64_t* x = (const uint64_t*)in; uint64_t* y = (uint64_t*)out;
    if (a < N) {
        for (int i = 0; i < a; i++) y[i] = 0ull - x[i - a + N];
        for (int i = a; i < N; i++) y[i] = x[i - a];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) y[i] = x[i - aa + N];
        for (int i = aa; i < N; i++) y[i] = 0ull - x[i - aa];
    }
}

/* ================================================================== gate path (Torus32) */
void orc_gate_params_default(orc_gate_params* p) {
    p->n = 500; p->N = 1024; p->k = 1; p->bk_l = 2; p->bk_Bgbit = 10;   /* misc/params-gb.html:124-131 */
    p->ks_t = 8; p->ks_basebit = 2;                                     /* [UPSTREAM] 80-bit default */
    p->bk_stdev = 7.18e-9; p->ks_stdev = 2.44e-5;                       /* [UPSTREAM] */
}

/* cb/lwe_functions.cpp:43-54 */
void orc_lweSymEncrypt(Torus32* result, Torus32 message, double alpha, const int32_t* key, int n, orc_rng* r) {
    uint32_t b = (uint32_t)orc_gaussian32(r, message, alpha);
    for (int i = 0; i < n; i++) {
        uint32_t a = (uint32_t)orc_rng_u64(r);
        result[i] = (Torus32)a;
        b += a * (uint32_t)key[i];
    }
    result[n] = (Torus32)b;
}
/* cb/lwe_functions.cpp:56-65 */
Torus32 orc_lwePhase(const Torus32* sample, const int32_t* key, int n) {
    uint32_t axs = 0;
    for (int i = 0; i < n; i++) axs += (uint32_t)sample[i] * (uint32_t)key[i];
    return (Torus32)((uint32_t)sample[n] - axs);
}

/* TLWE32 encryption of zero: cb/poc_CircuitBootstrapping.cpp:143-152 (b = e + a*s) */
static void tlwe32_encrypt_zero(Torus32* c /*[2][N]*/, double stdev, const double* keyFFT, int N, orc_rng* r) {
    Torus32* a = c; Torus32* b = c + N;
    for (int j = 0; j < N; j++) b[j] = orc_gaussian32(r, 0, stdev);
    for (int j = 0; j < N; j++) a[j] = (Torus32)(uint32_t)orc_rng_u64(r);
    orc_torus32PolynomialMultAddBinKey(b, keyFFT, a, N);
}

orc_gate_keys* orc_gate_keygen(const orc_gate_params* p, uint64_t seed) {
    orc_gate_keys* K = xmalloc(sizeof(*K));
    K->p = *p;
    const int n = p->n, N = p->N, l = p->bk_l, kpl = 2 * l, t = p->ks_t, base = 1 << p->ks_basebit;
    orc_rng r; orc_rng_seed(&r, seed);
    K->lwe_key = xmalloc(sizeof(int32_t) * n);
    K->tlwe_key = xmalloc(sizeof(int32_t) * N);
    for (int i = 0; i < n; i++) K->lwe_key[i] = orc_rng_bit(&r);   /* cb/lwe_functions.cpp:35-41 */
    for (int i = 0; i < N; i++) K->tlwe_key[i] = orc_rng_bit(&r);
    double* keyFFT = key_to_fft(K->tlwe_key, N);
    /* bk[i] = TGSW(s_i): zero rows + s_i * h on the diagonal blocks.
     * cb/tgsw_functions.cpp:122-142,174-177 ; cb/lwe_functions.cpp:504-506 */
    K->bk = xmalloc(sizeof(Torus32) * (size_t)n * kpl * 2 * N);
    for (int i = 0; i < n; i++)
        for (int bloc = 0; bloc <= K1; bloc++)
            for (int j = 0; j < l; j++) {
                Torus32* row = K->bk + (((size_t)i * kpl + bloc * l + j) * 2) * N;
                tlwe32_encrypt_zero(row, p->bk_stdev, keyFFT, N, &r);
                uint32_t h = 1u << (32 - (j + 1) * p->bk_Bgbit);
                row[bloc * N + 0] = (Torus32)((uint32_t)row[bloc * N + 0] + (uint32_t)K->lwe_key[i] * h);
            }
    free(keyFFT);
    /* ks[i][j][d] = LWE( s'_i * d * 2^(32-(j+1)basebit) )  cb/lwe_functions.cpp:120-133 */
    K->ks = xmalloc(sizeof(Torus32) * (size_t)N * t * base * (n + 1));
    for (int i = 0; i < N; i++)
        for (int j = 0; j < t; j++)
            for (int d = 0; d < base; d++) {
                Torus32 x = (Torus32)((uint32_t)(K->tlwe_key[i] * d) * (1u << (32 - (j + 1) * p->ks_basebit)));
                orc_lweSymEncrypt(K->ks + (((size_t)i * t + j) * base + d) * (n + 1), x, p->ks_stdev, K->lwe_key, n, &r);
            }
    K->bkFFT = NULL;
    orc_gate_keys_rebuild_fft(K);
    return K;
}
/* tGswToFFTConvert: cb/tgsw_functions.cpp:389-394 ; cb/lwe_functions.cpp:309-313 */
void orc_gate_keys_rebuild_fft(orc_gate_keys* K) {
    const size_t total = (size_t)K->p.n * 2 * K->p.bk_l * 2;
    const int N = K->p.N;
    if (!K->bkFFT) K->bkFFT = xmalloc(sizeof(double) * total * N);
    for (size_t q = 0; q < total; q++) g_backend->ifft_int(N, K->bkFFT + q * N, K->bk + q * N);
}
void orc_gate_keys_free(orc_gate_keys* K) {
    if (!K) return;
    free(K->lwe_key); free(K->tlwe_key); free(K->bk); free(K->bkFFT); free(K->ks); free(K);
}

/* cb/tgsw_functions.cpp:30-36 */
uint32_t orc_tgsw32_offset(int l, int Bgbit) {
    uint32_t temp1 = 0;
    for (int i = 0; i < l; i++) temp1 += 1u << (32 - (i + 1) * Bgbit);
    return temp1 * (uint32_t)((1 << Bgbit) / 2);
}
/* cb/tgsw_functions.cpp:224-337 (scalar branch) */
void orc_tGswTorus32PolynomialDecompH(int32_t* result, const Torus32* sample, int N, int l, int Bgbit) {
    const uint32_t maskMod = (1u << Bgbit) - 1;
    const int32_t halfBg = (1 << Bgbit) / 2;
    const uint32_t offset = orc_tgsw32_offset(l, Bgbit);
    for (int p = 0; p < l; p++) {
        const int decal = 32 - (p + 1) * Bgbit;
        for (int j = 0; j < N; j++) {
            uint32_t temp1 = (((uint32_t)sample[j] + offset) >> decal) & maskMod;
            result[p * N + j] = (int32_t)temp1 - halfBg;
        }
    }
}
/* cb/tgsw_functions.cpp:424-449 */
void orc_tGswFFTExternMulToTLwe(Torus32* accum, const double* gswFFT, int N, int l, int Bgbit) {
    const int kpl = 2 * l;
    int32_t deca[kpl * N];
    double decaFFT[kpl * N];
    double tmpa[2 * N];
    for (int i = 0; i <= K1; i++) orc_tGswTorus32PolynomialDecompH(deca + i * l * N, accum + i * N, N, l, Bgbit);
    for (int p = 0; p < kpl; p++) g_backend->ifft_int(N, decaFFT + p * N, deca + p * N);
    memset(tmpa, 0, sizeof(tmpa));
    for (int p = 0; p < kpl; p++)
        for (int q = 0; q <= K1; q++)       /* tLweFFTAddMulRTo cb/tlwe_functions.cpp:318-325 */
            g_backend->addmul(N, tmpa + q * N, decaFFT + p * N, gswFFT + ((size_t)p * 2 + q) * N);
    for (int q = 0; q <= K1; q++) g_backend->fft_torus32(N, accum + q * N, tmpa + q * N); /* tLweFromFFTConvert :299-305 */
}
/* cb/tgsw_functions.cpp:150-164 : the exact (coefficient-domain) external product, tLweAddMulRTo per row */
void orc_tGswExternMulToTLwe(Torus32* accum, const Torus32* gsw, int N, int l, int Bgbit) {
    const int kpl = 2 * l;
    int32_t dec[kpl * N];
    for (int i = 0; i <= K1; i++) orc_tGswTorus32PolynomialDecompH(dec + i * l * N, accum + i * N, N, l, Bgbit);
    memset(accum, 0, sizeof(Torus32) * 2 * N);
    for (int p = 0; p < kpl; p++)
        for (int q = 0; q <= K1; q++)
            orc_torus32PolynomialMultAddNaive(accum + q * N, dec + p * N, gsw + ((size_t)p * 2 + q) * N, N);
}
/* cb/lwe_functions.cpp:328-333 */
void orc_tfhe_MuxRotate_FFT(Torus32* result, const Torus32* accum, const double* bki, int barai, int N, int l, int Bgbit) {
    for (int q = 0; q <= K1; q++) orc_torusPolynomialMulByXaiMinusOne(result + q * N, barai, accum + q * N, N);
    orc_tGswFFTExternMulToTLwe(result, bki, N, l, Bgbit);
    for (int j = 0; j < 2 * N; j++) result[j] = (Torus32)((uint32_t)result[j] + (uint32_t)accum[j]);
}
/* cb/lwe_functions.cpp:337-361 */
void orc_tfhe_blindRotate_FFT(Torus32* accum, const double* bkFFT, const int32_t* bara, int n, int N, int l, int Bgbit) {
    Torus32 temp[2 * N];
    const size_t stride = (size_t)2 * l * 2 * N;
    for (int i = 0; i < n; i++) {
        const int barai = bara[i];
        if (barai == 0) continue;
        orc_tfhe_MuxRotate_FFT(temp, accum, bkFFT + i * stride, barai, N, l, Bgbit);
        memcpy(accum, temp, sizeof(temp));
    }
}
/* cb/tlwe_functions.cpp:351-367 (index 0) */
void orc_tLweExtractLweSample(Torus32* result, const Torus32* tlwe, int N) {
    result[0] = tlwe[0];
    for (int j = 1; j < N; j++) result[j] = (Torus32)(0u - (uint32_t)tlwe[N - j]);
    result[N] = tlwe[N + 0];
}
/* cb/lwe_functions.cpp:366-395 */
void orc_tfhe_blindRotateAndExtract_FFT(Torus32* result, const Torus32* v, const double* bkFFT, int barb,
                                        const int32_t* bara, int n, int N, int l, int Bgbit) {
    Torus32 acc[2 * N];
    memset(acc, 0, sizeof(Torus32) * N);
    if (barb != 0) orc_torusPolynomialMulByXai(acc + N, 2 * N - barb, v, N);
    else memcpy(acc + N, v, sizeof(Torus32) * N);
    orc_tfhe_blindRotate_FFT(acc, bkFFT, bara, n, N, l, Bgbit);
    orc_tLweExtractLweSample(result, acc, N);
}
/* cb/lwe_functions.cpp:399-430 */
void orc_tfhe_bootstrap_woKS_FFT(Torus32* result, const orc_gate_keys* K, Torus32 mu, const Torus32* x) {
    const int n = K->p.n, N = K->p.N;
    Torus32 testvect[N];
    int32_t bara[n];
    int barb = orc_modSwitchFromTorus32(x[n], 2 * N);
    for (int i = 0; i < n; i++) bara[i] = orc_modSwitchFromTorus32(x[i], 2 * N);
    for (int i = 0; i < N; i++) testvect[i] = mu;
    orc_tfhe_blindRotateAndExtract_FFT(result, testvect, K->bkFFT, barb, bara, n, N, K->p.bk_l, K->p.bk_Bgbit);
}
/* cb/lwe_functions.cpp:136-171 ; identical arithmetic in cb/poc_CircuitBootstrapping.cpp:437-465 */
void orc_lweKeySwitch(Torus32* result, const Torus32* ks, const Torus32* sample, int n_in, int n_out, int t, int basebit) {
    const int base = 1 << basebit;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    const uint32_t mask = (uint32_t)base - 1;
    uint32_t* res = (uint32_t*)result;
    for (int h = 0; h < n_out; h++) res[h] = 0;
    res[n_out] = (uint32_t)sample[n_in];
    for (int i = 0; i < n_in; i++) {
        const uint32_t aibar = (uint32_t)sample[i] + prec_offset;
        for (int j = 0; j < t; j++) {
            const uint32_t aij = (aibar >> (32 - (j + 1) * basebit)) & mask;
            if (aij != 0) {
                const uint32_t* row = (const uint32_t*)ks + (((size_t)i * t + j) * base + aij) * (n_out + 1);
                for (int h = 0; h <= n_out; h++) res[h] -= row[h];
            }
        }
    }
}
/* cb/lwe_functions.cpp:434-446 */
void orc_tfhe_bootstrap_FFT(Torus32* result, const orc_gate_keys* K, Torus32 mu, const Torus32* x) {
    Torus32 u[K->p.N + 1];
    orc_tfhe_bootstrap_woKS_FFT(u, K, mu, x);
    orc_lweKeySwitch(result, K->ks, u, K->p.N, K->p.n, K->p.ks_t, K->p.ks_basebit);
}

/* boots* [UPSTREAM, SURVEY Appendix C]: tmp = (0,c) + ka*ca + kb*cb ; bootstrap with MU = 1/8 */
static const struct { int c8; int ka; int kb; } g_gate[ORC_NUM_GATES] = {
    /* NAND */ { 1, -1, -1}, /* AND */ {-1, 1, 1}, /* OR */ { 1, 1, 1}, /* NOR */ {-1, -1, -1},
    /* XOR  */ { 2,  2,  2}, /* XNOR*/ {-2, -2, -2},
    /* ANDNY*/ {-1, -1,  1}, /* ANDYN*/ {-1, 1, -1}, /* ORNY */ { 1, -1, 1}, /* ORYN */ { 1, 1, -1},
};
void orc_gate_lincomb(Torus32* tmp, int op, const Torus32* ca, const Torus32* cb, int n) {
    const uint32_t ka = (uint32_t)g_gate[op].ka, kb = (uint32_t)g_gate[op].kb;
    for (int i = 0; i <= n; i++) tmp[i] = (Torus32)(ka * (uint32_t)ca[i] + kb * (uint32_t)cb[i]);
    tmp[n] = (Torus32)((uint32_t)tmp[n] + (uint32_t)orc_modSwitchToTorus32(g_gate[op].c8, 8));
}
void orc_bootsGate(Torus32* result, int op, const Torus32* ca, const Torus32* cb, const orc_gate_keys* K) {
    Torus32 tmp[K->p.n + 1];
    orc_gate_lincomb(tmp, op, ca, cb, K->p.n);
    orc_tfhe_bootstrap_FFT(result, K, orc_modSwitchToTorus32(1, 8), tmp);
}
void orc_bootsNOT(Torus32* result, const Torus32* ca, int n) {
    for (int i = 0; i <= n; i++) result[i] = (Torus32)(0u - (uint32_t)ca[i]);
}
void orc_bootsMUX(Torus32* result, const Torus32* a, const Torus32* b, const Torus32* c, const orc_gate_keys* K) {
    const int n = K->p.n, N = K->p.N;
    const Torus32 MU = orc_modSwitchToTorus32(1, 8);
    Torus32 t1[n + 1], t2[n + 1], u1[N + 1], u2[N + 1];
    orc_gate_lincomb(t1, ORC_AND, a, b, n);       /* AND(a,b)      */
    orc_gate_lincomb(t2, ORC_ANDNY, a, c, n);     /* AND(not a, c) */
    orc_tfhe_bootstrap_woKS_FFT(u1, K, MU, t1);
    orc_tfhe_bootstrap_woKS_FFT(u2, K, MU, t2);
    for (int i = 0; i <= N; i++) u1[i] = (Torus32)((uint32_t)u1[i] + (uint32_t)u2[i]);
    u1[N] = (Torus32)((uint32_t)u1[N] + (uint32_t)MU);
    orc_lweKeySwitch(result, K->ks, u1, N, n, K->p.ks_t, K->p.ks_basebit);
}
void orc_bootsSymEncrypt(Torus32* result, int message, const orc_gate_keys* K, orc_rng* r) {
    Torus32 MU = orc_modSwitchToTorus32(1, 8);
    orc_lweSymEncrypt(result, message ? MU : -MU, K->p.ks_stdev, K->lwe_key, K->p.n, r);
}
int orc_bootsSymDecrypt(const Torus32* sample, const orc_gate_keys* K) {
    return orc_lwePhase(sample, K->lwe_key, K->p.n) > 0;
}
int orc_gate_plain(int op, int a, int b) {
    switch (op) {
        case ORC_NAND: return !(a && b); case ORC_AND: return a && b; case ORC_OR: return a || b;
        case ORC_NOR: return !(a || b); case ORC_XOR: return a ^ b; case ORC_XNOR: return !(a ^ b);
        case ORC_ANDNY: return (!a) && b; case ORC_ANDYN: return a && !b;
        case ORC_ORNY: return (!a) || b; case ORC_ORYN: return a || !b;
    }
    return -1;
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_tfhe_bootstrap_woKS_FFT_batch(Torus32* result, const orc_gate_keys* K, Torus32 mu, const Torus32* x, int count, int threads) {
    const size_t si = (size_t)K->p.n + 1, so = (size_t)K->p.N + 1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int g = 0; g < count; g++) orc_tfhe_bootstrap_woKS_FFT(result + g * so, K, mu, x + g * si);
}
void orc_bootsGate_batch(Torus32* result, int op, const Torus32* ca, const Torus32* cb, const orc_gate_keys* K, int count, int threads) {
    const size_t s = (size_t)K->p.n + 1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int g = 0; g < count; g++) orc_bootsGate(result + g * s, op, ca + g * s, cb + g * s, K);
}

/* ================================================================== circuit bootstrapping */
void orc_cb_params_default(orc_cb_params* p) {      /* cb/poc_CircuitBootstrapping.cpp:70-85 */
    p->n_lvl0 = 500; p->N_lvl1 = 1024; p->N_lvl2 = 2048;
    p->bgbit_lvl1 = 8; p->ell_lvl1 = 2; p->bgbit_lvl2 = 9; p->ell_lvl2 = 4;
    p->bkstdev_lvl2 = ldexp(1.0, -44); p->ksstdev_lvl10 = ldexp(1.0, -14);
    p->kslength_lvl10 = 6; p->ksbasebit_lvl10 = 2;
    p->ksstdev_lvl21 = ldexp(1.0, -31); p->kslength_lvl21 = 10; p->ksbasebit_lvl21 = 3;
}
/* cb/poc_CircuitBootstrapping.cpp:191-200 */
static void tlwe64_encrypt(Torus64* c /*[2][N]*/, Torus64 mess, double stdev, const double* keyFFT, int N, orc_rng* r) {
    Torus64* a = c; Torus64* b = c + N;
    b[0] = orc_gaussian64(r, mess, stdev);
    for (int j = 1; j < N; j++) b[j] = orc_gaussian64(r, 0, stdev);
    for (int j = 0; j < N; j++) a[j] = (Torus64)orc_rng_u64(r);
    orc_torus64PolynomialMultAddBinKey(b, keyFFT, a, N);
}
orc_cb_keys* orc_cb_keygen(const orc_cb_params* p, uint64_t seed, int with_privks) {
    orc_cb_keys* K = xmalloc(sizeof(*K));
    K->p = *p;
    const int n0 = p->n_lvl0, N1 = p->N_lvl1, N2 = p->N_lvl2, l2 = p->ell_lvl2;
    orc_rng r; orc_rng_seed(&r, seed);
    /* secret keys  :357-369 */
    K->key_lvl0 = xmalloc(sizeof(int32_t) * n0);
    K->key_lvl1 = xmalloc(sizeof(int32_t) * N1);
    K->key_lvl2 = xmalloc(sizeof(int32_t) * (N2 + 1));
    for (int i = 0; i < n0; i++) K->key_lvl0[i] = orc_rng_bit(&r);
    for (int i = 0; i < N1; i++) K->key_lvl1[i] = orc_rng_bit(&r);
    for (int i = 0; i < N2; i++) K->key_lvl2[i] = orc_rng_bit(&r);
    K->key_lvl2[N2] = -1;
    /* preKS :372-383 */
    const int t10 = p->kslength_lvl10, bb10 = p->ksbasebit_lvl10, base10 = 1 << bb10;
    K->preKS = xmalloc(sizeof(Torus32) * (size_t)N1 * t10 * base10 * (n0 + 1));
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < t10; j++)
            for (int u = 0; u < base10; u++) {
                Torus32 mess = (Torus32)(((uint32_t)K->key_lvl1[i] << (32 - (j + 1) * bb10)) * (uint32_t)u);
                orc_lweSymEncrypt(K->preKS + (((size_t)i * t10 + j) * base10 + u) * (n0 + 1), mess, p->ksstdev_lvl10, K->key_lvl0, n0, &r);
            }
    /* bk :388-391 with tGsw64Encrypt_lvl2 :215-227 */
    double* key2FFT = key_to_fft(K->key_lvl2, N2);
    K->bk = xmalloc(sizeof(Torus64) * (size_t)n0 * 2 * l2 * 2 * N2);
    for (int i = 0; i < n0; i++)
        for (int bloc = 0; bloc <= K1; bloc++)
            for (int j = 0; j < l2; j++) {
                Torus64* row = K->bk + (((size_t)i * 2 * l2 + bloc * l2 + j) * 2) * N2;
                tlwe64_encrypt(row, 0, p->bkstdev_lvl2, key2FFT, N2, &r);
                row[bloc * N2] = (Torus64)((uint64_t)row[bloc * N2] + (uint64_t)(int64_t)K->key_lvl0[i] * (UINT64_C(1) << (64 - (j + 1) * p->bgbit_lvl2)));
            }
    free(key2FFT);
    K->bkFFT = NULL;
    orc_cb_keys_rebuild_fft(K);
    /* privKS :405-419 */
    K->privKS = NULL;
    if (with_privks) {
        const int t21 = p->kslength_lvl21, bb21 = p->ksbasebit_lvl21, base21 = 1 << bb21;
        double* key1FFT = key_to_fft(K->key_lvl1, N1);
        size_t rows = (size_t)2 * (N2 + 1) * t21 * base21;
        K->privKS = xmalloc(sizeof(Torus32) * rows * 2 * N1);
        /* one RNG stream per row so the loop can run under OpenMP and stay deterministic */
        uint64_t base_seed = orc_rng_u64(&r);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
        for (long long row = 0; row < (long long)rows; row++) {
            int u = (int)(row % base21);
            int j = (int)((row / base21) % t21);
            int i = (int)((row / ((size_t)base21 * t21)) % (N2 + 1));
            int z = (int)(row / ((size_t)base21 * t21 * (N2 + 1)));
            orc_rng rr; orc_rng_seed(&rr, base_seed + 0x632BE59BD9B4E019ull * (uint64_t)(row + 1));
            Torus32* c = K->privKS + (size_t)row * 2 * N1;
            Torus32 mess = (Torus32)(((uint32_t)K->key_lvl2[i] << (32 - (j + 1) * bb21)) * (uint32_t)u);
            tlwe32_encrypt_zero(c, p->ksstdev_lvl21, key1FFT, N1, &rr);
            c[z * N1] = (Torus32)((uint32_t)c[z * N1] + (uint32_t)mess);
        }
        free(key1FFT);
    }
    return K;
}
/* :394-402 */
void orc_cb_keys_rebuild_fft(orc_cb_keys* K) {
    const size_t total = (size_t)K->p.n_lvl0 * 2 * K->p.ell_lvl2 * 2;
    const int N = K->p.N_lvl2;
    if (!K->bkFFT) K->bkFFT = xmalloc(sizeof(double) * total * N);
    for (size_t q = 0; q < total; q++) g_backend->ifft_torus64(N, K->bkFFT + q * N, K->bk + q * N);
}
void orc_cb_keys_free(orc_cb_keys* K) {
    if (!K) return;
    free(K->key_lvl0); free(K->key_lvl1); free(K->key_lvl2); free(K->preKS); free(K->bk); free(K->bkFFT); free(K->privKS); free(K);
}

/* :349-350 */
uint64_t orc_tgsw64_offset(int l, int Bgbit) {
    uint64_t off = 0;
    for (int i = 0; i <= l; i++) off |= UINT64_C(1) << (63 - i * Bgbit);
    return off;
}
/* :492-515 */
void orc_tGswTorus64PolynomialDecompH(int32_t* result, const Torus64* sample, int N, int l, int Bgbit) {
    const uint64_t mask = (UINT64_C(1) << Bgbit) - 1;
    const int32_t halfBg = (int32_t)((UINT64_C(1) << Bgbit) / 2);
    const uint64_t offset = orc_tgsw64_offset(l, Bgbit);
    for (int p = 0; p < l; p++) {
        const int decal = 64 - (p + 1) * Bgbit;
        for (int j = 0; j < N; j++) {
            uint32_t temp1 = (uint32_t)((((uint64_t)sample[j] + offset) >> decal) & mask);
            result[p * N + j] = (int32_t)(temp1 - (uint32_t)halfBg);
        }
    }
}
/* :437-465 */
void orc_preKeySwitch(Torus32* result, const Torus32* x, const orc_cb_keys* K) {
    orc_lweKeySwitch(result, K->preKS, x, K->p.N_lvl1, K->p.n_lvl0, K->p.kslength_lvl10, K->p.ksbasebit_lvl10);
}
/* :472-484 */
void orc_preModSwitch(int32_t* result, const Torus32* x, int n0, int N2) {
    for (int i = 0; i <= n0; i++) result[i] = orc_modSwitchFromTorus32(x[i], 2 * N2);
}
/* :609-620 (decompose, 2l ifft, clear, 2l*(k+1) AddMul, (k+1) fft) */
void orc_tGsw64FFTExternMulToTLwe(Torus64* accum, const double* gswFFT, int N, int l, int Bgbit) {
    const int kpl = 2 * l;
    int32_t* decomp = xmalloc(sizeof(int32_t) * kpl * N);
    double* decompFFT = xmalloc(sizeof(double) * kpl * N);
    double* accFFT = xmalloc(sizeof(double) * 2 * N);
    for (int i = 0; i <= K1; i++) orc_tGswTorus64PolynomialDecompH(decomp + i * l * N, accum + i * N, N, l, Bgbit);
    for (int p = 0; p < kpl; p++) g_backend->ifft_int(N, decompFFT + p * N, decomp + p * N);
    memset(accFFT, 0, sizeof(double) * 2 * N);
    for (int p = 0; p < kpl; p++)
        for (int q = 0; q <= K1; q++)
            g_backend->addmul(N, accFFT + q * N, decompFFT + p * N, gswFFT + ((size_t)p * 2 + q) * N);
    for (int q = 0; q <= K1; q++) g_backend->fft_torus64(N, accum + q * N, accFFT + q * N);
    free(decomp); free(decompFFT); free(accFFT);
}
/* the same external product with exact integer products: the reference's non-USE_FFT build ("fake FFT",
 * cb/poc_CircuitBootstrapping.cpp:285-316) routes AddMul through torus64PolynomialMultAddKaratsuba_lvl2 */
void orc_tGsw64ExternMulToTLwe_exact(Torus64* accum, const Torus64* gsw, int N, int l, int Bgbit) {
    const int kpl = 2 * l;
    int32_t* decomp = xmalloc(sizeof(int32_t) * kpl * N);
    for (int i = 0; i <= K1; i++) orc_tGswTorus64PolynomialDecompH(decomp + i * l * N, accum + i * N, N, l, Bgbit);
    memset(accum, 0, sizeof(Torus64) * 2 * N);
    for (int p = 0; p < kpl; p++)
        for (int q = 0; q <= K1; q++)
            orc_torus64PolynomialMultAddNaive(accum + q * N, decomp + p * N, gsw + ((size_t)p * 2 + q) * N, N);
    free(decomp);
}
/* :530-659 with corrections D1 (bkFFT[i]), D2 (signs/indices of (X^a-1)), D3 (rotate test vector by 2N-bbar) */
void orc_circuitBootstrapWoKS(Torus64* result, Torus64 mu, const int32_t* abar, const orc_cb_keys* K) {
    const int N = K->p.N_lvl2, n0 = K->p.n_lvl0, l = K->p.ell_lvl2, N2 = N / 2;
    const Torus64 mu2 = mu / 2;
    Torus64* tv = xmalloc(sizeof(Torus64) * N);
    Torus64* acc = xmalloc(sizeof(Torus64) * 2 * N);
    Torus64* acc2 = xmalloc(sizeof(Torus64) * 2 * N);
    const int bbar = (2 * N - abar[n0]) % (2 * N);                 /* D3 */
    for (int j = 0; j < N2; j++) tv[j] = -mu2;                     /* :552-553 */
    for (int j = N2; j < N; j++) tv[j] = mu2;
    memset(acc, 0, sizeof(Torus64) * N);                           /* :565-568 */
    orc_torus64PolynomialMulByXai(acc + N, bbar, tv, N);           /* :555-562 */
    const size_t stride = (size_t)2 * l * 2 * N;
    for (int i = 0; i < n0; i++) {                                 /* :580-642 */
        const int aibar = abar[i];
        if (aibar == 0) continue;
        for (int q = 0; q <= K1; q++) orc_torus64PolynomialMulByXaiMinusOne(acc2 + q * N, aibar, acc + q * N, N);   /* D2 */
        orc_tGsw64FFTExternMulToTLwe(acc2, K->bkFFT + i * stride, N, l, K->p.bgbit_lvl2);                          /* D1 */
        for (int j = 0; j < 2 * N; j++) acc[j] = (Torus64)((uint64_t)acc[j] + (uint64_t)acc2[j]);                   /* :631-632 */
    }
    result[0] = acc[0];                                            /* :646-648 */
    for (int j = 1; j < N; j++) result[j] = (Torus64)(0ull - (uint64_t)acc[N - j]);
    result[N] = (Torus64)((uint64_t)acc[N] + (uint64_t)mu2);
    free(tv); free(acc); free(acc2);
}
/* :667-698 */
void orc_circuitPrivKS(Torus32* result, int u, const Torus64* x, const orc_cb_keys* K) {
    const int n2 = K->p.N_lvl2, N1 = K->p.N_lvl1, kslen = K->p.kslength_lvl21, bb = K->p.ksbasebit_lvl21;
    const int base = 1 << bb;
    const uint64_t mask = (uint64_t)base - 1;
    const uint64_t prec_offset = UINT64_C(1) << (64 - (1 + bb * kslen));
    uint32_t* res = (uint32_t*)result;
    memset(res, 0, sizeof(uint32_t) * 2 * N1);
    for (int i = 0; i <= n2; i++) {
        const uint64_t aibar = (uint64_t)x[i] + prec_offset;
        for (int j = 0; j < kslen; j++) {
            const uint64_t aij = (aibar >> (64 - (j + 1) * bb)) & mask;
            if (aij != 0) {
                const uint32_t* row = (const uint32_t*)K->privKS + ((((size_t)u * (n2 + 1) + i) * kslen + j) * base + aij) * 2 * N1;
                for (int p = 0; p < 2 * N1; p++) res[p] -= row[p];
            }
        }
    }
}
/* :823-873 ; result laid out as samples[u][w] -> result[((u*l1)+w)][2][N1] */
void orc_tfhe_CircuitBootstrapFFT(Torus32* result, const Torus32* sample, const orc_cb_keys* K) {
    const int n0 = K->p.n_lvl0, N1 = K->p.N_lvl1, N2 = K->p.N_lvl2, ell1 = K->p.ell_lvl1;
    Torus32 res_preKS[n0 + 1];
    int32_t res_preMS[n0 + 1];
    Torus64* res_boot = xmalloc(sizeof(Torus64) * (N2 + 1));
    orc_preKeySwitch(res_preKS, sample, K);
    orc_preModSwitch(res_preMS, res_preKS, n0, N2);
    for (int w = 0; w < ell1; w++) {
        const Torus64 mu1 = (Torus64)(UINT64_C(1) << (64 - (w + 1) * K->p.bgbit_lvl1));
        orc_circuitBootstrapWoKS(res_boot, mu1, res_preMS, K);
        for (int u = 0; u <= K1; u++) orc_circuitPrivKS(result + ((size_t)(u * ell1 + w) * 2) * N1, u, res_boot, K);
    }
    free(res_boot);
}
/* :98-106 */
void orc_lwe32Encrypt_lvl1(Torus32* cipher, Torus32 mess, double stdev, const orc_cb_keys* K, orc_rng* r) {
    orc_lweSymEncrypt(cipher, mess, stdev, K->key_lvl1, K->p.N_lvl1, r);
}
/* :127-134 */
Torus64 orc_lwe64Phase_lvl2(const Torus64* cipher, const orc_cb_keys* K) {
    const int n = K->p.N_lvl2;
    uint64_t res = (uint64_t)cipher[n];
    for (int i = 0; i < n; i++) res -= (uint64_t)cipher[i] * (uint64_t)(int64_t)K->key_lvl2[i];
    return (Torus64)res;
}
/* :155-171 : phase = b - a*K */
void orc_tLwe32Phase_lvl1(Torus32* phase, const Torus32* cipher, const orc_cb_keys* K) {
    const int N = K->p.N_lvl1;
    Torus32 t[N];
    memset(t, 0, sizeof(t));
    orc_torus32PolynomialMultAddNaive(t, K->key_lvl1, cipher, N);
    for (int j = 0; j < N; j++) phase[j] = (Torus32)((uint32_t)cipher[N + j] - (uint32_t)t[j]);
}
