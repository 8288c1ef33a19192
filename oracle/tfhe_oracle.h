/*
 * tfhe_oracle.h -- CPU restatement of the experimental-tfhe bootstrapping hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call into it.  The shipped
 * library (experimental-tfhe_b200/csrc) never links or loads anything from oracle/.
 *
 * Every function cites the reference file:line it restates.  Path abbreviations:
 *   cb/  = /root/reference/circuit-bootstrapping/src/
 *   hp/  = /root/reference/high-precision-anticyclic-fft/src/
 *
 * Parity status: pinned against the reference compiled in place (oracle/_ref, see
 * oracle/Makefile + oracle/ref_harness.cpp); golden vectors under tests/golden/ are
 * produced by that build (tests/golden/make_golden.py).  The gate-level (boots*) layer
 * is NOT in the reference (upstream tfhe/tfhe, unversioned): gate truth tables are pinned
 * by decryption, not by reference vectors.
 *
 * Flat layouts (shared with include/tfhe_b200.h):
 *   LWE sample (n)      : torus[n+1], b = [n]                  cb/poc_types.h:137-158
 *   TLWE sample (N,k=1) : torus[2][N], b = poly 1              cb/poc_types.h:164-184
 *   TGSW sample         : torus[2*l][2][N], row p = bloc*l+i   cb/poc_types.h:206-234
 *   bootstrapping key   : TGSW[n]
 *   LWE key-switch key  : torus32[N_in][t][base][n_out+1]      cb/lwe_functions.cpp:96-110
 *   private KS key      : torus32[2][n_in+1][t][base][2][N_out] cb/poc_CircuitBootstrapping.cpp:408
 *   LagrangeHalfC       : double[N], re = [0,N/2), im = [N/2,N) cb/poc_types.h:96-102
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t Torus32;
typedef int64_t Torus64;

/* ------------------------------------------------------------------ RNG */
typedef struct { uint64_t s; int has_spare; double spare; } orc_rng;
void     orc_rng_seed(orc_rng* r, uint64_t seed);
uint64_t orc_rng_u64(orc_rng* r);
int      orc_rng_bit(orc_rng* r);
double   orc_rng_normal(orc_rng* r);                       /* N(0,1) */
Torus32  orc_gaussian32(orc_rng* r, Torus32 center, double stdev); /* cb/generic_utils.h:176-181 */
Torus64  orc_gaussian64(orc_rng* r, Torus64 center, double stdev); /* cb/generic_utils.h:183-189 */

/* ------------------------------------------------------------------ FFT backend
 * The half-complex negacyclic transform.  The default backend is a portable
 * radix-2 restatement of cb/spqlios/spqlios-fft-impl.cpp:204-397,469-641; the reference
 * harness swaps in the reference's own spqlios assembly through orc_set_fft_backend. */
typedef struct {
    void (*ifft_int)(int N, double* res, const int32_t* a);      /* execute_reverse_int     cb/spqlios/fft_processor_spqlios.cpp:27-67 */
    void (*ifft_torus64)(int N, double* res, const int64_t* a);  /* execute_reverse_torus64 :166-170 */
    void (*fft_torus32)(int N, int32_t* res, const double* a);   /* execute_direct_torus32  :77-103 */
    void (*fft_torus64)(int N, int64_t* res, const double* a);   /* execute_direct_torus64  :105-156 */
    void (*addmul)(int N, double* res, const double* a, const double* b); /* LagrangeHalfCPolynomialAddMulASM lagrangehalfc_impl_fma.s:78-135 */
} orc_fft_backend;
void orc_set_fft_backend(const orc_fft_backend* b);   /* NULL restores the portable one */
const orc_fft_backend* orc_get_fft_backend(void);
/* raw portable transforms (double in/out, no conversion): fft(ifft(x)) == (N/2) x */
void orc_ifft_raw(int N, double* data);                /* twist + DIF, in place  */
void orc_fft_raw(int N, double* data);                 /* DIT + untwist, in place, unscaled */
/* double -> torus conversions, SURVEY A.8 */
Torus32 orc_double_to_torus32(double x);               /* int32(int64(x))        fft_processor_spqlios.cpp:102 */
Torus64 orc_double_to_torus64(double x);               /* mantissa shift         fft_processor_spqlios.cpp:131-142 */

/* ------------------------------------------------------------------ exact products (Karatsuba stand-in)
 * result (+)= poly1 * poly2 mod X^N+1, wrap-around integer arithmetic.
 * cb/poc_karatsuba.cpp:80-99,188-206 ; naive form par/poc_karatsuba.cpp:10-21. */
void orc_torus32PolynomialMultAddNaive(Torus32* result, const int32_t* poly1, const Torus32* poly2, int N);
void orc_torus64PolynomialMultAddNaive(Torus64* result, const int32_t* poly1, const Torus64* poly2, int N);
/* same product through the double FFT, exact when |poly1| <= 1 (binary keys); used by keygen only */
void orc_torus32PolynomialMultAddBinKey(Torus32* result, const double* keyFFT, const Torus32* poly2, int N);
void orc_torus64PolynomialMultAddBinKey(Torus64* result, const double* keyFFT, const Torus64* poly2, int N);

/* ------------------------------------------------------------------ numeric (cb/numeric_functions.cpp) */
int     orc_modSwitchFromTorus32(Torus32 phase, int Msize);   /* :54-60 */
Torus32 orc_modSwitchToTorus32(int mu, int Msize);            /* :62-67 */
void    orc_torusPolynomialMulByXaiMinusOne(Torus32* out, int a, const Torus32* in, int N); /* :304-323 */
void    orc_torusPolynomialMulByXai(Torus32* out, int a, const Torus32* in, int N);         /* :327-347 */
void    orc_torus64PolynomialMulByXaiMinusOne(Torus64* out, int a, const Torus64* in, int N);
void    orc_torus64PolynomialMulByXai(Torus64* out, int a, const Torus64* in, int N);

/* ------------------------------------------------------------------ gate path (Torus32) */
typedef struct {
    int n;          /* LWE dimension */
    int N;          /* ring degree   */
    int k;          /* always 1 (cb/poc_types.h:10) */
    int bk_l;       /* gadget length */
    int bk_Bgbit;   /* log2 gadget base */
    int ks_t;       /* key-switch length */
    int ks_basebit; /* key-switch log2 base */
    double bk_stdev;
    double ks_stdev;
} orc_gate_params;
/* P_gate defaults: n=500 N=1024 k=1 l=2 Bgbit=10 (misc/params-gb.html:124-131), KS t=8 basebit=2 [UPSTREAM] */
void orc_gate_params_default(orc_gate_params* p);

typedef struct {
    orc_gate_params p;
    int32_t* lwe_key;      /* [n]  binary */
    int32_t* tlwe_key;     /* [N]  binary */
    Torus32* bk;           /* [n][2l][2][N] coefficient domain */
    double*  bkFFT;        /* [n][2l][2][N] LagrangeHalfC (backend layout) */
    Torus32* ks;           /* [N][t][base][n+1] */
} orc_gate_keys;
orc_gate_keys* orc_gate_keygen(const orc_gate_params* p, uint64_t seed);
void orc_gate_keys_rebuild_fft(orc_gate_keys* K);   /* recompute bkFFT with the current backend */
void orc_gate_keys_free(orc_gate_keys* K);

void    orc_lweSymEncrypt(Torus32* result, Torus32 message, double alpha, const int32_t* key, int n, orc_rng* r); /* cb/lwe_functions.cpp:43-54 */
Torus32 orc_lwePhase(const Torus32* sample, const int32_t* key, int n);                                         /* :56-65 */

uint32_t orc_tgsw32_offset(int l, int Bgbit);                                                  /* cb/tgsw_functions.cpp:30-36 */
void orc_tGswTorus32PolynomialDecompH(int32_t* result /*[l][N]*/, const Torus32* sample, int N, int l, int Bgbit); /* :224-337 */
void orc_tGswFFTExternMulToTLwe(Torus32* accum /*[2][N]*/, const double* gswFFT /*[2l][2][N]*/, int N, int l, int Bgbit); /* :424-449 */
void orc_tGswExternMulToTLwe(Torus32* accum /*[2][N]*/, const Torus32* gsw /*[2l][2][N] coefficients*/, int N, int l, int Bgbit); /* exact, :150-164 */
void orc_tfhe_MuxRotate_FFT(Torus32* result, const Torus32* accum, const double* bki, int barai, int N, int l, int Bgbit); /* cb/lwe_functions.cpp:328-333 */
void orc_tfhe_blindRotate_FFT(Torus32* accum, const double* bkFFT, const int32_t* bara, int n, int N, int l, int Bgbit);   /* :337-361 */
void orc_tfhe_blindRotateAndExtract_FFT(Torus32* result /*[N+1]*/, const Torus32* v, const double* bkFFT, int barb,
                                        const int32_t* bara, int n, int N, int l, int Bgbit);                         /* :366-395 */
void orc_tfhe_bootstrap_woKS_FFT(Torus32* result /*[N+1]*/, const orc_gate_keys* K, Torus32 mu, const Torus32* x);        /* :399-430 */
void orc_lweKeySwitch(Torus32* result /*[n_out+1]*/, const Torus32* ks, const Torus32* sample /*[n_in+1]*/,
                      int n_in, int n_out, int t, int basebit);                                                        /* :136-171 */
void orc_tfhe_bootstrap_FFT(Torus32* result /*[n+1]*/, const orc_gate_keys* K, Torus32 mu, const Torus32* x);             /* :434-446 */
void orc_tLweExtractLweSample(Torus32* result /*[N+1]*/, const Torus32* tlwe /*[2][N]*/, int N);                          /* cb/tlwe_functions.cpp:351-367 */

/* boots* gates [UPSTREAM semantics, SURVEY Appendix C].  op codes shared with include/tfhe_b200.h */
enum { ORC_NAND = 0, ORC_AND, ORC_OR, ORC_NOR, ORC_XOR, ORC_XNOR, ORC_ANDNY, ORC_ANDYN, ORC_ORNY, ORC_ORYN, ORC_NUM_GATES };
void orc_gate_lincomb(Torus32* tmp /*[n+1]*/, int op, const Torus32* ca, const Torus32* cb, int n);
void orc_bootsGate(Torus32* result, int op, const Torus32* ca, const Torus32* cb, const orc_gate_keys* K);
void orc_bootsNOT(Torus32* result, const Torus32* ca, int n);
void orc_bootsMUX(Torus32* result, const Torus32* a, const Torus32* b, const Torus32* c, const orc_gate_keys* K);
void orc_bootsSymEncrypt(Torus32* result, int message, const orc_gate_keys* K, orc_rng* r);
int  orc_bootsSymDecrypt(const Torus32* sample, const orc_gate_keys* K);
int  orc_gate_plain(int op, int a, int b);
/* OpenMP batch (the reference's only threading idiom, par/test_parallel_multiplications.cpp:62) */
void orc_tfhe_bootstrap_woKS_FFT_batch(Torus32* result, const orc_gate_keys* K, Torus32 mu, const Torus32* x, int count, int threads);
void orc_bootsGate_batch(Torus32* result, int op, const Torus32* ca, const Torus32* cb, const orc_gate_keys* K, int count, int threads);
int  orc_num_threads(void);

/* ------------------------------------------------------------------ circuit bootstrapping (cb/poc_CircuitBootstrapping.cpp) */
typedef struct {
    int n_lvl0, N_lvl1, N_lvl2;         /* :71-73 */
    int bgbit_lvl1, ell_lvl1;           /* :74-75 */
    int bgbit_lvl2, ell_lvl2;           /* :76-77 */
    int kslength_lvl10, ksbasebit_lvl10;/* :80-81 */
    int kslength_lvl21, ksbasebit_lvl21;/* :83-84 */
    double bkstdev_lvl2, ksstdev_lvl10, ksstdev_lvl21; /* :78,79,82 */
} orc_cb_params;
void orc_cb_params_default(orc_cb_params* p);     /* the active "#if 1" block :70-85 */

typedef struct {
    orc_cb_params p;
    int32_t* key_lvl0;     /* [n0] */
    int32_t* key_lvl1;     /* [N1] */
    int32_t* key_lvl2;     /* [N2+1], last = -1  (:365-367) */
    Torus32* preKS;        /* [N1][t10][base10][n0+1] (:375) */
    Torus64* bk;           /* [n0][2*l2][2][N2] */
    double*  bkFFT;        /* [n0][2*l2][2][N2] */
    Torus32* privKS;       /* [2][N2+1][t21][base21][2][N1] (:408), may be NULL */
} orc_cb_keys;
/* with_privks=0 skips the 2.69 GB private key-switch key */
orc_cb_keys* orc_cb_keygen(const orc_cb_params* p, uint64_t seed, int with_privks);
void orc_cb_keys_rebuild_fft(orc_cb_keys* K);
void orc_cb_keys_free(orc_cb_keys* K);

uint64_t orc_tgsw64_offset(int l, int Bgbit);                                        /* :349-350 */
void orc_tGswTorus64PolynomialDecompH(int32_t* result /*[l][N]*/, const Torus64* sample, int N, int l, int Bgbit); /* :492-515 */
void orc_preKeySwitch(Torus32* result /*[n0+1]*/, const Torus32* x /*[N1+1]*/, const orc_cb_keys* K);  /* :437-465 */
void orc_preModSwitch(int32_t* result /*[n0+1]*/, const Torus32* x /*[n0+1]*/, int n0, int N2);          /* :472-484 */
/* blind rotation + extract with the three corrections of SURVEY Appendix B (D1,D2,D3) */
void orc_circuitBootstrapWoKS(Torus64* result /*[N2+1]*/, Torus64 mu, const int32_t* abar, const orc_cb_keys* K); /* :530-659 */
void orc_tGsw64FFTExternMulToTLwe(Torus64* accum /*[2][N]*/, const double* gswFFT, int N, int l, int Bgbit);      /* :609-620 */
void orc_tGsw64ExternMulToTLwe_exact(Torus64* accum /*[2][N]*/, const Torus64* gsw /*coefficients*/, int N, int l, int Bgbit); /* :285-316 */
void orc_circuitPrivKS(Torus32* result /*[2][N1]*/, int u, const Torus64* x /*[N2+1]*/, const orc_cb_keys* K);   /* :667-698 */
void orc_tfhe_CircuitBootstrapFFT(Torus32* result /*[2][l1][2][N1] = samples[u][w]*/, const Torus32* sample /*[N1+1]*/,
                                  const orc_cb_keys* K);                                                        /* :823-873 */
void    orc_lwe32Encrypt_lvl1(Torus32* cipher, Torus32 mess, double stdev, const orc_cb_keys* K, orc_rng* r);   /* :98-106 */
Torus64 orc_lwe64Phase_lvl2(const Torus64* cipher, const orc_cb_keys* K);                                       /* :127-134 */
void    orc_tLwe32Phase_lvl1(Torus32* phase /*[N1]*/, const Torus32* cipher /*[2][N1]*/, const orc_cb_keys* K); /* :155-171 */

/* ------------------------------------------------------------------ high-precision FFT (hp/code.cpp) */
typedef unsigned __int128 orc_u128;
typedef struct { orc_u128 re, im; } orc_cplx96;           /* complex<Real96>  hp/code.cpp:25-75,212 */
orc_u128 orc_real96_mul(orc_u128 a, orc_u128 b);           /* intmul_best hp/code.cpp:148-169 */
/* twiddle table of n = 2N entries: round(cos,sin(2 pi i / n) * 2^64), 1.0 -> 2^64-1 (hp/code.cpp:246-277,378-388) */
void orc_hp_precomp_iFFT(orc_cplx96* powomega, int n);
void orc_hp_precomp_FFT(orc_cplx96* powombar, int n);
void orc_hp_iFFT(orc_cplx96* out /*[n/4]*/, const Torus64* in /*[n/2]*/, int n, const orc_cplx96* powomega); /* :391-443 */
void orc_hp_FFT(Torus64* out /*[n/2]*/, orc_cplx96* in /*[n/4], clobbered*/, int n, const orc_cplx96* powombar); /* :446-512 */

#ifdef __cplusplus
}
#endif
#endif
