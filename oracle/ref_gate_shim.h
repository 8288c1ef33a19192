/* ref_gate_shim.h -- test infrastructure.  The minimum of upstream tfhe/tfhe's headers that the reference's library-style
 * extracts (cb/numeric_functions.cpp, cb/tgsw_functions.cpp, cb/tlwe_functions.cpp, cb/lwe_functions.cpp) need in order to
 * compile function by function.  Those headers are NOT in the reference tree; the struct and field names below are the ones
 * the extracts themselves use (SURVEY.md Appendix C lists where each is visible).  Nothing here computes anything except
 * lweSubTo, which the extracts call but do not define ([UPSTREAM]: coefficient-wise subtraction of LWE samples).
 *
 * oracle/Makefile cuts the function bodies out of the reference sources BY LINE RANGE at build time (sed -n 'a,bp', like
 * patch_poc.sed for the PoC) into _ref/build/gate_extract.cpp, includes this header in front and compiles the result twice
 * (scalar branch and __AVX2__ inline-asm branch of tGswTorus32PolynomialDecompH); ref_harness then pins the oracle's
 * restatements against them bit for bit (tests/golden/pin_log.txt).  No reference source enters the repository. */
#pragma once
#include <cstdint>
#include <cassert>
#include <cstdlib>
#define EXPORT
typedef int32_t Torus32;
struct IntPolynomial { int N; int* coefs; };
struct TorusPolynomial { int N; Torus32* coefsT; };
struct LweParams { int n; double alpha_min, alpha_max; };
struct LweSample { Torus32* a; Torus32 b; double current_variance; };
struct TLweParams { int N, k; double alpha_min, alpha_max; LweParams extracted_lweparams; };
struct TLweSample { TorusPolynomial* a; TorusPolynomial* b; double current_variance; int k; };
struct TGswParams { int l, Bgbit, Bg; int32_t halfBg; uint32_t maskMod; const TLweParams* tlwe_params; int kpl; Torus32* h; uint32_t offset; };
struct LweKeySwitchKey { int n, t, basebit, base; const LweParams* out_params; LweSample* ks0_raw; LweSample** ks1_raw; LweSample*** ks; };
/* [UPSTREAM] lwe-functions: result -= sample */
static inline void lweSubTo(LweSample* result, const LweSample* sample, const LweParams* params) {
    for (int i = 0; i < params->n; ++i) result->a[i] -= sample->a[i];
    result->b -= sample->b;
    result->current_variance += sample->current_variance;
}
