// tfhe_b200_compat.hpp -- C++ host shim: the reference's function names and struct shapes over the C ABI.
//
// A user of tfhe/experimental-tfhe calls free functions on host structs made of separately allocated polynomials, one
// sample per call.  This header keeps those names, argument orders and struct fields and forwards to include/tfhe_b200.h:
// it flattens the structs, stages them through small device buffers and calls the batched entry points with count = 1;
// every function also has a `_batch` twin on flat host arrays.  Nothing here computes: without libtfhe_b200.so and a
// B200 the calls abort() (the reference's own error behaviour is assert/abort, cb/spqlios/spqlios-fft-impl.cpp:92-97).
//
// Two families, as in the reference:
//   library style (cb/lwe_functions.cpp, cb/tlwe_functions.cpp, cb/tgsw_functions.cpp; struct fields per those files):
//       tfhe_MuxRotate_FFT, tfhe_blindRotate_FFT, tfhe_blindRotateAndExtract_FFT, tfhe_bootstrap_woKS_FFT,
//       tfhe_bootstrap_FFT, lweKeySwitch, init_LweBootstrappingKeyFFT, boots* gates (upstream semantics)
//   proof-of-concept style (cb/poc_CircuitBootstrapping.cpp, cb/poc_types.h):
//       preKeySwitch, preModSwitch, circuitBootstrapWoKS, circuitPrivKS, tfhe_CircuitBootstrapFFT over a `Globals`-like env
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/tfhe_b200.h"

namespace tfhe_b200_compat {

typedef int32_t Torus32;   // cb/poc_types.h:13
typedef int64_t Torus64;   // cb/poc_types.h:14

[[noreturn]] inline void die(const char* what, tfhe_b200_ctx* ctx) {
    fprintf(stderr, "tfhe_b200: %s failed: %s\n", what, tfhe_b200_last_error(ctx));
    abort();
}
#define TFHE_B200_CK(call, ctx) do { if ((call) != TFHE_B200_OK) ::tfhe_b200_compat::die(#call, ctx); } while (0)
#define TFHE_B200_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "tfhe_b200: %s: %s\n", #call, cudaGetErrorString(e_)); abort(); } } while (0)

// small RAII device buffer used for count = 1 marshalling
template <typename T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    explicit DevBuf(size_t count) : n(count) { TFHE_B200_CUDA(cudaMalloc(&p, count * sizeof(T))); }
    ~DevBuf() { cudaFree(p); }
    void up(const T* h) { TFHE_B200_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
    void down(T* h) const { TFHE_B200_CUDA(cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost)); }
};

// ============================================================================================================
// Library-style structs (field names as in the reference extracts)
// ============================================================================================================
struct LweParams { int n; double alpha_min, alpha_max; };                                   // cb/lwe_functions.cpp:17
struct LweSample {                                                                           // cb/lwe_functions.cpp:20-25
    Torus32* a; Torus32 b; double current_variance;
    explicit LweSample(const LweParams* params) : a(new Torus32[params->n]), b(0), current_variance(0.) {}
    ~LweSample() { delete[] a; }
    LweSample(const LweSample&) = delete;
};
struct LweKey { const LweParams* params; int* key; };                                        // cb/lwe_functions.cpp:27-31
struct TLweParams { int N, k; double alpha_min, alpha_max; LweParams extracted_lweparams; }; // cb/tlwe_functions.cpp:14-19
struct TorusPolynomial { int N; Torus32* coefsT; explicit TorusPolynomial(int N_) : N(N_), coefsT(new Torus32[N_]) {} ~TorusPolynomial() { delete[] coefsT; } };
struct TLweSample {                                                                          // cb/tlwe_functions.cpp:27-31
    TorusPolynomial* a; TorusPolynomial* b; double current_variance; int k;
    explicit TLweSample(const TLweParams* p) : current_variance(0.), k(p->k) {
        a = static_cast<TorusPolynomial*>(operator new[]((k + 1) * sizeof(TorusPolynomial)));
        for (int i = 0; i <= k; i++) new (a + i) TorusPolynomial(p->N);
        b = a + k;
    }
    ~TLweSample() { for (int i = 0; i <= k; i++) a[i].~TorusPolynomial(); operator delete[](a); }
};
struct TGswParams {                                                                          // cb/tgsw_functions.cpp:15-38
    int l, Bgbit, Bg, halfBg; uint32_t maskMod; const TLweParams* tlwe_params; int kpl; uint32_t offset;
    TGswParams(int l_, int Bgbit_, const TLweParams* tp) : l(l_), Bgbit(Bgbit_), Bg(1 << Bgbit_), halfBg(Bg / 2), maskMod(Bg - 1),
                                                           tlwe_params(tp), kpl((tp->k + 1) * l_) {
        uint32_t t = 0; for (int i = 0; i < l; i++) t += 1u << (32 - (i + 1) * Bgbit);
        offset = t * (uint32_t)halfBg;
    }
};
struct TGswSample {                                                                          // cb/tgsw_functions.cpp:66-76
    TLweSample* all_sample; TLweSample** bloc_sample; int k, l;
    explicit TGswSample(const TGswParams* p) : k(p->tlwe_params->k), l(p->l) {
        all_sample = static_cast<TLweSample*>(operator new[]((k + 1) * l * sizeof(TLweSample)));
        for (int i = 0; i < (k + 1) * l; i++) new (all_sample + i) TLweSample(p->tlwe_params);
        bloc_sample = new TLweSample*[k + 1];
        for (int i = 0; i <= k; i++) bloc_sample[i] = all_sample + i * l;
    }
    ~TGswSample() { for (int i = 0; i < (k + 1) * l; i++) all_sample[i].~TLweSample(); operator delete[](all_sample); delete[] bloc_sample; }
};
struct LweKeySwitchKey {                                                                     // cb/lwe_functions.cpp:96-114
    int n, t, basebit, base; const LweParams* out_params;
    Torus32* ks0_raw;          // flat [n][t][base][out_n+1] (the reference stores LweSample objects; same content, contiguous)
    tfhe_b200_ctx* engine = nullptr;   // set by init_LweBootstrappingKeyFFT
    LweKeySwitchKey(int n_, int t_, int basebit_, const LweParams* out) : n(n_), t(t_), basebit(basebit_), base(1 << basebit_), out_params(out),
        ks0_raw(new Torus32[(size_t)n_ * t_ * (1 << basebit_) * (out->n + 1)]) {}
    ~LweKeySwitchKey() { delete[] ks0_raw; }
    Torus32* entry(int i, int j, int d) { return ks0_raw + (((size_t)i * t + j) * base + d) * (out_params->n + 1); }   // ks[i][j][d]: a[0..n) then b
};
struct LweBootstrappingKey {                                                                 // cb/lwe_functions.cpp:259-269
    const LweParams* in_out_params; const TGswParams* bk_params; const TLweParams* accum_params; const LweParams* extract_params;
    TGswSample** bk;           // bk[i], i < n  (array of pointers: TGswSample is not default constructible)
    LweKeySwitchKey* ks;
};
// The FFT-domain key lives on the device in the engine's private layout; this handle replaces TGswSampleFFT[n].
struct TGswSampleFFT { tfhe_b200_ctx* engine; int index; };
struct LweBootstrappingKeyFFT {                                                              // cb/lwe_functions.cpp:272-282
    const LweParams* in_out_params; const TGswParams* bk_params; const TLweParams* accum_params; const LweParams* extract_params;
    const TGswSampleFFT* bkFFT; const LweKeySwitchKey* ks;
    tfhe_b200_ctx* engine;
};

// init_LweBootstrappingKeyFFT (cb/lwe_functions.cpp:287-316): copies ks, transforms bk (on the device here)
inline void init_LweBootstrappingKeyFFT(LweBootstrappingKeyFFT* obj, const LweBootstrappingKey* bk, int device = 0) {
    tfhe_b200_ctx* ctx = nullptr;
    if (tfhe_b200_ctx_create(&ctx, device) != TFHE_B200_OK) die("tfhe_b200_ctx_create", nullptr);
    const int n = bk->in_out_params->n, N = bk->accum_params->N, kk = bk->accum_params->k, l = bk->bk_params->l;
    std::vector<Torus32> flat((size_t)n * (kk + 1) * l * (kk + 1) * N);
    size_t o = 0;
    for (int i = 0; i < n; i++)
        for (int p = 0; p < (kk + 1) * l; p++)
            for (int q = 0; q <= kk; q++, o += N) memcpy(flat.data() + o, bk->bk[i]->all_sample[p].a[q].coefsT, sizeof(Torus32) * N);
    tfhe_b200_gate_params gp{n, N, kk, l, bk->bk_params->Bgbit, bk->ks->t, bk->ks->basebit};
    TFHE_B200_CK(tfhe_b200_gate_load_keys(ctx, &gp, flat.data(), bk->ks->ks0_raw), ctx);
    TGswSampleFFT* handles = new TGswSampleFFT[n];
    for (int i = 0; i < n; i++) handles[i] = TGswSampleFFT{ctx, i};
    bk->ks->engine = ctx;
    obj->in_out_params = bk->in_out_params; obj->bk_params = bk->bk_params; obj->accum_params = bk->accum_params;
    obj->extract_params = bk->extract_params; obj->bkFFT = handles; obj->ks = bk->ks; obj->engine = ctx;
}
inline void destroy_LweBootstrappingKeyFFT(LweBootstrappingKeyFFT* obj) {                   // cb/lwe_functions.cpp:320-324
    delete[] obj->bkFFT; tfhe_b200_ctx_destroy(obj->engine); obj->engine = nullptr;
}

inline void flatten(const LweSample* s, int n, Torus32* out) { memcpy(out, s->a, sizeof(Torus32) * n); out[n] = s->b; }
inline void unflatten(LweSample* s, int n, const Torus32* in) { memcpy(s->a, in, sizeof(Torus32) * n); s->b = in[n]; }

// ---- batched forms on flat host arrays ([count][n+1] etc.)
inline void tfhe_bootstrap_woKS_FFT_batch(Torus32* result, const LweBootstrappingKeyFFT* bk, Torus32 mu, const Torus32* x, int count) {
    const int n = bk->in_out_params->n, N = bk->accum_params->N;
    DevBuf<Torus32> dx((size_t)count * (n + 1)), dr((size_t)count * (N + 1));
    dx.up(x);
    TFHE_B200_CK(tfhe_b200_bootstrap_woKS_FFT_batch(bk->engine, dr.p, mu, dx.p, count, nullptr), bk->engine);
    dr.down(result);
}
inline void tfhe_bootstrap_FFT_batch(Torus32* result, const LweBootstrappingKeyFFT* bk, Torus32 mu, const Torus32* x, int count) {
    const int n = bk->in_out_params->n;
    DevBuf<Torus32> dx((size_t)count * (n + 1)), dr((size_t)count * (n + 1));
    dx.up(x);
    TFHE_B200_CK(tfhe_b200_bootstrap_FFT_batch(bk->engine, dr.p, mu, dx.p, count, nullptr), bk->engine);
    dr.down(result);
}
inline void lweKeySwitch_batch(Torus32* result, const LweKeySwitchKey* ks, const Torus32* sample, int count) {
    DevBuf<Torus32> ds((size_t)count * (ks->n + 1)), dr((size_t)count * (ks->out_params->n + 1));
    ds.up(sample);
    TFHE_B200_CK(tfhe_b200_lweKeySwitch_batch(ks->engine, dr.p, ds.p, count, nullptr), ks->engine);
    dr.down(result);
}
// bkFFT may point INTO the key array (the reference passes bkFFT+i, cb/lwe_functions.cpp:352) and n may be smaller than the key's n:
// the engine always walks its whole key, so the caller's n rotation amounts are placed at offset bkFFT->index of a zero-padded row
// (a step with bara = 0 is skipped, :350).  A range that does not fit the key aborts, like the reference's asserts.
inline void tfhe_blindRotate_FFT_batch(Torus32* accum /*[count][k+1][N]*/, const TGswSampleFFT* bkFFT, const int* bara /*[count][n]*/, int n,
                                       const TGswParams* bk_params, int count) {
    const int N = bk_params->tlwe_params->N;
    tfhe_b200_gate_params gp;
    TFHE_B200_CK(tfhe_b200_gate_get_params(bkFFT->engine, &gp), bkFFT->engine);
    if (n < 0 || bkFFT->index < 0 || bkFFT->index + n > gp.n || N != gp.N) {
        fprintf(stderr, "tfhe_blindRotate_FFT: steps [%d, %d) do not fit the loaded key (n = %d, N = %d)\n", bkFFT->index, bkFFT->index + n, gp.n, gp.N);
        abort();
    }
    std::vector<int> padded((size_t)count * gp.n, 0);
    for (int c = 0; c < count; c++) memcpy(padded.data() + (size_t)c * gp.n + bkFFT->index, bara + (size_t)c * n, sizeof(int) * n);
    DevBuf<Torus32> da((size_t)count * 2 * N); DevBuf<int> db((size_t)count * gp.n);
    da.up(accum); db.up(padded.data());
    TFHE_B200_CK(tfhe_b200_blindRotate_FFT_batch(bkFFT->engine, da.p, db.p, count, nullptr), bkFFT->engine);
    da.down(accum);
}

// ---- the reference's single-sample signatures
// tfhe_blindRotate_FFT (cb/lwe_functions.cpp:337-361)
inline void tfhe_blindRotate_FFT(TLweSample* accum, const TGswSampleFFT* bkFFT, const int* bara, const int n, const TGswParams* bk_params) {
    const int N = bk_params->tlwe_params->N;
    std::vector<Torus32> flat(2 * N);
    for (int q = 0; q < 2; q++) memcpy(flat.data() + q * N, accum->a[q].coefsT, sizeof(Torus32) * N);
    tfhe_blindRotate_FFT_batch(flat.data(), bkFFT, bara, n, bk_params, 1);
    for (int q = 0; q < 2; q++) memcpy(accum->a[q].coefsT, flat.data() + q * N, sizeof(Torus32) * N);
}
// tfhe_MuxRotate_FFT (cb/lwe_functions.cpp:328-333), the reference's signature: result = accum + bki (x) ((X^barai - 1) accum); bki = bkFFT + i
inline void tfhe_MuxRotate_FFT(TLweSample* result, const TLweSample* accum, const TGswSampleFFT* bki, const int barai, const TGswParams* bk_params) {
    const int N = bk_params->tlwe_params->N;
    for (int q = 0; q < 2; q++) memcpy(result->a[q].coefsT, accum->a[q].coefsT, sizeof(Torus32) * N);
    if (barai == 0) return;                        // (X^0 - 1) accum = 0
    tfhe_blindRotate_FFT(result, bki, &barai, 1, bk_params);      // one step at the key entry bki points at
}
// tfhe_blindRotateAndExtract_FFT (cb/lwe_functions.cpp:366-395)
inline void tfhe_blindRotateAndExtract_FFT(LweSample* result, const TorusPolynomial* v, const TGswSampleFFT* bk, const int barb, const int* bara,
                                           const int n, const TGswParams* bk_params) {
    const int N = bk_params->tlwe_params->N;
    DevBuf<Torus32> dv(N), dr(N + 1); DevBuf<int> dbb(1), dba(n);
    dv.up(v->coefsT); dbb.up(&barb); dba.up(bara);
    TFHE_B200_CK(tfhe_b200_blindRotateAndExtract_FFT_batch(bk->engine, dr.p, dv.p, dbb.p, dba.p, 1, nullptr), bk->engine);
    std::vector<Torus32> out(N + 1); dr.down(out.data());
    unflatten(result, N, out.data());
}
// tfhe_bootstrap_woKS_FFT (cb/lwe_functions.cpp:399-430)
inline void tfhe_bootstrap_woKS_FFT(LweSample* result, const LweBootstrappingKeyFFT* bk, Torus32 mu, const LweSample* x) {
    const int n = bk->in_out_params->n, N = bk->accum_params->N;
    std::vector<Torus32> in(n + 1), out(N + 1);
    flatten(x, n, in.data());
    tfhe_bootstrap_woKS_FFT_batch(out.data(), bk, mu, in.data(), 1);
    unflatten(result, N, out.data());
}
// tfhe_bootstrap_FFT (cb/lwe_functions.cpp:434-446)
inline void tfhe_bootstrap_FFT(LweSample* result, const LweBootstrappingKeyFFT* bk, Torus32 mu, const LweSample* x) {
    const int n = bk->in_out_params->n;
    std::vector<Torus32> in(n + 1), out(n + 1);
    flatten(x, n, in.data());
    tfhe_bootstrap_FFT_batch(out.data(), bk, mu, in.data(), 1);
    unflatten(result, n, out.data());
}
// lweKeySwitch (cb/lwe_functions.cpp:163-171)
inline void lweKeySwitch(LweSample* result, const LweKeySwitchKey* ks, const LweSample* sample) {
    std::vector<Torus32> in(ks->n + 1), out(ks->out_params->n + 1);
    flatten(sample, ks->n, in.data());
    lweKeySwitch_batch(out.data(), ks, in.data(), 1);
    unflatten(result, ks->out_params->n, out.data());
}

// ---- boots* gates (upstream tfhe/tfhe boot-gates.cpp)
struct TFheGateBootstrappingCloudKeySet { const LweBootstrappingKeyFFT* bkFFT; };
inline void bootsGate_batch(int op, Torus32* result, const Torus32* ca, const Torus32* cb, int count, const TFheGateBootstrappingCloudKeySet* bk) {
    TFHE_B200_CK(tfhe_b200_bootsGate_batch_host(bk->bkFFT->engine, op, result, ca, cb, count), bk->bkFFT->engine);
}
inline void bootsGate(int op, LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk) {
    const int n = bk->bkFFT->in_out_params->n;
    std::vector<Torus32> a(n + 1), b(n + 1), r(n + 1);
    flatten(ca, n, a.data()); flatten(cb, n, b.data());
    bootsGate_batch(op, r.data(), a.data(), b.data(), 1, bk);
    unflatten(result, n, r.data());
}
#define TFHE_B200_GATE(NAME, OP)                                                                                                            \
    inline void boots##NAME(LweSample* result, const LweSample* ca, const LweSample* cb, const TFheGateBootstrappingCloudKeySet* bk) {     \
        bootsGate(OP, result, ca, cb, bk);                                                                                                  \
    }                                                                                                                                       \
    inline void boots##NAME##_batch(Torus32* result, const Torus32* ca, const Torus32* cb, int count, const TFheGateBootstrappingCloudKeySet* bk) { \
        bootsGate_batch(OP, result, ca, cb, count, bk);                                                                                     \
    }
TFHE_B200_GATE(NAND, TFHE_B200_NAND) TFHE_B200_GATE(AND, TFHE_B200_AND) TFHE_B200_GATE(OR, TFHE_B200_OR) TFHE_B200_GATE(NOR, TFHE_B200_NOR)
TFHE_B200_GATE(XOR, TFHE_B200_XOR) TFHE_B200_GATE(XNOR, TFHE_B200_XNOR) TFHE_B200_GATE(ANDNY, TFHE_B200_ANDNY) TFHE_B200_GATE(ANDYN, TFHE_B200_ANDYN)
TFHE_B200_GATE(ORNY, TFHE_B200_ORNY) TFHE_B200_GATE(ORYN, TFHE_B200_ORYN)
#undef TFHE_B200_GATE
inline void bootsNOT(LweSample* result, const LweSample* ca, const TFheGateBootstrappingCloudKeySet* bk) {
    const int n = bk->bkFFT->in_out_params->n;                     // negation, no bootstrapping
    for (int i = 0; i < n; i++) result->a[i] = (Torus32)(0u - (uint32_t)ca->a[i]);
    result->b = (Torus32)(0u - (uint32_t)ca->b);
}
inline void bootsMUX(LweSample* result, const LweSample* a, const LweSample* b, const LweSample* c, const TFheGateBootstrappingCloudKeySet* bk) {
    tfhe_b200_ctx* ctx = bk->bkFFT->engine;
    const int n = bk->bkFFT->in_out_params->n;
    std::vector<Torus32> h(n + 1);
    DevBuf<Torus32> da(n + 1), db(n + 1), dc(n + 1), dr(n + 1);
    flatten(a, n, h.data()); da.up(h.data()); flatten(b, n, h.data()); db.up(h.data()); flatten(c, n, h.data()); dc.up(h.data());
    TFHE_B200_CK(tfhe_b200_bootsMUX_batch(ctx, dr.p, da.p, db.p, dc.p, 1, nullptr), ctx);
    dr.down(h.data()); unflatten(result, n, h.data());
}

// ============================================================================================================
// Proof-of-concept style (cb/poc_types.h, cb/poc_CircuitBootstrapping.cpp)
// ============================================================================================================
struct LweSample32 { Torus32* const a; Torus32* const b; explicit LweSample32(int n) : a(new Torus32[n + 1]), b(&a[n]) {} ~LweSample32() { delete[] a; } };   // poc_types.h:137-144
struct LweSample64 { Torus64* const a; Torus64* const b; explicit LweSample64(int n) : a(new Torus64[n + 1]), b(&a[n]) {} ~LweSample64() { delete[] a; } };   // :151-158
struct Torus32Polynomial { Torus32* const coefs; explicit Torus32Polynomial(int N) : coefs(new Torus32[N]) {} ~Torus32Polynomial() { delete[] coefs; } };    // :40-46
struct TLweSample32 {                                                                                                                                       // :164-171
    Torus32Polynomial* a; Torus32Polynomial* b;
    explicit TLweSample32(int N) { a = static_cast<Torus32Polynomial*>(operator new[](2 * sizeof(Torus32Polynomial))); new (a) Torus32Polynomial(N); new (a + 1) Torus32Polynomial(N); b = a + 1; }
    ~TLweSample32() { a[0].~Torus32Polynomial(); a[1].~Torus32Polynomial(); operator delete[](a); }
};
struct TGswSample32 {                                                                                                                                       // :206-217  samples[k+1][l]
    int l; std::vector<TLweSample32*> store; TLweSample32*** samples;
    TGswSample32(int l_, int N) : l(l_) {
        samples = new TLweSample32**[2];
        for (int u = 0; u < 2; u++) { samples[u] = new TLweSample32*[l]; for (int w = 0; w < l; w++) { samples[u][w] = new TLweSample32(N); store.push_back(samples[u][w]); } }
    }
    ~TGswSample32() { for (auto* s : store) delete s; delete[] samples[0]; delete[] samples[1]; delete[] samples; }
};
// `Globals` (cb/poc_types.h:267-312) reduced to what the hot path reads: the parameters and the engine holding the cloud keys.
struct Globals {
    tfhe_b200_cb_params p;
    int n_lvl0, n_lvl1, n_lvl2, N_lvl1, N_lvl2, ell_lvl1, bgbit_lvl1;
    tfhe_b200_ctx* engine;
    // replaces the cloud-key part of Globals::Globals (cb/poc_CircuitBootstrapping.cpp:372-419): keys come in as flat host arrays
    Globals(const tfhe_b200_cb_params& params, const int32_t* preKS, const int64_t* bk, const int32_t* privKS, int device = 0) : p(params), engine(nullptr) {
        n_lvl0 = p.n_lvl0; n_lvl1 = N_lvl1 = p.N_lvl1; n_lvl2 = N_lvl2 = p.N_lvl2; ell_lvl1 = p.ell_lvl1; bgbit_lvl1 = p.bgbit_lvl1;
        if (tfhe_b200_ctx_create(&engine, device) != TFHE_B200_OK) die("tfhe_b200_ctx_create", nullptr);
        TFHE_B200_CK(tfhe_b200_cb_load_keys(engine, &p, preKS, bk, privKS), engine);
    }
    ~Globals() { tfhe_b200_ctx_destroy(engine); }
};
// preKeySwitch (cb/poc_CircuitBootstrapping.cpp:437-465)
inline void preKeySwitch(LweSample32* result, const LweSample32* x, const Globals* env) {
    DevBuf<Torus32> dx(env->n_lvl1 + 1), dr(env->n_lvl0 + 1);
    dx.up(x->a);
    TFHE_B200_CK(tfhe_b200_preKeySwitch_batch(env->engine, dr.p, dx.p, 1, nullptr), env->engine);
    dr.down(result->a);
}
// preModSwitch (:472-484)
inline void preModSwitch(int* result, const LweSample32* x, const Globals* env) {
    DevBuf<Torus32> dx(env->n_lvl0 + 1), dr(env->n_lvl0 + 1);
    dx.up(x->a);
    TFHE_B200_CK(tfhe_b200_preModSwitch_batch(env->engine, dr.p, dx.p, 1, nullptr), env->engine);
    dr.down(result);
}
// circuitBootstrapWoKS (:530-659)
inline void circuitBootstrapWoKS(LweSample64* result, const Torus64 mu, const int* abar, const Globals* env) {
    DevBuf<int> da(env->n_lvl0 + 1); DevBuf<Torus64> dr(env->n_lvl2 + 1);
    da.up(abar);
    TFHE_B200_CK(tfhe_b200_circuitBootstrapWoKS_batch(env->engine, dr.p, mu, da.p, 1, nullptr), env->engine);
    dr.down(result->a);
}
// circuitPrivKS (:667-698)
inline void circuitPrivKS(TLweSample32* result, const int u, const LweSample64* x, const Globals* env) {
    DevBuf<Torus64> dx(env->n_lvl2 + 1); DevBuf<Torus32> dr(2 * env->N_lvl1);
    dx.up(x->a);
    TFHE_B200_CK(tfhe_b200_circuitPrivKS_batch(env->engine, dr.p, u, dx.p, 1, nullptr), env->engine);
    std::vector<Torus32> h(2 * env->N_lvl1); dr.down(h.data());
    memcpy(result->a[0].coefs, h.data(), sizeof(Torus32) * env->N_lvl1);
    memcpy(result->a[1].coefs, h.data() + env->N_lvl1, sizeof(Torus32) * env->N_lvl1);
}
// tfhe_CircuitBootstrapFFT (:823-873), batched on flat arrays: result[count][2][l1][2][N1], sample[count][N1+1]
inline void tfhe_CircuitBootstrapFFT_batch(Torus32* result, const Torus32* sample, int count, const Globals* env) {
    TFHE_B200_CK(tfhe_b200_CircuitBootstrapFFT_batch_host(env->engine, result, sample, count), env->engine);
}
inline void tfhe_CircuitBootstrapFFT(TGswSample32* result, const LweSample32* sample, const Globals* env) {
    const int N1 = env->N_lvl1, l1 = env->ell_lvl1;
    std::vector<Torus32> out((size_t)2 * l1 * 2 * N1);
    tfhe_CircuitBootstrapFFT_batch(out.data(), sample->a, 1, env);
    for (int u = 0; u < 2; u++) for (int w = 0; w < l1; w++) for (int q = 0; q < 2; q++)
        memcpy(result->samples[u][w]->a[q].coefs, out.data() + (((size_t)u * l1 + w) * 2 + q) * N1, sizeof(Torus32) * N1);
}

// ---------------------------------------------------------------------------------------------------------------
// Gate-level circuits (BASELINE configs[2]; no reference counterpart, see include/tfhe_b200.h: tfhe_b200_circuit_eval_batch).
// Ripple-carry adder, `bits` wide: per bit t = a^b, g = a&b, s = t^c, p = t&c, c' = g|p  (2 XOR + 2 AND + 1 OR = 5 bootstrapped
// gates per bit).  Wire map: a[i] = i, b[i] = bits+i, cin = 2 bits, s[i] = 2 bits+1+i, cout = c[bits];
// internal t, g, c follow.  The 2*bits level-0 gates come first and merge into two launches of bits*count samples.
// ---------------------------------------------------------------------------------------------------------------
struct AdderNetlist {
    int bits, n_wires;
    int a0, b0, cin, s0, c0;          // first wire of each bus; carry c[i] = c0 + i, c[0] = cin copied, cout = c0 + bits
    std::vector<tfhe_b200_gate> gates;
};
inline AdderNetlist ripple_carry_adder_netlist(int bits) {
    AdderNetlist nl;
    nl.bits = bits; nl.a0 = 0; nl.b0 = bits; nl.cin = 2 * bits; nl.s0 = 2 * bits + 1;
    const int t0 = nl.s0 + bits, g0 = t0 + bits, p0 = g0 + bits;
    nl.c0 = p0 + bits; nl.n_wires = nl.c0 + bits + 1;
    for (int i = 0; i < bits; i++) nl.gates.push_back({TFHE_B200_XOR, t0 + i, nl.a0 + i, nl.b0 + i, 0});
    for (int i = 0; i < bits; i++) nl.gates.push_back({TFHE_B200_AND, g0 + i, nl.a0 + i, nl.b0 + i, 0});
    nl.gates.push_back({TFHE_B200_COPY, nl.c0, nl.cin, 0, 0});
    for (int i = 0; i < bits; i++) {
        nl.gates.push_back({TFHE_B200_XOR, nl.s0 + i, t0 + i, nl.c0 + i, 0});
        nl.gates.push_back({TFHE_B200_AND, p0 + i, t0 + i, nl.c0 + i, 0});
        nl.gates.push_back({TFHE_B200_OR, nl.c0 + i + 1, g0 + i, p0 + i, 0});
    }
    return nl;
}

}  // namespace tfhe_b200_compat
