// ref_harness.cpp -- drives the REFERENCE's own code (compiled in place from /root/reference by oracle/Makefile,
// with the three corrections of SURVEY.md Appendix B) next to the oracle restatement.  TEST INFRASTRUCTURE.
//
//   ref_harness golden <dir>           pin the oracle against the reference and write golden vectors
//   ref_harness bench-gate <count> <threads> [reps]   gate bootstraps/s: oracle gate path on the reference's spqlios kernels
//   ref_harness bench-cb <count>       circuit bootstraps/s: the reference's tfhe_CircuitBootstrapFFT (single thread: it is
//                                      not re-entrant)
//
// Key material always comes from the oracle's deterministic keygen (orc_*_keygen) and is copied INTO the reference's
// structs, so both sides see identical keys and inputs (the reference's RNG is libstdc++-specific).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <chrono>
#include <omp.h>
#include "generic_utils.h"
#define k 1
#include "spqlios/lagrangehalfc_impl.h"
#include "poc_types.h"
#undef k
#include "tfhe_oracle.h"

// ---- reference entry points (defined in the patched PoC translation unit / poc_karatsuba.cpp)
void preKeySwitch(LweSample32* result, const LweSample32* x, const Globals* env);
void preModSwitch(int* result, const LweSample32* x, const Globals* env);
void circuitBootstrapWoKS(LweSample64* result, const Torus64 mu, const int* abar, const Globals* env);
void circuitPrivKS(TLweSample32* result, const int u, const LweSample64* x, const Globals* env);
void tfhe_CircuitBootstrapFFT(TGswSample32* result, const LweSample32* sample, const Globals* env);
void tGsw64DecompH(IntPolynomial* result, const TLweSample64* sample, const Globals* env);
void TorusPolynomial64_ifft_lvl2(LagrangeHalfCPolynomial* result, const Torus64Polynomial* source, const Globals* env);
void torus32PolynomialMultAddKaratsuba_lvl1(Torus32Polynomial* result, const IntPolynomial* poly1, const Torus32Polynomial* poly2, const Globals* env);
void torus64PolynomialMultAddKaratsuba_lvl2(Torus64Polynomial* result, const IntPolynomial* poly1, const Torus64Polynomial* poly2, const Globals* env);

// ---- the reference's spqlios kernels as an oracle FFT backend (one processor per thread: they are not re-entrant)
static FFT_Processor_Spqlios* proc(int N) {
    static thread_local FFT_Processor_Spqlios* p1024 = nullptr;
    static thread_local FFT_Processor_Spqlios* p2048 = nullptr;
    if (N == 1024) { if (!p1024) p1024 = new FFT_Processor_Spqlios(1024); return p1024; }
    if (N == 2048) { if (!p2048) p2048 = new FFT_Processor_Spqlios(2048); return p2048; }
    fprintf(stderr, "spqlios backend: unsupported N=%d\n", N); abort();
}
static void s_ifft_int(int N, double* res, const int32_t* a) { proc(N)->execute_reverse_int(res, (const int*)a); }
static void s_ifft_t64(int N, double* res, const int64_t* a) { proc(N)->execute_reverse_torus64(res, a); }
// execute_direct_* keep `static const double _2sN` initialised by the first N that calls them; the harness
// only ever uses torus32 with N=1024 and torus64 with N=2048, like the reference.
static void s_fft_t32(int N, int32_t* res, const double* a) { if (N != 1024) abort(); proc(N)->execute_direct_torus32(res, a); }
static void s_fft_t64(int N, int64_t* res, const double* a) { if (N != 2048) abort(); proc(N)->execute_direct_torus64(res, a); }
static void s_addmul(int N, double* res, const double* a, const double* b) {
    LagrangeHalfCPolynomialAddMulASM(res, const_cast<double*>(a), const_cast<double*>(b), N / 2);
}
static const orc_fft_backend kSpqlios = { s_ifft_int, s_ifft_t64, s_fft_t32, s_fft_t64, s_addmul };

// ---- Globals filled from oracle keys, without running Globals::Globals (which draws its own ~100 s of keys)
static Globals* make_env(const orc_cb_keys* K, bool with_priv) {
    Globals* env = (Globals*)calloc(1, sizeof(Globals));
    const int n0 = Globals::n_lvl0, n1 = Globals::n_lvl1, n2 = Globals::n_lvl2, l2 = Globals::ell_lvl2;
    if (K->p.n_lvl0 != n0 || K->p.N_lvl1 != n1 || K->p.N_lvl2 != n2 || K->p.ell_lvl2 != l2 || K->p.bgbit_lvl2 != Globals::bgbit_lvl2 ||
        K->p.kslength_lvl10 != Globals::kslength_lvl10 || K->p.ksbasebit_lvl10 != Globals::ksbasebit_lvl10 ||
        K->p.kslength_lvl21 != Globals::kslength_lvl21 || K->p.ksbasebit_lvl21 != Globals::ksbasebit_lvl21 ||
        K->p.ell_lvl1 != Globals::ell_lvl1 || K->p.bgbit_lvl1 != Globals::bgbit_lvl1) {
        fprintf(stderr, "oracle default parameters differ from the reference's active set\n"); abort();
    }
    env->t_lvl0 = Globals::kslength_lvl10 * Globals::ksbasebit_lvl10;
    env->t_lvl1 = Globals::kslength_lvl21 * Globals::ksbasebit_lvl21;
    env->N_lvl1 = n1; env->N_lvl2 = n2;
    env->torusDecompOffset = 0;
    for (int i = 0; i <= l2; ++i) env->torusDecompOffset |= (UINT64_C(1) << (63 - i * Globals::bgbit_lvl2));
    env->torusDecompBuf = new uint64_t[n2];
    env->key_lvl0 = new int[n0]; for (int i = 0; i < n0; i++) env->key_lvl0[i] = K->key_lvl0[i];
    env->key_lvl1 = new int[n1]; for (int i = 0; i < n1; i++) env->key_lvl1[i] = K->key_lvl1[i];
    env->Key_lvl1 = new IntPolynomial(n1); for (int i = 0; i < n1; i++) env->Key_lvl1->coefs[i] = K->key_lvl1[i];
    env->key_lvl2 = new int[n2 + 1]; for (int i = 0; i <= n2; i++) env->key_lvl2[i] = K->key_lvl2[i];
    env->Key_lvl2 = new IntPolynomial(n2); for (int i = 0; i < n2; i++) env->Key_lvl2->coefs[i] = K->key_lvl2[i];
    const int t10 = Globals::kslength_lvl10, base10 = 1 << Globals::ksbasebit_lvl10;
    env->preKS = new_array3<LweSample32>(n1, t10, base10, n0);
    for (int i = 0; i < n1; i++) for (int j = 0; j < t10; j++) for (int u = 0; u < base10; u++)
        memcpy(env->preKS[i][j][u].a, K->preKS + (((size_t)i * t10 + j) * base10 + u) * (n0 + 1), sizeof(Torus32) * (n0 + 1));
    env->bk = new_array1<TGswSample64>(n0, l2, n2);
    env->bkFFT = new_array1<TGswSampleFFT>(n0, l2, n2);
    for (int i = 0; i < n0; i++)
        for (int p = 0; p < 2 * l2; p++)
            for (int q = 0; q < 2; q++) {
                memcpy(env->bk[i].allsamples[p].a[q].coefs, K->bk + ((((size_t)i * 2 * l2 + p) * 2) + q) * n2, sizeof(Torus64) * n2);
                TorusPolynomial64_ifft_lvl2(&env->bkFFT[i].allsamples[p].a[q], &env->bk[i].allsamples[p].a[q], env);   // poc:396-402
            }
    env->privKS = nullptr;
    if (with_priv && K->privKS) {
        const int t21 = Globals::kslength_lvl21, base21 = 1 << Globals::ksbasebit_lvl21;
        env->privKS = new_array4<TLweSample32>(2, n2 + 1, t21, base21, n1);
        for (int z = 0; z < 2; z++) for (int i = 0; i <= n2; i++) for (int j = 0; j < t21; j++) for (int u = 0; u < base21; u++) {
            const Torus32* src = K->privKS + (((((size_t)z * (n2 + 1) + i) * t21 + j) * base21 + u) * 2) * n1;
            memcpy(env->privKS[z][i][j][u].a[0].coefs, src, sizeof(Torus32) * n1);
            memcpy(env->privKS[z][i][j][u].a[1].coefs, src + n1, sizeof(Torus32) * n1);
        }
    }
    return env;
}

static void dump(const std::string& dir, const char* name, const void* data, size_t bytes) {
    std::string path = dir + "/" + name;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(data, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fclose(f);
}
static int g_fail = 0;
#define PIN(cond, msg) do { if (!(cond)) { printf("PIN FAILED: %s\n", msg); g_fail++; } else printf("pinned: %s\n", msg); } while (0)

extern "C" int ref_run_pins(const char* golden_dir);
static int cmd_golden(const std::string& dir) {
    orc_rng r; orc_rng_seed(&r, 2026);
    // ---------------- spqlios transforms vs the portable restatement (same ordering, same conventions)
    for (int N : {1024, 2048}) {
        std::vector<int32_t> a(N); std::vector<double> ref(N), port(N);
        for (int i = 0; i < N; i++) a[i] = (int32_t)(orc_rng_u64(&r) % 1024) - 512;
        kSpqlios.ifft_int(N, ref.data(), a.data());
        orc_get_fft_backend()->ifft_int(N, port.data(), a.data());
        double md = 0, mx = 0;
        for (int i = 0; i < N; i++) { md = fmax(md, fabs(ref[i] - port[i])); mx = fmax(mx, fabs(ref[i])); }
        char msg[128]; snprintf(msg, sizeof msg, "ifft_int N=%d: |spqlios - portable| = %.3g (max |value| %.3g)", N, md, mx);
        PIN(md <= 1e-5, msg);     // the reference's own asm-vs-model bar, cb/spqlios/spqlios-bench.cpp:63-68
        char nm[64];
        snprintf(nm, sizeof nm, "fft_in_int_N%d.i32", N); dump(dir, nm, a.data(), N * 4);
        snprintf(nm, sizeof nm, "fft_out_spqlios_N%d.f64", N); dump(dir, nm, ref.data(), N * 8);
        // round trip through the reference: execute_direct_torus32(execute_reverse_int(x)) truncates toward zero
        // (cb/spqlios/fft_processor_spqlios.cpp:102), so it is the identity only up to 1 LSB
        if (N == 1024) {
            std::vector<int32_t> back(N);
            kSpqlios.fft_torus32(N, back.data(), ref.data());
            int worst = 0, off = 0;
            for (int i = 0; i < N; i++) { int d = abs(back[i] - a[i]); if (d > worst) worst = d; off += d != 0; }
            snprintf(msg, sizeof msg, "spqlios round trip N=1024 within 1 LSB (worst %d, %d of %d coefficients off by one)", worst, off, N);
            PIN(worst <= 1, msg);
            dump(dir, "fft_roundtrip_spqlios_N1024.i32", back.data(), N * 4);
        }
    }
    // ---------------- Karatsuba (reference) == naive restatement, exactly
    {
        const int N = 1024;
        Globals* sizes = (Globals*)calloc(1, sizeof(Globals)); sizes->N_lvl1 = 1024; sizes->N_lvl2 = 2048;
        IntPolynomial p1(N); Torus32Polynomial p2(N), res(N); std::vector<int32_t> mine(N, 0);
        for (int i = 0; i < N; i++) { p1.coefs[i] = (int)(orc_rng_u64(&r) % 1024) - 512; p2.coefs[i] = (Torus32)orc_rng_u64(&r); res.coefs[i] = 0; }
        torus32PolynomialMultAddKaratsuba_lvl1(&res, &p1, &p2, sizes);
        orc_torus32PolynomialMultAddNaive(mine.data(), p1.coefs, p2.coefs, N);
        int bad = 0; for (int i = 0; i < N; i++) bad += mine[i] != res.coefs[i];
        PIN(bad == 0, "torus32PolynomialMultAddKaratsuba_lvl1 == orc_torus32PolynomialMultAddNaive (N=1024)");
        dump(dir, "kara32_p1.i32", p1.coefs, N * 4); dump(dir, "kara32_p2.i32", p2.coefs, N * 4); dump(dir, "kara32_out.i32", res.coefs, N * 4);
        const int N2 = 2048;
        IntPolynomial q1(N2); Torus64Polynomial q2(N2), qres(N2); std::vector<int64_t> mine64(N2, 0);
        for (int i = 0; i < N2; i++) { q1.coefs[i] = (int)(orc_rng_u64(&r) % 512) - 256; q2.coefs[i] = (Torus64)orc_rng_u64(&r); qres.coefs[i] = 0; }
        torus64PolynomialMultAddKaratsuba_lvl2(&qres, &q1, &q2, sizes);
        orc_torus64PolynomialMultAddNaive(mine64.data(), q1.coefs, q2.coefs, N2);
        bad = 0; for (int i = 0; i < N2; i++) bad += mine64[i] != qres.coefs[i];
        PIN(bad == 0, "torus64PolynomialMultAddKaratsuba_lvl2 == orc_torus64PolynomialMultAddNaive (N=2048)");
    }
    // ---------------- circuit-bootstrapping pipeline on oracle keys, seed 42
    orc_cb_params cp; orc_cb_params_default(&cp);
    printf("generating oracle keys (seed 42, with privKS)...\n"); fflush(stdout);
    orc_cb_keys* K = orc_cb_keygen(&cp, 42, 1);
    Globals* env = make_env(K, true);
    const int n0 = cp.n_lvl0, N1 = cp.N_lvl1, N2 = cp.N_lvl2, ell1 = cp.ell_lvl1;
    const int NS = 4;
    std::vector<int32_t> in(NS * (N1 + 1));
    orc_rng rin; orc_rng_seed(&rin, 45);
    for (int s = 0; s < NS; s++) orc_lwe32Encrypt_lvl1(in.data() + s * (N1 + 1), (s & 1) ? (Torus32)(1u << 31) : 0, ldexp(1.0, -20), K, &rin);
    dump(dir, "cb_in.i32", in.data(), in.size() * 4);
    std::vector<int32_t> g_pre(NS * (n0 + 1)), g_ms(NS * (n0 + 1));
    std::vector<int64_t> g_boot((size_t)NS * ell1 * (N2 + 1));
    std::vector<int32_t> g_out((size_t)NS * 2 * ell1 * 2 * N1);
    orc_set_fft_backend(&kSpqlios);
    orc_cb_keys_rebuild_fft(K);                 // oracle bkFFT through the reference's own transform
    int bad_pre = 0, bad_ms = 0, bad_boot = 0, bad_out = 0, bad_priv = 0;
    for (int s = 0; s < NS; s++) {
        LweSample32 x(N1); memcpy(x.a, in.data() + s * (N1 + 1), sizeof(Torus32) * (N1 + 1));
        LweSample32 pre(n0); preKeySwitch(&pre, &x, env);
        memcpy(g_pre.data() + s * (n0 + 1), pre.a, sizeof(Torus32) * (n0 + 1));
        std::vector<int> ms(n0 + 1); preModSwitch(ms.data(), &pre, env);
        for (int i = 0; i <= n0; i++) g_ms[s * (n0 + 1) + i] = ms[i];
        std::vector<int32_t> o_pre(n0 + 1), o_ms(n0 + 1);
        orc_preKeySwitch(o_pre.data(), in.data() + s * (N1 + 1), K);
        orc_preModSwitch(o_ms.data(), o_pre.data(), n0, N2);
        bad_pre += memcmp(o_pre.data(), pre.a, sizeof(Torus32) * (n0 + 1)) != 0;
        for (int i = 0; i <= n0; i++) bad_ms += o_ms[i] != ms[i];
        for (int w = 0; w < ell1; w++) {
            const Torus64 mu = (Torus64)(UINT64_C(1) << (64 - (w + 1) * cp.bgbit_lvl1));
            LweSample64 boot(N2); circuitBootstrapWoKS(&boot, mu, ms.data(), env);
            memcpy(g_boot.data() + ((size_t)s * ell1 + w) * (N2 + 1), boot.a, sizeof(Torus64) * (N2 + 1));
            std::vector<int64_t> o_boot(N2 + 1);
            orc_circuitBootstrapWoKS(o_boot.data(), mu, o_ms.data(), K);
            bad_boot += memcmp(o_boot.data(), boot.a, sizeof(Torus64) * (N2 + 1)) != 0;
            for (int u = 0; u < 2; u++) {
                TLweSample32 row(N1); circuitPrivKS(&row, u, &boot, env);
                std::vector<int32_t> o_row(2 * N1);
                orc_circuitPrivKS(o_row.data(), u, o_boot.data(), K);
                bad_priv += memcmp(o_row.data(), row.a[0].coefs, sizeof(Torus32) * N1) != 0;
                bad_priv += memcmp(o_row.data() + N1, row.a[1].coefs, sizeof(Torus32) * N1) != 0;
            }
        }
        TGswSample32 res(ell1, N1); tfhe_CircuitBootstrapFFT(&res, &x, env);
        std::vector<int32_t> o_res((size_t)2 * ell1 * 2 * N1);
        orc_tfhe_CircuitBootstrapFFT(o_res.data(), in.data() + s * (N1 + 1), K);
        for (int u = 0; u < 2; u++) for (int w = 0; w < ell1; w++) for (int q = 0; q < 2; q++) {
            int32_t* dst = g_out.data() + ((((size_t)s * 2 + u) * ell1 + w) * 2 + q) * N1;
            memcpy(dst, res.samples[u][w].a[q].coefs, sizeof(Torus32) * N1);
            bad_out += memcmp(dst, o_res.data() + (((size_t)u * ell1 + w) * 2 + q) * N1, sizeof(Torus32) * N1) != 0;
        }
    }
    PIN(bad_pre == 0, "preKeySwitch: oracle == reference, bit-exact");
    PIN(bad_ms == 0, "preModSwitch: oracle == reference, bit-exact");
    PIN(bad_boot == 0, "circuitBootstrapWoKS (patched D1-D3): oracle on spqlios backend == reference, bit-exact");
    PIN(bad_priv == 0, "circuitPrivKS: oracle == reference, bit-exact");
    PIN(bad_out == 0, "tfhe_CircuitBootstrapFFT: oracle on spqlios backend == reference, bit-exact");
    dump(dir, "cb_preks.i32", g_pre.data(), g_pre.size() * 4);
    dump(dir, "cb_prems.i32", g_ms.data(), g_ms.size() * 4);
    dump(dir, "cb_boot.i64", g_boot.data(), g_boot.size() * 8);
    dump(dir, "cb_out.i32", g_out.data(), g_out.size() * 4);
    // decryption of the reference outputs (so the golden file also carries the expected plaintexts)
    for (int s = 0; s < NS; s++) {
        Torus64 ph = orc_lwe64Phase_lvl2(g_boot.data() + ((size_t)s * ell1) * (N2 + 1), K);
        printf("sample %d: lvl2 phase (w=0) = %lld (mu = %lld)\n", s, (long long)ph, (long long)(UINT64_C(1) << 56));
    }
    // ---------------- gate path: oracle restatement, spqlios backend vs portable backend (no compiled reference exists)
    {
        orc_gate_params gp; orc_gate_params_default(&gp);
        orc_gate_keys* G = orc_gate_keygen(&gp, 42);      // keygen is backend independent except bkFFT
        orc_gate_keys_rebuild_fft(G);
        orc_rng rg; orc_rng_seed(&rg, 43);
        const int NG = 16; int bad_dec = 0;
        std::vector<int32_t> ca(NG * (gp.n + 1)), cb(NG * (gp.n + 1)), out(NG * (gp.n + 1));
        for (int g = 0; g < NG; g++) { orc_bootsSymEncrypt(ca.data() + g * (gp.n + 1), g & 1, G, &rg); orc_bootsSymEncrypt(cb.data() + g * (gp.n + 1), (g >> 1) & 1, G, &rg); }
        orc_bootsGate_batch(out.data(), ORC_NAND, ca.data(), cb.data(), G, NG, 0);
        for (int g = 0; g < NG; g++) bad_dec += orc_bootsSymDecrypt(out.data() + g * (gp.n + 1), G) != !((g & 1) && ((g >> 1) & 1));
        PIN(bad_dec == 0, "gate path on spqlios backend: 16 NAND gates decrypt correctly");
        dump(dir, "gate_ca.i32", ca.data(), ca.size() * 4); dump(dir, "gate_cb.i32", cb.data(), cb.size() * 4);
        dump(dir, "gate_nand_spqlios.i32", out.data(), out.size() * 4);
        orc_gate_keys_free(G);
    }
    orc_set_fft_backend(nullptr);
    // gate-path function bodies and the high-precision FFT of the reference, compiled in place (ref_pins.cpp)
    g_fail += ref_run_pins(dir.c_str());
    printf("%s\n", g_fail ? "GOLDEN: FAILURES" : "GOLDEN: all pins hold");
    return g_fail ? 1 : 0;
}

static int cmd_bench_gate(int count, int threads, int reps) {
    orc_set_fft_backend(&kSpqlios);
    orc_gate_params gp; orc_gate_params_default(&gp);
    orc_gate_keys* G = orc_gate_keygen(&gp, 42);
    orc_rng rg; orc_rng_seed(&rg, 44);
    const size_t s = gp.n + 1;
    std::vector<int32_t> ca(count * s), cb(count * s), out(count * s);
    for (auto& v : ca) v = (int32_t)orc_rng_u64(&rg);
    for (auto& v : cb) v = (int32_t)orc_rng_u64(&rg);
    if (threads <= 0) threads = omp_get_max_threads();
    orc_bootsGate_batch(out.data(), ORC_NAND, ca.data(), cb.data(), G, threads < count ? threads : count, threads);   // warm-up
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        orc_bootsGate_batch(out.data(), ORC_NAND, ca.data(), cb.data(), G, count, threads);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    printf("{\"bench\": \"gate\", \"count\": %d, \"threads\": %d, \"seconds\": %.6f, \"gates_per_s\": %.3f, \"fft\": \"reference spqlios-fma\"}\n",
           count, threads, best, count / best);
    return 0;
}

static int cmd_bench_cb(int count) {
    orc_cb_params cp; orc_cb_params_default(&cp);
    orc_cb_keys* K = orc_cb_keygen(&cp, 42, 1);
    Globals* env = make_env(K, true);
    const int N1 = cp.N_lvl1;
    orc_rng rin; orc_rng_seed(&rin, 45);
    LweSample32 x(N1); orc_lwe32Encrypt_lvl1(x.a, (Torus32)(1u << 31), ldexp(1.0, -20), K, &rin);
    TGswSample32 res(cp.ell_lvl1, N1);
    tfhe_CircuitBootstrapFFT(&res, &x, env);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < count; i++) tfhe_CircuitBootstrapFFT(&res, &x, env);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"bench\": \"cb\", \"count\": %d, \"threads\": 1, \"seconds\": %.6f, \"cb_per_s\": %.4f}\n", count, dt, count / dt);
    return 0;
}

// circuit bootstraps on `threads` cores: the oracle's tfhe_CircuitBootstrapFFT (pinned bit for bit against the reference's own,
// cmd_golden) over the reference's spqlios kernels, OpenMP over independent samples -- the reference's only threading idiom
// (par/test_parallel_multiplications.cpp:62); its own entry point keeps shared scratch in Globals and cannot run on two threads.
static int cmd_bench_cb_mt(int count, int threads) {
    orc_set_fft_backend(&kSpqlios);
    orc_cb_params cp; orc_cb_params_default(&cp);
    orc_cb_keys* K = orc_cb_keygen(&cp, 42, 1);
    orc_cb_keys_rebuild_fft(K);
    const int N1 = cp.N_lvl1;
    const size_t in_s = N1 + 1, out_s = (size_t)2 * cp.ell_lvl1 * 2 * N1;
    if (threads <= 0) threads = omp_get_max_threads();
    std::vector<int32_t> in(count * in_s), out(count * out_s);
    orc_rng rg; orc_rng_seed(&rg, 45);
    for (auto& v : in) v = (int32_t)orc_rng_u64(&rg);
    auto pass = [&](int n) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int i = 0; i < n; i++) orc_tfhe_CircuitBootstrapFFT(out.data() + i * out_s, in.data() + i * in_s, K);
    };
    pass(threads < count ? threads : count);      // warm-up: per-thread FFT processors, page faults of the 2.7 GB key
    auto t0 = std::chrono::steady_clock::now();
    pass(count);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"bench\": \"cb\", \"count\": %d, \"threads\": %d, \"seconds\": %.6f, \"cb_per_s\": %.4f, \"fft\": \"reference spqlios-fma\"}\n",
           count, threads, dt, count / dt);
    return 0;
}
// 128-bit fixed-point transforms (hp/code.cpp restated, oracle/hpfft_oracle.c) on `threads` cores
static int cmd_bench_hp(int N, int count, int threads) {
    const int n = 2 * N;
    std::vector<orc_cplx96> om(n), ob(n);
    orc_hp_precomp_iFFT(om.data(), n); orc_hp_precomp_FFT(ob.data(), n);
    if (threads <= 0) threads = omp_get_max_threads();
    std::vector<int64_t> in((size_t)count * N), back((size_t)count * N);
    std::vector<orc_cplx96> spec((size_t)count * (N / 2));
    orc_rng rg; orc_rng_seed(&rg, 46);
    for (auto& v : in) v = (int64_t)orc_rng_u64(&rg);
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < count; i++) orc_hp_iFFT(spec.data() + (size_t)i * (N / 2), in.data() + (size_t)i * N, n, om.data());
    double t_i = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < count; i++) orc_hp_FFT(back.data() + (size_t)i * N, spec.data() + (size_t)i * (N / 2), n, ob.data());
    double t_f = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"bench\": \"hp\", \"N\": %d, \"count\": %d, \"threads\": %d, \"ifft_per_s\": %.2f, \"fft_per_s\": %.2f}\n", N, count, threads,
           count / t_i, count / t_f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && !strcmp(argv[1], "golden")) return cmd_golden(argv[2]);
    if (argc >= 4 && !strcmp(argv[1], "bench-gate")) return cmd_bench_gate(atoi(argv[2]), atoi(argv[3]), argc >= 5 ? atoi(argv[4]) : 1);
    if (argc >= 4 && !strcmp(argv[1], "bench-cb")) return cmd_bench_cb_mt(atoi(argv[2]), atoi(argv[3]));
    if (argc >= 3 && !strcmp(argv[1], "bench-cb")) return cmd_bench_cb(atoi(argv[2]));
    if (argc >= 5 && !strcmp(argv[1], "bench-hp")) return cmd_bench_hp(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    fprintf(stderr, "usage: ref_harness golden <dir> | bench-gate <count> <threads> [reps] | bench-cb <count> [threads] | bench-hp <N> <count> <threads>\n");
    return 2;
}
