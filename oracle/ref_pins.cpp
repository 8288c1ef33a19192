// ref_pins.cpp -- test infrastructure: pins the oracle's gate-path and high-precision restatements against the REFERENCE's own
// function bodies, compiled in place (oracle/Makefile: gate_extract_{scalar,avx2}.cpp cut out of cb/*_functions.cpp by line range
// behind ref_gate_shim.h; hp_patched.cpp = hp/code.cpp through patch_hp.sed).  Called from `ref_harness golden`; every line it
// prints ends up in tests/golden/pin_log.txt.  Separate translation unit because the shim's struct names collide with poc_types.h.
#include "ref_gate_shim.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
extern "C" {
#include "tfhe_oracle.h"
}

#define DECL_GATE(NS)                                                                                                         \
    namespace NS {                                                                                                            \
    int modSwitchFromTorus32(Torus32 phase, int Msize);                                                                       \
    Torus32 modSwitchToTorus32(int mu, int Msize);                                                                            \
    void torusPolynomialMulByXaiMinusOne(TorusPolynomial* result, int a, const TorusPolynomial* source);                       \
    void torusPolynomialMulByXai(TorusPolynomial* result, int a, const TorusPolynomial* source);                               \
    void tGswTorus32PolynomialDecompH(IntPolynomial* result, const TorusPolynomial* sample, const TGswParams* params);         \
    void tLweExtractLweSampleIndex(LweSample* result, const TLweSample* x, const int index, const LweParams* params,           \
                                   const TLweParams* rparams);                                                                \
    void lweKeySwitch(LweSample* result, const LweKeySwitchKey* ks, const LweSample* sample);                                  \
    }
DECL_GATE(refgate_scalar)
DECL_GATE(refgate_avx2)
extern "C" void hpref_precomp(uint64_t* powomega, uint64_t* powombar, int n);
extern "C" void hpref_iFFT(uint64_t* out, const int64_t* in, int n);
extern "C" void hpref_FFT(int64_t* out, uint64_t* in, int n);

static int g_bad = 0;
#define PIN(cond, msg) do { if (!(cond)) { printf("PIN FAILED: %s\n", msg); g_bad++; } else printf("pinned: %s\n", msg); } while (0)

template <typename F> static void for_both(F f) { f(0); f(1); }

static void dump(const std::string& dir, const char* name, const void* data, size_t bytes) {
    std::string path = dir + "/" + name;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(data, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fclose(f);
}

extern "C" int ref_run_pins(const char* golden_dir) {
    orc_rng r; orc_rng_seed(&r, 777);
    // ---------------------------------------------------------------- modulus switch, both directions  (cb/numeric_functions.cpp:54-67)
    {
        int bad = 0;
        const Torus32 edge[] = {0, -1, 1, INT32_MAX, INT32_MIN, 1 << 20, (1 << 20) - 1, -(1 << 20), 1 << 21, 0x7FEFFFFF, (Torus32)0x80100000};
        for (int Msize : {2048, 4096, 8, 16}) {
            for (Torus32 e : edge) bad += refgate_scalar::modSwitchFromTorus32(e, Msize) != orc_modSwitchFromTorus32(e, Msize);
            for (int i = 0; i < 20000; i++) {
                const Torus32 x = (Torus32)orc_rng_u64(&r);
                bad += refgate_scalar::modSwitchFromTorus32(x, Msize) != orc_modSwitchFromTorus32(x, Msize);
                bad += refgate_avx2::modSwitchFromTorus32(x, Msize) != orc_modSwitchFromTorus32(x, Msize);
            }
            for (int mu = -Msize; mu <= Msize; mu++) bad += refgate_scalar::modSwitchToTorus32(mu, Msize) != orc_modSwitchToTorus32(mu, Msize);
        }
        PIN(bad == 0, "modSwitchFromTorus32 / modSwitchToTorus32 == reference function bodies (cb/numeric_functions.cpp:54-67), bit-exact");
    }
    // ---------------------------------------------------------------- negacyclic monomial products  (:304-347)
    {
        int bad = 0;
        for (int N : {1024, 2048}) {
            std::vector<Torus32> in(N), o_ref(N), o_orc(N);
            for (auto& v : in) v = (Torus32)orc_rng_u64(&r);
            TorusPolynomial src{N, in.data()}, dst{N, o_ref.data()};
            std::vector<int> as = {0, 1, N - 1, N, N + 1, 2 * N - 1, 777, N + 333};
            for (int a : as) {
                refgate_scalar::torusPolynomialMulByXaiMinusOne(&dst, a, &src);
                orc_torusPolynomialMulByXaiMinusOne(o_orc.data(), a, in.data(), N);
                bad += memcmp(o_ref.data(), o_orc.data(), sizeof(Torus32) * N) != 0;
                refgate_avx2::torusPolynomialMulByXai(&dst, a, &src);
                orc_torusPolynomialMulByXai(o_orc.data(), a, in.data(), N);
                bad += memcmp(o_ref.data(), o_orc.data(), sizeof(Torus32) * N) != 0;
            }
        }
        PIN(bad == 0, "torusPolynomialMulByXaiMinusOne / torusPolynomialMulByXai == reference function bodies (:304-347), bit-exact");
    }
    // ---------------------------------------------------------------- gadget decomposition  (cb/tgsw_functions.cpp:224-337; offset :30-36)
    {
        int bad = 0;
        const int N = 1024;
        const int sets[4][2] = {{2, 10}, {3, 8}, {1, 10}, {4, 8}};
        for (auto& s : sets) {
            const int l = s[0], Bgbit = s[1];
            TLweParams tp{N, 1, 0, 0, {N, 0, 0}};
            TGswParams gp{};
            gp.l = l; gp.Bgbit = Bgbit; gp.Bg = 1 << Bgbit; gp.halfBg = gp.Bg / 2; gp.maskMod = gp.Bg - 1; gp.tlwe_params = &tp; gp.kpl = 2 * l;
            uint32_t temp1 = 0;                                     // TGswParams constructor, cb/tgsw_functions.cpp:30-36
            for (int i = 0; i < l; ++i) temp1 += 1u << (32 - (i + 1) * Bgbit);
            gp.offset = temp1 * (uint32_t)gp.halfBg;
            bad += gp.offset != orc_tgsw32_offset(l, Bgbit);
            std::vector<Torus32> in(N), keep;
            for (auto& v : in) v = (Torus32)orc_rng_u64(&r);
            in[0] = 0; in[1] = -1; in[2] = INT32_MAX; in[3] = INT32_MIN; in[4] = (Torus32)(0u - gp.offset); in[5] = (Torus32)(1u << (31 - l * Bgbit));
            keep = in;
            std::vector<int32_t> o_orc((size_t)l * N);
            orc_tGswTorus32PolynomialDecompH(o_orc.data(), in.data(), N, l, Bgbit);
            for (int which = 0; which < 2; which++) {
                std::vector<std::vector<int>> rows(l, std::vector<int>(N));
                std::vector<IntPolynomial> res(l);
                for (int p = 0; p < l; p++) res[p] = IntPolynomial{N, rows[p].data()};
                TorusPolynomial src{N, in.data()};
                if (which == 0) refgate_scalar::tGswTorus32PolynomialDecompH(res.data(), &src, &gp);
                else            refgate_avx2::tGswTorus32PolynomialDecompH(res.data(), &src, &gp);
                for (int p = 0; p < l; p++) bad += memcmp(rows[p].data(), o_orc.data() + (size_t)p * N, sizeof(int32_t) * N) != 0;
                bad += in != keep;                                  // the reference restores its input (offset added then removed)
            }
        }
        PIN(bad == 0, "tGswTorus32PolynomialDecompH (scalar and AVX2 branches) == oracle, 4 gadget sets (cb/tgsw_functions.cpp:224-337,30-36), bit-exact");
    }
    // ---------------------------------------------------------------- sample extraction  (cb/tlwe_functions.cpp:351-363)
    {
        int bad = 0;
        const int N = 1024;
        std::vector<Torus32> tl(2 * N), o_ref(N + 1), o_orc(N + 1);
        for (auto& v : tl) v = (Torus32)orc_rng_u64(&r);
        TorusPolynomial polys[2] = {{N, tl.data()}, {N, tl.data() + N}};
        TLweSample x{polys, &polys[1], 0., 1};
        LweParams lp{N, 0, 0};
        TLweParams tp{N, 1, 0, 0, lp};
        LweSample res{o_ref.data(), 0, 0.};
        refgate_scalar::tLweExtractLweSampleIndex(&res, &x, 0, &lp, &tp);
        o_ref[N] = res.b;
        orc_tLweExtractLweSample(o_orc.data(), tl.data(), N);
        bad += memcmp(o_ref.data(), o_orc.data(), sizeof(Torus32) * (N + 1)) != 0;
        PIN(bad == 0, "tLweExtractLweSampleIndex(index 0) == oracle (cb/tlwe_functions.cpp:351-363), bit-exact");
    }
    // ---------------------------------------------------------------- key switch  (cb/lwe_functions.cpp:136-151,163-171)
    {
        int bad = 0;
        const int cfg[3][4] = {{64, 37, 8, 2}, {48, 20, 16, 1}, {40, 33, 5, 3}};          // n_in, n_out, t, basebit
        for (auto& c : cfg) {
            const int n_in = c[0], n_out = c[1], t = c[2], basebit = c[3], base = 1 << basebit;
            std::vector<Torus32> ks((size_t)n_in * t * base * (n_out + 1));
            for (auto& v : ks) v = (Torus32)orc_rng_u64(&r);
            // reference view of the same array: ks[i][j][d] = LweSample over row (i,j,d)
            std::vector<LweSample> raw((size_t)n_in * t * base);
            std::vector<LweSample*> l1((size_t)n_in * t);
            std::vector<LweSample**> l2(n_in);
            for (size_t e = 0; e < raw.size(); e++) { Torus32* row = ks.data() + e * (n_out + 1); raw[e] = LweSample{row, row[n_out], 0.}; }
            for (size_t e = 0; e < l1.size(); e++) l1[e] = raw.data() + e * base;
            for (int i = 0; i < n_in; i++) l2[i] = l1.data() + (size_t)i * t;
            LweParams outp{n_out, 0, 0};
            LweKeySwitchKey K{n_in, t, basebit, base, &outp, raw.data(), l1.data(), l2.data()};
            for (int rep = 0; rep < 8; rep++) {
                std::vector<Torus32> x(n_in + 1), o_ref(n_out + 1), o_orc(n_out + 1);
                for (auto& v : x) v = (Torus32)orc_rng_u64(&r);
                if (rep == 0) { x[0] = 0; x[1] = -1; x[2] = INT32_MAX; x[3] = INT32_MIN; }
                LweSample smp{x.data(), x[n_in], 0.}, res{o_ref.data(), 0, 0.};
                refgate_scalar::lweKeySwitch(&res, &K, &smp);
                o_ref[n_out] = res.b;
                orc_lweKeySwitch(o_orc.data(), ks.data(), x.data(), n_in, n_out, t, basebit);
                bad += memcmp(o_ref.data(), o_orc.data(), sizeof(Torus32) * (n_out + 1)) != 0;
            }
        }
        PIN(bad == 0, "lweKeySwitch / lweKeySwitchTranslate_fromArray == oracle at bases 2, 4, 8 (cb/lwe_functions.cpp:136-171), bit-exact");
    }
    // ---------------------------------------------------------------- 128-bit fixed-point FFT  (hp/code.cpp compiled from a patched copy)
    for (int N : {2048, 4096}) {
        const int n = 2 * N;
        std::vector<uint64_t> om_ref((size_t)4 * n), ob_ref((size_t)4 * n);
        std::vector<orc_cplx96> om(n), ob(n);
        hpref_precomp(om_ref.data(), ob_ref.data(), n);
        orc_hp_precomp_iFFT(om.data(), n); orc_hp_precomp_FFT(ob.data(), n);
        char msg[200];
        snprintf(msg, sizeof msg, "hp twiddle tables (2N = %d entries, both directions) == reference precomp_iFFT / precomp_FFT (hp/code.cpp:376-388), bit-exact", n);
        PIN(memcmp(om_ref.data(), om.data(), (size_t)32 * n) == 0 && memcmp(ob_ref.data(), ob.data(), (size_t)32 * n) == 0, msg);
        int bad_i = 0, bad_f = 0;
        const int NP = 4;
        std::vector<int64_t> in((size_t)NP * N), back_ref((size_t)NP * N), back_orc(N);
        std::vector<uint64_t> spec_ref((size_t)NP * (N / 2) * 4);
        std::vector<orc_cplx96> spec_orc(N / 2), tmp(N / 2);
        for (auto& v : in) v = (int64_t)orc_rng_u64(&r);
        for (int j = 0; j < N; j++) in[(size_t)1 * N + j] = 0;
        const int64_t edge[8] = {INT64_MAX, INT64_MIN, -1, 1, 0, (int64_t)1 << 62, -((int64_t)1 << 62), 12345};
        for (int j = 0; j < 8; j++) in[(size_t)2 * N + j] = edge[j];
        for (int p = 0; p < NP; p++) {
            hpref_iFFT(spec_ref.data() + (size_t)p * (N / 2) * 4, in.data() + (size_t)p * N, n);
            orc_hp_iFFT(spec_orc.data(), in.data() + (size_t)p * N, n, om.data());
            bad_i += memcmp(spec_orc.data(), spec_ref.data() + (size_t)p * (N / 2) * 4, (size_t)32 * (N / 2)) != 0;
            std::vector<uint64_t> clob(spec_ref.begin() + (size_t)p * (N / 2) * 4, spec_ref.begin() + (size_t)(p + 1) * (N / 2) * 4);
            hpref_FFT(back_ref.data() + (size_t)p * N, clob.data(), n);
            memcpy(tmp.data(), spec_orc.data(), (size_t)32 * (N / 2));
            orc_hp_FFT(back_orc.data(), tmp.data(), n, ob.data());
            if (N == 2048) bad_f += memcmp(back_orc.data(), back_ref.data() + (size_t)p * N, sizeof(int64_t) * N) != 0;
            else {
                // The reference divides by N/2 with a literal `>>10` (:502-503, "divide by N/2"), right for the only N it runs (2048).
                // The oracle shifts by log2(N/2) = 11 here, so it keeps bits [11,75) of the same 128-bit value where the unmodified
                // reference keeps [10,74): they share 63 bits, compared here (the top bit is covered by the N = 2048 pin).
                for (int j = 0; j < N; j++)
                    bad_f += ((uint64_t)back_orc[j] & (UINT64_MAX >> 1)) != ((uint64_t)back_ref[(size_t)p * N + j] >> 1);
            }
        }
        snprintf(msg, sizeof msg, "hp iFFT N=%d == reference iFFT (hp/code.cpp:391-443), 4 polynomials incl. zero and extreme inputs, bit-exact", N);
        PIN(bad_i == 0, msg);
        if (N == 2048) snprintf(msg, sizeof msg, "hp FFT N=%d == reference FFT (hp/code.cpp:446-512), bit-exact", N);
        else snprintf(msg, sizeof msg, "hp FFT N=%d == reference FFT (hp/code.cpp:446-512) on the 63 bits shared with its hard-coded >>10 (:502-503)", N);
        PIN(bad_f == 0, msg);
        if (golden_dir) {
            char name[64];
            snprintf(name, sizeof name, "hp_in_N%d.i64", N); dump(golden_dir, name, in.data(), in.size() * 8);
            snprintf(name, sizeof name, "hp_spec_N%d.u64", N); dump(golden_dir, name, spec_ref.data(), spec_ref.size() * 8);
            snprintf(name, sizeof name, "hp_back_N%d.i64", N); dump(golden_dir, name, back_ref.data(), back_ref.size() * 8);      // reference output as is (N = 4096: >>10)
        }
    }
    return g_bad;
}
