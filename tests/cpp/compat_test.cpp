// compat_test.cpp -- drives the C++ host shim (experimental-tfhe_b200/host/tfhe_b200_compat.hpp) through the reference's own
// function names on reference-shaped structs, and checks every result against the oracle (tests only link the oracle).
// Reads like the reference's usage: keygen -> init_LweBootstrappingKeyFFT -> bootsNAND / tfhe_bootstrap_FFT / lweKeySwitch ...
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../experimental-tfhe_b200/host/tfhe_b200_compat.hpp"
extern "C" {
#include "../../oracle/tfhe_oracle.h"
}
using namespace tfhe_b200_compat;

static int g_fail = 0;
#define CHECK(cond, msg) do { if (!(cond)) { printf("FAIL: %s\n", msg); g_fail++; } else printf("ok: %s\n", msg); } while (0)

int main() {
    // ---------------- library-style gate path
    orc_gate_params gp; orc_gate_params_default(&gp);
    orc_gate_keys* K = orc_gate_keygen(&gp, 42);
    LweParams in_out{gp.n, gp.ks_stdev, 0.012467};
    TLweParams accum{gp.N, gp.k, gp.bk_stdev, 3.73e-9, LweParams{gp.N, gp.bk_stdev, 3.73e-9}};
    TGswParams bkp(gp.bk_l, gp.bk_Bgbit, &accum);
    // fill the reference-shaped key structs from the oracle's flat key material
    LweKeySwitchKey ks(gp.N, gp.ks_t, gp.ks_basebit, &in_out);
    memcpy(ks.ks0_raw, K->ks, sizeof(Torus32) * (size_t)gp.N * gp.ks_t * (1 << gp.ks_basebit) * (gp.n + 1));
    std::vector<TGswSample*> bks(gp.n);
    for (int i = 0; i < gp.n; i++) {
        bks[i] = new TGswSample(&bkp);
        for (int p = 0; p < bkp.kpl; p++)
            for (int q = 0; q < 2; q++)
                memcpy(bks[i]->all_sample[p].a[q].coefsT, K->bk + ((((size_t)i * bkp.kpl + p) * 2) + q) * gp.N, sizeof(Torus32) * gp.N);
    }
    LweBootstrappingKey bk{&in_out, &bkp, &accum, &accum.extracted_lweparams, bks.data(), &ks};
    LweBootstrappingKeyFFT bkFFT;
    init_LweBootstrappingKeyFFT(&bkFFT, &bk);
    TFheGateBootstrappingCloudKeySet cloud{&bkFFT};

    orc_rng r; orc_rng_seed(&r, 7);
    const Torus32 MU = orc_modSwitchToTorus32(1, 8);
    int bad = 0;
    for (int t = 0; t < 8; t++) {
        const int a = t & 1, b = (t >> 1) & 1, c = (t >> 2) & 1;
        LweSample ca(&in_out), cb(&in_out), cc(&in_out), res(&in_out);
        std::vector<Torus32> tmp(gp.n + 1);
        orc_bootsSymEncrypt(tmp.data(), a, K, &r); unflatten(&ca, gp.n, tmp.data());
        orc_bootsSymEncrypt(tmp.data(), b, K, &r); unflatten(&cb, gp.n, tmp.data());
        orc_bootsSymEncrypt(tmp.data(), c, K, &r); unflatten(&cc, gp.n, tmp.data());
        bootsNAND(&res, &ca, &cb, &cloud); flatten(&res, gp.n, tmp.data());
        bad += orc_bootsSymDecrypt(tmp.data(), K) != !(a && b);
        bootsXOR(&res, &ca, &cb, &cloud); flatten(&res, gp.n, tmp.data());
        bad += orc_bootsSymDecrypt(tmp.data(), K) != (a ^ b);
        bootsMUX(&res, &ca, &cb, &cc, &cloud); flatten(&res, gp.n, tmp.data());
        bad += orc_bootsSymDecrypt(tmp.data(), K) != (a ? b : c);
        bootsNOT(&res, &ca, &cloud); flatten(&res, gp.n, tmp.data());
        bad += orc_bootsSymDecrypt(tmp.data(), K) != !a;
        tfhe_bootstrap_FFT(&res, &bkFFT, MU, &ca); flatten(&res, gp.n, tmp.data());
        bad += orc_bootsSymDecrypt(tmp.data(), K) != a;
    }
    CHECK(bad == 0, "bootsNAND / bootsXOR / bootsMUX / bootsNOT / tfhe_bootstrap_FFT decrypt correctly through the shim");

    {   // lweKeySwitch: bit-exact
        LweSample u(&accum.extracted_lweparams), res(&in_out);
        std::vector<Torus32> in(gp.N + 1), ref(gp.n + 1), got(gp.n + 1);
        for (auto& v : in) v = (Torus32)orc_rng_u64(&r);
        unflatten(&u, gp.N, in.data());
        lweKeySwitch(&res, &ks, &u); flatten(&res, gp.n, got.data());
        orc_lweKeySwitch(ref.data(), K->ks, in.data(), gp.N, gp.n, gp.ks_t, gp.ks_basebit);
        CHECK(got == ref, "lweKeySwitch bit-exact vs oracle");
    }
    {   // tfhe_MuxRotate_FFT: one CMUX within 1 LSB of the exact external product
        TLweSample acc(&accum), res(&accum);
        std::vector<Torus32> flat(2 * gp.N), tmp(2 * gp.N);
        for (auto& v : flat) v = (Torus32)orc_rng_u64(&r);
        for (int q = 0; q < 2; q++) memcpy(acc.a[q].coefsT, flat.data() + q * gp.N, sizeof(Torus32) * gp.N);
        const int i = 123, barai = 1500;
        tfhe_MuxRotate_FFT(&res, &acc, bkFFT.bkFFT + i, barai, &bkp);          // the reference's own signature (cb/lwe_functions.cpp:328)
        for (int q = 0; q < 2; q++) orc_torusPolynomialMulByXaiMinusOne(tmp.data() + q * gp.N, barai, flat.data() + q * gp.N, gp.N);
        orc_tGswExternMulToTLwe(tmp.data(), K->bk + (size_t)i * bkp.kpl * 2 * gp.N, gp.N, gp.bk_l, gp.bk_Bgbit);
        int worst = 0;
        for (int q = 0; q < 2; q++)
            for (int j = 0; j < gp.N; j++) {
                int32_t exact = (int32_t)((uint32_t)tmp[q * gp.N + j] + (uint32_t)flat[q * gp.N + j]);
                int d = abs((int32_t)((uint32_t)res.a[q].coefsT[j] - (uint32_t)exact));
                if (d > worst) worst = d;
            }
        CHECK(worst <= 1, "tfhe_MuxRotate_FFT within 1 LSB of the exact external product");
        // tfhe_blindRotate_FFT on a SUB-RANGE of the key, as the reference calls it with bkFFT+i and a shorter n (:352): two steps at
        // i, i+1 must equal two MuxRotate steps in a row
        TLweSample two(&accum), ref1(&accum), ref2(&accum);
        for (int q = 0; q < 2; q++) memcpy(two.a[q].coefsT, flat.data() + q * gp.N, sizeof(Torus32) * gp.N);
        const int bara2[2] = {barai, 77};
        tfhe_blindRotate_FFT(&two, bkFFT.bkFFT + i, bara2, 2, &bkp);
        tfhe_MuxRotate_FFT(&ref1, &acc, bkFFT.bkFFT + i, bara2[0], &bkp);
        tfhe_MuxRotate_FFT(&ref2, &ref1, bkFFT.bkFFT + i + 1, bara2[1], &bkp);
        bool same = true;
        for (int q = 0; q < 2; q++) same = same && memcmp(two.a[q].coefsT, ref2.a[q].coefsT, sizeof(Torus32) * gp.N) == 0;
        CHECK(same, "tfhe_blindRotate_FFT(bkFFT+i, n=2) == two tfhe_MuxRotate_FFT steps");
    }
    {   // tfhe_bootstrap_woKS_FFT + tfhe_blindRotateAndExtract_FFT: phase == +-mu up to bootstrapping noise
        LweSample x(&in_out), u(&accum.extracted_lweparams);
        std::vector<Torus32> tmp(gp.n + 1), out(gp.N + 1);
        orc_bootsSymEncrypt(tmp.data(), 1, K, &r); unflatten(&x, gp.n, tmp.data());
        tfhe_bootstrap_woKS_FFT(&u, &bkFFT, MU, &x); flatten(&u, gp.N, out.data());
        int32_t ph = orc_lwePhase(out.data(), K->tlwe_key, gp.N);
        CHECK(abs(ph - MU) < (1 << 28), "tfhe_bootstrap_woKS_FFT phase");
        TorusPolynomial v(gp.N); for (int j = 0; j < gp.N; j++) v.coefsT[j] = MU;
        std::vector<int> bara(gp.n);
        for (int i = 0; i < gp.n; i++) bara[i] = orc_modSwitchFromTorus32(tmp[i], 2 * gp.N);
        const int barb = orc_modSwitchFromTorus32(tmp[gp.n], 2 * gp.N);
        tfhe_blindRotateAndExtract_FFT(&u, &v, bkFFT.bkFFT, barb, bara.data(), gp.n, &bkp); flatten(&u, gp.N, out.data());
        ph = orc_lwePhase(out.data(), K->tlwe_key, gp.N);
        CHECK(abs(ph - MU) < (1 << 28), "tfhe_blindRotateAndExtract_FFT phase");
    }
    {   // gate-level circuit: 6 instances of a 4-bit ripple-carry adder through tfhe_b200_circuit_eval_batch
        const int bits = 4, B = 6, row = gp.n + 1;
        AdderNetlist nl = ripple_carry_adder_netlist(bits);
        std::vector<Torus32> wires((size_t)nl.n_wires * B * row, 0);
        int xa[B] = {0, 15, 7, 9, 5, 12}, xb[B] = {0, 1, 8, 9, 10, 15}, xc[B] = {0, 1, 0, 1, 1, 0};
        for (int k = 0; k < B; k++) {
            for (int i = 0; i < bits; i++) {
                orc_bootsSymEncrypt(&wires[((size_t)(nl.a0 + i) * B + k) * row], (xa[k] >> i) & 1, K, &r);
                orc_bootsSymEncrypt(&wires[((size_t)(nl.b0 + i) * B + k) * row], (xb[k] >> i) & 1, K, &r);
            }
            orc_bootsSymEncrypt(&wires[((size_t)nl.cin * B + k) * row], xc[k], K, &r);
        }
        Torus32* d = nullptr;
        cudaMalloc(&d, wires.size() * sizeof(Torus32));
        cudaMemcpy(d, wires.data(), wires.size() * sizeof(Torus32), cudaMemcpyHostToDevice);
        int rc = tfhe_b200_circuit_eval_batch(bkFFT.engine, nl.gates.data(), (int)nl.gates.size(), d, nl.n_wires, B, nullptr);
        cudaDeviceSynchronize();
        cudaMemcpy(wires.data(), d, wires.size() * sizeof(Torus32), cudaMemcpyDeviceToHost);
        cudaFree(d);
        int wrong = rc != 0;
        for (int k = 0; k < B; k++) {
            int sum = 0;
            for (int i = 0; i < bits; i++) sum |= orc_bootsSymDecrypt(&wires[((size_t)(nl.s0 + i) * B + k) * row], K) << i;
            sum |= orc_bootsSymDecrypt(&wires[((size_t)(nl.c0 + bits) * B + k) * row], K) << bits;
            wrong += sum != xa[k] + xb[k] + xc[k];
        }
        CHECK(wrong == 0, "4-bit ripple-carry adders (ripple_carry_adder_netlist + tfhe_b200_circuit_eval_batch) add correctly");
    }
    destroy_LweBootstrappingKeyFFT(&bkFFT);
    for (auto* p : bks) delete p;
    orc_gate_keys_free(K);

    // ---------------- proof-of-concept style circuit bootstrap (without the 2.7 GB private key-switch key: stage functions only)
    orc_cb_params cp; orc_cb_params_default(&cp);
    orc_cb_keys* C = orc_cb_keygen(&cp, 42, 0);
    tfhe_b200_cb_params ep{cp.n_lvl0, cp.N_lvl1, cp.N_lvl2, cp.bgbit_lvl1, cp.ell_lvl1, cp.bgbit_lvl2, cp.ell_lvl2,
                           cp.kslength_lvl10, cp.ksbasebit_lvl10, cp.kslength_lvl21, cp.ksbasebit_lvl21};
    {
        Globals env(ep, C->preKS, C->bk, nullptr);
        LweSample32 in(cp.N_lvl1), pre(cp.n_lvl0);
        orc_lwe32Encrypt_lvl1(in.a, (Torus32)(1u << 31), ldexp(1.0, -20), C, &r);
        preKeySwitch(&pre, &in, &env);
        std::vector<Torus32> ref(cp.n_lvl0 + 1);
        orc_preKeySwitch(ref.data(), in.a, C);
        CHECK(memcmp(ref.data(), pre.a, sizeof(Torus32) * (cp.n_lvl0 + 1)) == 0, "preKeySwitch bit-exact vs oracle");
        std::vector<int> ms(cp.n_lvl0 + 1), ms_ref(cp.n_lvl0 + 1);
        preModSwitch(ms.data(), &pre, &env);
        orc_preModSwitch(ms_ref.data(), ref.data(), cp.n_lvl0, cp.N_lvl2);
        CHECK(ms == ms_ref, "preModSwitch bit-exact vs oracle");
        LweSample64 boot(cp.N_lvl2);
        const Torus64 mu = (Torus64)1 << 56;
        circuitBootstrapWoKS(&boot, mu, ms.data(), &env);
        const Torus64 ph = orc_lwe64Phase_lvl2(boot.a, C);
        CHECK(llabs(ph - mu) < ((Torus64)1 << 44), "circuitBootstrapWoKS phase == mu for an input of 1/2");
    }
    orc_cb_keys_free(C);
    printf("%s\n", g_fail ? "COMPAT: FAILURES" : "COMPAT: all checks passed");
    return g_fail ? 1 : 0;
}
