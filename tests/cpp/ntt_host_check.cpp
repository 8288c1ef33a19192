// ntt_host_check.cpp -- CPU check of experimental-tfhe_b200/csrc/exact_ntt.cuh (the header is __host__ __device__: the same field arithmetic and
// butterflies the CUDA kernels run).  Compares (1) gl_mul / gl_add / gl_sub with 128-bit integer arithmetic, (2) the limb-split
// NTT external-product recipe with a schoolbook negacyclic product mod 2^64.   g++ -O2 -std=c++17 -I<csrc> ntt_host_check.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include "exact_ntt.cuh"
using namespace tfhe_b200;
typedef unsigned __int128 u128;

int main() {
    std::mt19937_64 rng(12345);
    int bad = 0;
    const uint64_t edge[] = {0, 1, 2, GL_EPS, GL_EPS + 1, GL_P - 1, GL_P - 2, 1ull << 63, (1ull << 63) - 1, 0xFFFFFFFF00000000ull, 0x00000000FFFFFFFFull};
    std::vector<uint64_t> vals(edge, edge + sizeof(edge) / 8);
    for (int i = 0; i < 2000; i++) vals.push_back(rng() % GL_P);
    for (size_t i = 0; i < vals.size(); i++)
        for (size_t j = 0; j < vals.size(); j += (i < 11 ? 1 : 97)) {
            const uint64_t a = vals[i], b = vals[j];
            bad += gl_mul(a, b) != (uint64_t)(((u128)a * b) % GL_P);
            bad += gl_add(a, b) != (uint64_t)(((u128)a + b) % GL_P);
            bad += gl_sub(a, b) != (uint64_t)(((u128)a + GL_P - b) % GL_P);
        }
    const uint64_t his[] = {0, 1, GL_EPS, ~(uint64_t)0, (uint64_t)1 << 63, 0xFFFFFFFF00000000ull};
    const uint64_t los[] = {0, 1, GL_EPS, ~(uint64_t)0, GL_P, GL_P - 1};
    for (uint64_t hi : his)
        for (uint64_t lo : los)
            bad += gl_reduce128(hi, lo) != (uint64_t)((((u128)hi << 64) | lo) % GL_P);
    printf("field arithmetic: %s\n", bad ? "MISMATCH" : "ok");
    for (int logN : {10, 11}) {
        const int N = 1 << logN;
        std::vector<uint64_t> psi(N), psi_inv(N); uint64_t n_inv;
        gl_make_tables(logN, psi.data(), psi_inv.data(), &n_inv);
        bad += gl_mul(gl_mul(psi[1], psi[1]), 1) != GL_P - 1;      // psi_rev[1] = psi^(N/2), its square = psi^N = -1
        // round trip
        std::vector<uint64_t> a(N), keep;
        for (auto& v : a) v = rng() % GL_P;
        keep = a;
        gl_ntt_forward_serial(a.data(), psi.data(), N);
        gl_ntt_inverse_serial(a.data(), psi_inv.data(), n_inv, N);
        bad += a != keep;
        // external-product recipe: sum_p d_p (*) t_p mod (X^N+1, 2^64), digits in [-512, 511], KPL polynomials
        const int KPL = 12;
        std::vector<uint64_t> acc_lo(N, 0), acc_hi(N, 0);
        std::vector<uint64_t> exact(N, 0);
        for (int p = 0; p < KPL; p++) {
            std::vector<int64_t> d(N); std::vector<uint64_t> t(N);
            for (auto& v : d) v = (int64_t)(rng() % 1024) - 512;
            for (auto& v : t) v = rng();
            if (p == 0) { t[0] = ~0ull; t[1] = 1ull << 63; d[0] = -512; d[1] = 511; }
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    const uint64_t prod = (uint64_t)d[i] * t[j];
                    if (i + j < N) exact[i + j] += prod; else exact[i + j - N] -= prod;
                }
            std::vector<uint64_t> D(N), TL(N), TH(N);
            for (int i = 0; i < N; i++) { D[i] = gl_from_i64(d[i]); TL[i] = t[i] & GL_EPS; TH[i] = t[i] >> 32; }
            gl_ntt_forward_serial(D.data(), psi.data(), N);
            gl_ntt_forward_serial(TL.data(), psi.data(), N);
            gl_ntt_forward_serial(TH.data(), psi.data(), N);
            for (int i = 0; i < N; i++) { acc_lo[i] = gl_add(acc_lo[i], gl_mul(D[i], TL[i])); acc_hi[i] = gl_add(acc_hi[i], gl_mul(D[i], TH[i])); }
        }
        gl_ntt_inverse_serial(acc_lo.data(), psi_inv.data(), n_inv, N);
        gl_ntt_inverse_serial(acc_hi.data(), psi_inv.data(), n_inv, N);
        int mism = 0;
        for (int i = 0; i < N; i++) {
            const uint64_t r = (uint64_t)gl_lift(acc_lo[i]) + ((uint64_t)gl_lift(acc_hi[i]) << 32);
            mism += r != exact[i];
        }
        printf("N=%d: NTT external product vs schoolbook mod 2^64: %d mismatches\n", N, mism);
        bad += mism;
    }
    printf("%s\n", bad ? "NTT HOST CHECK: FAILED" : "NTT HOST CHECK: all ok");
    return bad != 0;
}
