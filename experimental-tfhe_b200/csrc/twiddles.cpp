// twiddles.cpp -- host-side twiddle tables for the device transforms.
//
// FP64 tables follow FftPlan<LOGM> (fft_device.cuh): twist[M] | tw1[7][M/8] | tw2[7][M/64] | tw3[7][M/512].
// The reference builds its tables with quadrant-reduced cos/sin in double (cb/spqlios/spqlios-fft-impl.cpp:99-113);
// here they are evaluated in binary128 and rounded once, so every entry is the correctly rounded double.
//
// High-precision tables follow hp/code.cpp:246-277,378-388: round(cos|sin(2 pi i/n) * 2^64) as 128-bit
// two's-complement fixed point, with 1.0 represented as 2^64-1.  The reference uses NTL RR (150 bits);
// binary128 rounds identically except on astronomically unlikely near-ties.
#include <quadmath.h>
#include <stdint.h>

namespace tfhe_b200 {

int fft_table_entries(int logM) {
    const int M = 1 << logM;
    return M + 7 * (M / 8) + 7 * (M / 64) + (logM == 10 ? 7 * (M / 512) : 0);
}

static void put(double*& p, __float128 num, __float128 den) {   // e^{2 pi i num/den}
    const __float128 ang = 2 * M_PIq * num / den;
    *p++ = (double)cosq(ang);
    *p++ = (double)sinq(ang);
}

void make_fft_tables(int logM, double* out) {
    const int M = 1 << logM, N = 2 * M;
    double* p = out;
    for (int j = 0; j < M; j++) put(p, j, 2 * N);                   // twist e^{i pi j/N}
    int L = M;
    const int npass = (logM == 10) ? 3 : 2;
    for (int pass = 0; pass < npass; pass++) {
        const int Lp = L / 8;
        for (int s = 1; s < 8; s++)
            for (int j = 0; j < Lp; j++) put(p, (__float128)(j * s), L);   // W_L^{j s}
        L = Lp;
    }
}

static void fix64(__float128 x, uint64_t* lo, uint64_t* hi) {
    __float128 s = roundq(ldexpq(x, 64));
    const bool neg = s < 0;
    if (neg) s = -s;
    const __float128 h = floorq(ldexpq(s, -64));
    const __float128 l = s - ldexpq(h, 64);
    unsigned __int128 v = ((unsigned __int128)(uint64_t)h << 64) | (unsigned __int128)(uint64_t)l;
    if (neg) v = (unsigned __int128)0 - v;
    *lo = (uint64_t)v;
    *hi = (uint64_t)(v >> 64);
}
static void hp_cos(int i, int n, uint64_t* lo, uint64_t* hi) {
    i = ((i % n) + n) % n;
    if (i == 0) { *lo = ~0ull; *hi = 0; return; }
    fix64(cosq(2 * M_PIq * i / n), lo, hi);
}
static void hp_sin(int i, int n, uint64_t* lo, uint64_t* hi) {
    i = ((i % n) + n) % n;
    if (i == n / 4) { *lo = ~0ull; *hi = 0; return; }
    fix64(sinq(2 * M_PIq * i / n), lo, hi);
}
void make_hp_tables(int n, int inverse, uint64_t* out) {
    for (int i = 0; i < n; i++) {
        hp_cos(i, n, out + 4 * i, out + 4 * i + 1);
        hp_sin(inverse ? (n - i) % n : i, n, out + 4 * i + 2, out + 4 * i + 3);
    }
}

}  // namespace tfhe_b200
