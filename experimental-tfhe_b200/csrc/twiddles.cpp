// twiddles.cpp -- host-side twiddle tables for the device transforms.
//
// FP64 tables follow TreePlan<LOGM> (tree_fft.cuh): every entry is the twiddle of a NODE of the product tree of
// X^M - i,  w(d,nu) = exp(2 pi i (1 + 4 bitrev_d(nu)) / 2^(d+3)),  stored for even nodes only where the sibling is i*w.
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
    const int M = 1 << logM, T = M / 16;
    return 8 + 128 + 4 * 32 + (logM == 10 ? 8 * T : 0);
}

static unsigned bitrev(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static void node(double*& p, int d, unsigned nu) {   // w(d,nu)
    const __float128 ang = 2 * M_PIq * (__float128)(1 + 4 * (long)bitrev(nu, d)) / (__float128)(1L << (d + 3));
    *p++ = (double)cosq(ang);
    *p++ = (double)sinq(ang);
}

void make_fft_tables(int logM, double* out) {
    const int M = 1 << logM, T = M / 16, NS = logM - 8;
    double* p = out;
    // TA: depths 0-3, even nodes
    node(p, 0, 0); node(p, 1, 0); node(p, 2, 0); node(p, 2, 2);
    for (int s = 0; s < 4; s++) node(p, 3, 2 * s);
    // TB[e][b]: depths 4-7 below depth-4 node b
    for (int e = 0; e < 8; e++)
        for (int b = 0; b < 16; b++) {
            if (e == 0) node(p, 4, b);
            else if (e == 1) node(p, 5, 2 * b);
            else if (e < 4) node(p, 6, 4 * b + 2 * (e - 2));
            else node(p, 7, 8 * b + 2 * (e - 4));
        }
    // TC0[k][g]: depth 8, nodes 8g + 2k
    for (int k = 0; k < 4; k++)
        for (int g = 0; g < 32; g++) node(p, 8, 8 * g + 2 * k);
    // TC1[k][t]: depth 9 (M = 1024), nodes 16(t>>1) + 2k + (t&1)
    if (NS > 1)
        for (int k = 0; k < 8; k++)
            for (int t = 0; t < T; t++) node(p, 9, 16 * (t >> 1) + 2 * k + (t & 1));
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
