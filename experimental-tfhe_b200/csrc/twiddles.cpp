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
#include <vector>

namespace tfhe_b200 {

int fft_table_entries(int logM) {
    const int M = 1 << logM, T = M / 16;
    return 8 + 256 + 8 * T + (logM == 10 ? 8 * T : 0) + 16 * T;     // TA | TB | TC0 | TC1 | TG  (TreePlan)
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

typedef __complex128 q128;
static q128 nodeq(int d, unsigned nu) {
    const __float128 ang = 2 * M_PIq * (__float128)(1 + 4 * (long)bitrev(nu, d)) / (__float128)(1L << (d + 3));
    return cosq(ang) + sinq(ang) * 1.0Qi;
}
static void put(double*& p, q128 z) { *p++ = (double)crealq(z); *p++ = (double)cimagq(z); }

// Layout and the select-free exchange construction: TreePlan in tree_fft.cuh; model: tools/tree_fft_model.py (tables2).
void make_fft_tables(int logM, double* out) {
    const int M = 1 << logM, T = M / 16, P = T / 16, NS = logM - 8;
    double* p = out;
    // TA: depths 0-3, even nodes
    node(p, 0, 0); node(p, 1, 0); node(p, 2, 0); node(p, 2, 2);
    for (int s = 0; s < 4; s++) node(p, 3, 2 * s);
    // TB[h][e][b]: depths 4-7 below depth-4 node b; side-1 lanes take depth 7 negated
    for (int h = 0; h < 2; h++)
        for (int e = 0; e < 8; e++)
            for (int b = 0; b < 16; b++) {
                if (e == 0) node(p, 4, b);
                else if (e == 1) node(p, 5, 2 * b);
                else if (e < 4) node(p, 6, 4 * b + 2 * (e - 2));
                else put(p, (h ? -1 : 1) * nodeq(7, 8 * b + 2 * (e - 4)));
            }
    // TC0[m][t], TC1[m][t], TG[i][t]
    std::vector<q128> c8((size_t)8 * T), c9((size_t)8 * T), g((size_t)16 * T);
    for (int t = 0; t < T; t++) {
        const int b = t / P, pp = t % P;
        const int h8 = (pp >> (NS - 1)) & 1, h9 = NS > 1 ? (pp & 1) : 0;
        for (int m = 0; m < 8; m++) {
            const unsigned A = 16 * b + 2 * m + h8;                 // the depth-8 node this lane completes
            const q128 w8 = nodeq(8, A);
            q128 c = h8 ? conjq(w8) : w8;
            if (NS > 1 && h9) c = -c;
            c8[(size_t)m * T + t] = c;
            const q128 g8[2] = {h8 ? conjq(w8) : (q128)1, h8 ? -conjq(w8) : (q128)1};     // on (plus, minus) of node A
            if (NS == 1) {
                g[(size_t)(2 * m) * T + t] = g8[0]; g[(size_t)(2 * m + 1) * T + t] = g8[1];
            } else {
                const unsigned D = 2 * A + h9;                      // depth-9 node: plus (h9 = 0) or minus (h9 = 1) child of A
                const q128 w9 = nodeq(9, D);
                c9[(size_t)m * T + t] = h9 ? conjq(w9) : w9;
                g[(size_t)(2 * m) * T + t] = g8[h9] * (h9 ? conjq(w9) : (q128)1);
                g[(size_t)(2 * m + 1) * T + t] = g8[h9] * (h9 ? -conjq(w9) : (q128)1);
            }
        }
    }
    for (size_t i = 0; i < c8.size(); i++) put(p, c8[i]);
    if (NS > 1) for (size_t i = 0; i < c9.size(); i++) put(p, c9[i]);
    for (size_t i = 0; i < g.size(); i++) put(p, g[i]);
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
