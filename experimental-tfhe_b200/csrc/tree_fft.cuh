// tree_fft.cuh -- warp-resident negacyclic FP64 transform for sm_100a ("twist-free" product-tree form).
//
// Replaces the reference's AVX2 kernels cb/spqlios/spqlios-ifft-fma.s:63-263 (coefficients -> spectrum) and
// cb/spqlios/spqlios-fft-fma.s:79-274 (spectrum -> coefficients) and their wrappers
// (cb/spqlios/fft_processor_spqlios.cpp:27-170).  Same mathematics as SURVEY A.7 -- fold the N real coefficients
// into M = N/2 complex z_j = c_j + i c_{j+M}; multiplication mod X^N+1 becomes multiplication mod X^M - i --
// but a different algorithm, chosen for the B200's FP64 : shared-memory ratio (64 FMA vs 128 B per clock per SM):
//
//   * Instead of "twist by w^j, then cyclic FFT", the transform walks the product tree of X^M - i:
//       X^L - c = (X^(L/2) - sqrt c)(X^(L/2) + sqrt c),  butterfly (lo, hi) -> (lo + w hi, lo - w hi), w = sqrt c.
//     Every twiddle belongs to a tree NODE, not to a coefficient, so the first four levels use warp-uniform constants
//     and there is no separate twist pass; a forward butterfly is 6 FMAs.
//   * T = M/16 lanes own one polynomial (one warp for N=1024, two for N=2048), 16 points per lane in registers:
//     depths 0-3 in registers, ONE transpose through shared memory, depths 4-7 in registers, and the last
//     log2(M)-8 depths by a half-register exchange with the neighbouring lane (__shfl_xor).
//     For N=1024 that is 128 B/lane of shared-memory traffic each way plus 32 shuffles per transform.
//   * Sibling nodes have twiddles (w, i w): only every other twiddle is stored/loaded.
//
// The spectral order is engine-private (leaf order of the tree, distributed over lanes); bk spectra are produced by
// this same code, so products line up slot by slot.   tools/tree_fft_model.py is the lane-level numpy model.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

typedef double2 cplx;

__device__ __forceinline__ cplx cmulf(cplx a, cplx b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// acc += a*b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

// forward butterfly at a node with twiddle w:  lo' = lo + w hi ; hi' = lo - w hi      (6 FMA)
__device__ __forceinline__ void bf_fwd(cplx& lo, cplx& hi, const cplx w) {
    double xr = fma(w.x, hi.x, lo.x); xr = fma(-w.y, hi.y, xr);
    double xi = fma(w.x, hi.y, lo.y); xi = fma(w.y, hi.x, xi);
    hi.x = fma(2.0, lo.x, -xr); hi.y = fma(2.0, lo.y, -xi);
    lo.x = xr; lo.y = xi;
}
// same with twiddle i*w (the sibling node)
__device__ __forceinline__ void bf_fwd_i(cplx& lo, cplx& hi, const cplx w) {
    double xr = fma(-w.x, hi.y, lo.x); xr = fma(-w.y, hi.x, xr);
    double xi = fma(w.x, hi.x, lo.y); xi = fma(-w.y, hi.y, xi);
    hi.x = fma(2.0, lo.x, -xr); hi.y = fma(2.0, lo.y, -xi);
    lo.x = xr; lo.y = xi;
}
// inverse butterfly:  lo' = lo + hi ; hi' = (lo - hi) conj(w)      (unscaled: the 1/M is folded into the key spectra)
__device__ __forceinline__ void bf_inv(cplx& lo, cplx& hi, const cplx w) {
    const double dr = lo.x - hi.x, di = lo.y - hi.y;
    lo.x += hi.x; lo.y += hi.y;
    hi.x = fma(dr, w.x, di * w.y);
    hi.y = fma(di, w.x, -dr * w.y);
}
// inverse with twiddle i*w: conj(i w) = -i conj(w)
__device__ __forceinline__ void bf_inv_i(cplx& lo, cplx& hi, const cplx w) {
    const double dr = lo.x - hi.x, di = lo.y - hi.y;
    lo.x += hi.x; lo.y += hi.y;
    hi.x = fma(di, w.x, -dr * w.y);
    hi.y = fma(-dr, w.x, -di * w.y);
}

// ---------------------------------------------------------------------------------------------
// Geometry and twiddle-table layout (cplx entries), LOGM = log2(M) in {9, 10}
//   TA  [8]          depths 0-3, essential (even) nodes: (0,0) (1,0) (2,0) (2,2) (3,0) (3,2) (3,4) (3,6)
//   TB  [2][8][16]   depths 4-7 under depth-4 node b: (4,b) (5,2b) (6,4b) (6,4b+2) (7,8b+2k), k<4    index h*128+e*16+b
//                    h = the lane's side in the first exchange; for h = 1 the depth-7 entries are NEGATED
//   TC0 [8][T]       depth 8, per lane                                                             index m*T+t
//   TC1 [8][T]       depth 9 (M=1024 only), per lane                                               index m*T+t
//   TG  [16][T]      unit factors g carried by the slots of a lane (global memory only; standalone transforms)
//   node twiddle w(d,nu) = exp(2 pi i (1 + 4 bitrev_d(nu)) / 2^(d+3))
//
// Select-free lane exchange (numpy model and derivation: tools/tree_fft_model.py).  After depths 4-7 a lane holds ONE
// coefficient of 16 depth-8 nodes; the depth-8 butterfly needs (lo, hi) from two lanes.  The exchange always sends the
// odd registers v[2m+1] and receives into them, on both sides -- no per-register selects.  To make that work, lanes on
// side h = 1 run depth 7 with negated twiddles (their two outputs land swapped), so after the exchange side 0 owns
// (lo, hi) of node 16b+2m and side 1 owns (hi, lo) of node 16b+2m+1; side 1 then uses conj(w) as its depth-8 twiddle and
// obtains conj(w)*plus and -conj(w)*minus: the true outputs times a unit factor g that depends only on (lane, slot).
// Digit spectra and the spectral accumulators of a blind rotation carry the same g, the key spectra are stored TRUE, so the
// multiply-accumulate needs no correction and the backward transform (same tables, mirrored) removes g exactly.
// For M = 1024 the second exchange repeats the construction (depth-8 twiddles negated on its side-1 lanes).
// ---------------------------------------------------------------------------------------------
template <int LOGM> struct TreePlan {
    static constexpr int M = 1 << LOGM;
    static constexpr int N = 2 * M;
    static constexpr int T = M / 16;               // lanes per polynomial
    static constexpr int P = T / 16;               // lanes per depth-4 node
    static constexpr int NS = LOGM - 8;            // shuffle stages
    static constexpr int S = T + T / 16;           // padded row stride of the transpose buffer
    static constexpr int BUF = 16 * S;             // cplx entries per polynomial
    static constexpr int TA = 0;
    static constexpr int TB = 8;
    static constexpr int TC0 = TB + 256;
    static constexpr int TC1 = TC0 + 8 * T;
    static constexpr int TW_TOTAL = TC1 + (NS > 1 ? 8 * T : 0);      // what the kernels keep in shared memory
    static constexpr int TG = TW_TOTAL;
    static constexpr int TABLE_ENTRIES = TW_TOTAL + 16 * T;
};

// Twiddles of tree depths 0-3 (even nodes).  They do not depend on M, so they are compile-time constants read through the
// constant bank instead of shared memory (correctly rounded from 200-bit values; identical to TreePlan table entries TA[0..8)).
__device__ __constant__ double c_tree_ta[16] = {
    0x1.6a09e667f3bcdp-1, 0x1.6a09e667f3bcdp-1,   // w(0,0)
    0x1.d906bcf328d46p-1, 0x1.87de2a6aea963p-2,   // w(1,0)
    0x1.f6297cff75cb0p-1, 0x1.8f8b83c69a60bp-3,   // w(2,0)
    0x1.1c73b39ae68c8p-1, 0x1.a9b66290ea1a3p-1,   // w(2,2)
    0x1.fd88da3d12526p-1, 0x1.917a6bc29b42cp-4,   // w(3,0)
    0x1.44cf325091dd6p-1, 0x1.8bc806b151741p-1,   // w(3,2)
    0x1.c38b2f180bdb1p-1, 0x1.e2b5d3806f63bp-2,   // w(3,4)
    0x1.294062ed59f06p-2, 0x1.e9f4156c62ddap-1,   // w(3,6)
};

// Development instrumentation (tests/dev/br_timeline.py builds a private library with -DBR_TIMELINE): per-warp cycle counts
// between marks, accumulated by lane 0.  Compiles to nothing in the product build.
#ifdef BR_TIMELINE
__device__ long long g_tl_acc[32 * 32];
__device__ __forceinline__ void tl_mark(const int k) {
    __shared__ long long tl_last[16];
    __shared__ long long tl_sum[16][20];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        const long long now = clock64();
        if (k < 0) { for (int i = 0; i < 20; i++) tl_sum[w][i] = 0; }
        else if (k == 99) { if (blockIdx.x == 0) for (int i = 0; i < 20; i++) g_tl_acc[w * 32 + i] = tl_sum[w][i]; }
        else tl_sum[w][k] += now - tl_last[w];
        tl_last[w] = clock64();
    }
}
#define TL(k) tl_mark(k)
#else
#define TL(k) do {} while (0)
#endif

template <int T> __device__ __forceinline__ void lanes_sync(int bar_id) {
    if (T == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Per-lane twiddles from TENSOR MEMORY (blind-rotation kernel, 8-warp configurations).  The 16 (N=1024) / 24 (N=2048) twiddles a
// lane needs after the transpose are lane-specific, so they cannot come from the constant bank; read from shared memory they are
// 16-24 LDS.128 per transform = 14 % of the kernel's LSU wavefronts, and the LSU pipe is its busiest unit (66 %).  Each lane keeps
// them in 64 / 96 of its tensor-memory columns instead: [0,32) depths 4-7 (its TB row), [32,64) depth 8, [64,96) depth 9.
// tcgen05.ld has its own datapath.  ttw = 0 selects the shared-memory path (standalone transforms, 12-warp variants).
// ---------------------------------------------------------------------------------------------
#define TREE_TLD16(r, addr)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
                 : "r"(addr) : "memory")
#define TREE_TST16(r, addr)                                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"       \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr) : "memory")
// eight twiddles (32 columns at taddr) -> registers, in two halves so that the tensor-memory latency can hide behind whatever
// the caller does between issue and collect (a transpose read, a lane exchange)
struct Tw8Regs { uint32_t r0[16], r1[16]; };
__device__ __forceinline__ void tw8_issue(Tw8Regs& q, const uint32_t taddr) {
    TREE_TLD16(q.r0, taddr);
    TREE_TLD16(q.r1, taddr + 16);
}
__device__ __forceinline__ void tw8_collect(cplx (&E)[8], const Tw8Regs& q) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const uint32_t (&r0)[16] = q.r0; const uint32_t (&r1)[16] = q.r1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        E[i] = make_double2(__hiloint2double((int)r0[4 * i + 1], (int)r0[4 * i]), __hiloint2double((int)r0[4 * i + 3], (int)r0[4 * i + 2]));
        E[4 + i] = make_double2(__hiloint2double((int)r1[4 * i + 1], (int)r1[4 * i]), __hiloint2double((int)r1[4 * i + 3], (int)r1[4 * i + 2]));
    }
}
__device__ __forceinline__ void tw8_to_tmem(const cplx (&E)[8], const uint32_t taddr) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            r[4 * i] = (uint32_t)__double2loint(E[4 * h + i].x); r[4 * i + 1] = (uint32_t)__double2hiint(E[4 * h + i].x);
            r[4 * i + 2] = (uint32_t)__double2loint(E[4 * h + i].y); r[4 * i + 3] = (uint32_t)__double2hiint(E[4 * h + i].y);
        }
        TREE_TST16(r, taddr + 16 * h);
    }
}

// four tree depths on the 16 registers of a lane; E = the 8 essential twiddles of these depths
// PRE0 (forward, depth 0 only): the caller has already formed (x - y, x + y) of every `hi` input, w(0,0) = (1 + i) c with c = 1/sqrt 2
template <bool INV, bool PRE0 = false> __device__ __forceinline__ void pass16(cplx (&v)[16], const cplx* __restrict__ E, const int estride) {
    if (!INV) {
        if constexpr (PRE0) {
            const double c = E[0].x;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double s = v[i + 8].x, t = v[i + 8].y;
                v[i + 8].x = fma(-c, s, v[i].x); v[i + 8].y = fma(-c, t, v[i].y);
                v[i].x = fma(c, s, v[i].x);      v[i].y = fma(c, t, v[i].y);
            }
        } else {
            const cplx w = E[0];
#pragma unroll
            for (int i = 0; i < 8; i++) bf_fwd(v[i], v[i + 8], w); }
        {   const cplx w = E[estride];
#pragma unroll
            for (int i = 0; i < 4; i++) { bf_fwd(v[i], v[i + 4], w); bf_fwd_i(v[8 + i], v[12 + i], w); } }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const cplx w = E[(2 + s) * estride];
#pragma unroll
            for (int i = 0; i < 2; i++) { bf_fwd(v[8 * s + i], v[8 * s + i + 2], w); bf_fwd_i(v[8 * s + 4 + i], v[8 * s + 6 + i], w); }
        }
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const cplx w = E[(4 + s) * estride];
            bf_fwd(v[4 * s], v[4 * s + 1], w); bf_fwd_i(v[4 * s + 2], v[4 * s + 3], w);
        }
    } else {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const cplx w = E[(4 + s) * estride];
            bf_inv(v[4 * s], v[4 * s + 1], w); bf_inv_i(v[4 * s + 2], v[4 * s + 3], w);
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const cplx w = E[(2 + s) * estride];
#pragma unroll
            for (int i = 0; i < 2; i++) { bf_inv(v[8 * s + i], v[8 * s + i + 2], w); bf_inv_i(v[8 * s + 4 + i], v[8 * s + 6 + i], w); }
        }
        {   const cplx w = E[estride];
#pragma unroll
            for (int i = 0; i < 4; i++) { bf_inv(v[i], v[i + 4], w); bf_inv_i(v[8 + i], v[12 + i], w); } }
        {   const cplx w = E[0];
#pragma unroll
            for (int i = 0; i < 8; i++) bf_inv(v[i], v[i + 8], w); }
    }
}

// every lane trades its odd registers with the lane at distance `mask` (an involution)
__device__ __forceinline__ void odd_swap(cplx (&v)[16], const int mask) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        v[2 * m + 1].x = __shfl_xor_sync(0xffffffffu, v[2 * m + 1].x, mask);
        v[2 * m + 1].y = __shfl_xor_sync(0xffffffffu, v[2 * m + 1].y, mask);
    }
}
template <int LOGM> __device__ __forceinline__ int tree_side(const int t) {      // side of lane t in the first exchange
    typedef TreePlan<LOGM> P;
    return ((t % P::P) >> (P::NS - 1)) & 1;
}

// Forward: in  v[m] = z_{t + T m}  (t = lane in [0,T), natural coefficient order, stride T)
//          out v[i] = spectrum slot i of this lane (leaf order, private)
// Split in two so the caller can reuse the transpose buffer between the halves (after part A nobody reads buf any more).
struct TreeNoHook { __device__ __forceinline__ void operator()() const {} };
template <int LOGM, bool PRE0 = false, typename Hook = TreeNoHook>
__device__ __forceinline__ void tree_forward_a(cplx (&v)[16], cplx* __restrict__ buf, const cplx* __restrict__ tw, const int t, const int bar_id,
                                               Hook mid = Hook()) {
    typedef TreePlan<LOGM> P;
    constexpr int T = P::T;
    pass16<false, PRE0>(v, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    TL(2);
    lanes_sync<T>(bar_id);                                   // WAR: earlier reads of buf
#pragma unroll
    for (int m = 0; m < 16; m++) buf[m * P::S + t] = v[m];
    mid();                                                   // v is dead here: room to start something long (twiddle loads)
    lanes_sync<T>(bar_id);
    const int b = t / P::P, p = t % P::P;
#pragma unroll
    for (int u = 0; u < 16; u++) v[u] = buf[b * P::S + p + P::P * u];
    lanes_sync<T>(bar_id);                                   // every lane is done with buf
    TL(3);
}
// this lane's post-transpose twiddles -> its tensor-memory columns at ttw (once per kernel)
template <int LOGM, bool TT9 = true>
__device__ __forceinline__ void tree_twiddles_to_tmem(const cplx* __restrict__ tw, const int t, const uint32_t ttw) {
    typedef TreePlan<LOGM> P;
    cplx E[8];
#pragma unroll
    for (int e = 0; e < 8; e++) E[e] = tw[P::TB + tree_side<LOGM>(t) * 128 + e * 16 + t / P::P];
    tw8_to_tmem(E, ttw);
#pragma unroll
    for (int m = 0; m < 8; m++) E[m] = tw[P::TC0 + m * P::T + t];
    tw8_to_tmem(E, ttw + 32);
    if (P::NS > 1 && TT9) {
#pragma unroll
        for (int m = 0; m < 8; m++) E[m] = tw[P::TC1 + m * P::T + t];
        tw8_to_tmem(E, ttw + 64);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
template <int LOGM>
__device__ __forceinline__ void tree_forward_b(cplx (&v)[16], const cplx* __restrict__ tw, const int t) {      // depths 4-7
    typedef TreePlan<LOGM> P;
    pass16<false>(v, tw + P::TB + tree_side<LOGM>(t) * 128 + t / P::P, 16);
}
// same with the twiddles already on their way from tensor memory (tw8_issue(q, ttw) before the transpose read)
__device__ __forceinline__ void tree_forward_b_tm(cplx (&v)[16], const Tw8Regs& q) {
    cplx E[8];
    tw8_collect(E, q);
    pass16<false>(v, E, 1);
}
template <int LOGM, bool TT = false, bool TT9 = TT>
__device__ __forceinline__ void tree_forward_c(cplx (&v)[16], const cplx* __restrict__ tw, const int t, const uint32_t ttw = 0) {      // depths 8..
    typedef TreePlan<LOGM> P;
    constexpr int T = P::T;
    {   // depth 8
        if constexpr (TT) {
            Tw8Regs q; tw8_issue(q, ttw + 32);               // rides behind the lane exchange
            odd_swap(v, P::P >> 1);
            cplx E[8]; tw8_collect(E, q);
#pragma unroll
            for (int m = 0; m < 8; m++) bf_fwd(v[2 * m], v[2 * m + 1], E[m]);
        } else {
            odd_swap(v, P::P >> 1);
            const cplx* e = tw + P::TC0 + t;
#pragma unroll
            for (int m = 0; m < 8; m++) bf_fwd(v[2 * m], v[2 * m + 1], e[m * T]);
        }
    }
    if (P::NS > 1) {   // depth 9 (M = 1024)
        if constexpr (TT9) {
            odd_swap(v, 1);
            Tw8Regs q; tw8_issue(q, ttw + 64);
            cplx E[8]; tw8_collect(E, q);
#pragma unroll
            for (int m = 0; m < 8; m++) bf_fwd(v[2 * m], v[2 * m + 1], E[m]);
        } else {
            odd_swap(v, 1);
            const cplx* e = tw + P::TC1 + t;
#pragma unroll
            for (int m = 0; m < 8; m++) bf_fwd(v[2 * m], v[2 * m + 1], e[m * T]);
        }
    }
}
template <int LOGM>
__device__ __forceinline__ void tree_forward(cplx (&v)[16], cplx* __restrict__ buf, const cplx* __restrict__ tw, const int t, const int bar_id) {
    tree_forward_a<LOGM>(v, buf, tw, t, bar_id);
    tree_forward_b<LOGM>(v, tw, t);
    tree_forward_c<LOGM>(v, tw, t);
}

// Two forward transforms side by side (the two gadget levels cut from one rotated accumulator polynomial): every latency-bound
// step -- the transposes, the twiddle fetches from tensor memory, the lane exchanges -- is followed by FP64 work of the OTHER data
// set, so one warp keeps the FP64 pipe fed where a single transform leaves it idle; the twiddles are fetched once for both.
// One transpose buffer serves both (v goes through first; u's pass A covers v's read-back).
template <int LOGM, bool TT9 = true>
__device__ __forceinline__ void tree_forward2(cplx (&v)[16], cplx (&u)[16], cplx* __restrict__ buf, const cplx* __restrict__ tw, const int t,
                                              const int bar_id, const uint32_t ttw) {
    typedef TreePlan<LOGM> P;
    constexpr int T = P::T;
    const int b = t / P::P, p = t % P::P;
    pass16<false>(v, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    lanes_sync<T>(bar_id);                                   // WAR: earlier reads of buf
#pragma unroll
    for (int m = 0; m < 16; m++) buf[m * P::S + t] = v[m];
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = buf[b * P::S + p + P::P * k];
    pass16<false>(u, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    Tw8Regs q; tw8_issue(q, ttw);                            // depths 4-7, shared by both data sets
    lanes_sync<T>(bar_id);                                   // everybody has read v back
#pragma unroll
    for (int m = 0; m < 16; m++) buf[m * P::S + t] = u[m];
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int k = 0; k < 16; k++) u[k] = buf[b * P::S + p + P::P * k];
    {
        cplx E[8];
        tw8_collect(E, q);
        pass16<false>(v, E, 1);                              // covers u's read-back
        pass16<false>(u, E, 1);
    }
    {   // depth 8
        Tw8Regs q8; tw8_issue(q8, ttw + 32);
        odd_swap(v, P::P >> 1); odd_swap(u, P::P >> 1);
        cplx E[8]; tw8_collect(E, q8);
#pragma unroll
        for (int m = 0; m < 8; m++) { bf_fwd(v[2 * m], v[2 * m + 1], E[m]); bf_fwd(u[2 * m], u[2 * m + 1], E[m]); }
    }
    if (P::NS > 1) {   // depth 9 (M = 1024)
        cplx E[8];
        if constexpr (TT9) {
            Tw8Regs q9; tw8_issue(q9, ttw + 64);
            odd_swap(v, 1); odd_swap(u, 1);
            tw8_collect(E, q9);
        } else {
            odd_swap(v, 1); odd_swap(u, 1);
#pragma unroll
            for (int m = 0; m < 8; m++) E[m] = tw[P::TC1 + m * T + t];
        }
#pragma unroll
        for (int m = 0; m < 8; m++) { bf_fwd(v[2 * m], v[2 * m + 1], E[m]); bf_fwd(u[2 * m], u[2 * m + 1], E[m]); }
    }
}

// Backward: the exact mirror.  in v[i] = spectrum slot i (times g) ; out v[m] = M * z_{t + T m}
template <int LOGM, bool TT = false>
__device__ __forceinline__ void tree_backward(cplx (&v)[16], cplx* __restrict__ buf, const cplx* __restrict__ tw, const int t, const int bar_id,
                                              const uint32_t ttw = 0) {
    typedef TreePlan<LOGM> P;
    constexpr int T = P::T;
    const int b = t / P::P, p = t % P::P;
    Tw8Regs qn;                                             // the next stage's twiddles, loaded one stage ahead
    if (P::NS > 1) {
        if constexpr (TT) {
            Tw8Regs q; tw8_issue(q, ttw + 64); tw8_issue(qn, ttw + 32);
            cplx E[8]; tw8_collect(E, q);
#pragma unroll
            for (int m = 0; m < 8; m++) bf_inv(v[2 * m], v[2 * m + 1], E[m]);
        } else {
            const cplx* e = tw + P::TC1 + t;
#pragma unroll
            for (int m = 0; m < 8; m++) bf_inv(v[2 * m], v[2 * m + 1], e[m * T]);
        }
        odd_swap(v, 1);
    }
    {
        if constexpr (TT) {
            if (P::NS == 1) tw8_issue(qn, ttw + 32);
            cplx E[8]; tw8_collect(E, qn);
            tw8_issue(qn, ttw);                              // depths 7-4, collected after the exchange
#pragma unroll
            for (int m = 0; m < 8; m++) bf_inv(v[2 * m], v[2 * m + 1], E[m]);
        } else {
            const cplx* e = tw + P::TC0 + t;
#pragma unroll
            for (int m = 0; m < 8; m++) bf_inv(v[2 * m], v[2 * m + 1], e[m * T]);
        }
        odd_swap(v, P::P >> 1);
    }
    TL(10);
    if constexpr (TT) { cplx E[8]; tw8_collect(E, qn); pass16<true>(v, E, 1); }
    else pass16<true>(v, tw + P::TB + tree_side<LOGM>(t) * 128 + b, 16);
    TL(11);
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int u = 0; u < 16; u++) buf[b * P::S + p + P::P * u] = v[u];
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int m = 0; m < 16; m++) v[m] = buf[m * P::S + t];
    TL(12);
    pass16<true>(v, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    TL(13);
}

// Two backward transforms side by side (the two polynomials of a TLWE accumulator): the stages are interleaved so each
// latency-bound step (twiddle fetch, lane exchange, transpose) is paid once for two independent data sets, and the twiddles are
// fetched once.  Uses the one transpose buffer twice.  Needs 128 data registers: only where no key values are in flight.
template <int LOGM, bool TT = false, bool TT9 = TT>
__device__ __forceinline__ void tree_backward2(cplx (&v)[16], cplx (&u)[16], cplx* __restrict__ buf, const cplx* __restrict__ tw, const int t,
                                               const int bar_id, const uint32_t ttw = 0) {
    typedef TreePlan<LOGM> P;
    constexpr int T = P::T;
    const int b = t / P::P, p = t % P::P;
    Tw8Regs qn;
    if (P::NS > 1) {
        cplx E[8];
        if constexpr (TT9) { Tw8Regs q; tw8_issue(q, ttw + 64); tw8_issue(qn, ttw + 32); tw8_collect(E, q); }
        else {
            if constexpr (TT) tw8_issue(qn, ttw + 32);
#pragma unroll
            for (int m = 0; m < 8; m++) E[m] = tw[P::TC1 + m * T + t];
        }
#pragma unroll
        for (int m = 0; m < 8; m++) { bf_inv(v[2 * m], v[2 * m + 1], E[m]); bf_inv(u[2 * m], u[2 * m + 1], E[m]); }
        odd_swap(v, 1); odd_swap(u, 1);
    }
    {
        cplx E[8];
        if constexpr (TT) { if (P::NS == 1) tw8_issue(qn, ttw + 32); tw8_collect(E, qn); tw8_issue(qn, ttw); }
        else {
#pragma unroll
            for (int m = 0; m < 8; m++) E[m] = tw[P::TC0 + m * T + t];
        }
#pragma unroll
        for (int m = 0; m < 8; m++) { bf_inv(v[2 * m], v[2 * m + 1], E[m]); bf_inv(u[2 * m], u[2 * m + 1], E[m]); }
        odd_swap(v, P::P >> 1); odd_swap(u, P::P >> 1);
    }
    TL(10);
    {
        cplx E[8];
        if constexpr (TT) tw8_collect(E, qn);
        else {
#pragma unroll
            for (int e = 0; e < 8; e++) E[e] = tw[P::TB + tree_side<LOGM>(t) * 128 + e * 16 + b];
        }
        pass16<true>(v, E, 1);
        pass16<true>(u, E, 1);
    }
    TL(11);
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int k = 0; k < 16; k++) buf[b * P::S + p + P::P * k] = v[k];
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int m = 0; m < 16; m++) v[m] = buf[m * P::S + t];
    lanes_sync<T>(bar_id);                                   // everybody has read the first polynomial
#pragma unroll
    for (int k = 0; k < 16; k++) buf[b * P::S + p + P::P * k] = u[k];
    lanes_sync<T>(bar_id);
#pragma unroll
    for (int m = 0; m < 16; m++) u[m] = buf[m * P::S + t];
    TL(12);
    pass16<true>(v, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    pass16<true>(u, reinterpret_cast<const cplx*>(c_tree_ta), 1);
    TL(13);
}

// ---------------------------------------------------------------------------------------------
// double -> torus, truncation toward zero then wrap, on the FP64 pipe (F2I is full rate there; the
// integer bit-twiddling form costs ~25 ALU ops, see profiles/microbench_r1.txt).
//   Torus32: int32_t(int64_t(x))                     cb/spqlios/fft_processor_spqlios.cpp:102
//   Torus64: significand shifted by the exponent     cb/spqlios/fft_processor_spqlios.cpp:131-142
//            == trunc(x) mod 2^64.  x - 2^64 rint(x 2^-64) is exact and lies in [-2^63, 2^63]; |x| >= 2^53 is already an
//            integer, so truncating the remainder equals truncating x.  +2^63 wraps to INT64_MIN like the reference.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t double_to_torus32(double x) { return (int32_t)__double2ll_rz(x); }
__device__ __forceinline__ int64_t double_to_torus64(double x) {
    const double q = rint(x * 5.42101086242752217e-20);              // 2^-64
    const double r = fma(q, -18446744073709551616.0, x);
    return r >= 9223372036854775808.0 ? (int64_t)0x8000000000000000ull : __double2ll_rz(r);
}

// (X^a - 1) * P at coefficient j, a in [0, 2N)   (cb/numeric_functions.cpp:304-323)
template <typename T, int N> __device__ __forceinline__ T rot_minus_one(const T* __restrict__ P, int j, int a) {
    const int idx = (j - a) & (2 * N - 1);
    const T r = P[idx & (N - 1)];
    return (T)(((idx & N) ? (T)0 - r : r) - P[j]);
}

}  // namespace tfhe_b200
