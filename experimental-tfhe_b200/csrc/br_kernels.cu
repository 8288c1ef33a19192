// br_kernels.cu -- fused blind rotation for sm_100a.
//
// One lane GROUP (T = N/32 lanes: one warp for N=1024, two warps for N=2048) owns one TLWE accumulator for all n
// CMUX steps; the accumulator lives in shared memory for the whole kernel and is touched in HBM only at the start
// (input LWE sample) and at the end (extracted LWE sample).  Per step the group fuses
//   (X^a - 1) * ACC            tLweMulByXaiMinusOne         cb/tlwe_functions.cpp:209-213
//   gadget decomposition       tGswTorus32PolynomialDecompH cb/tgsw_functions.cpp:224-337
//                              tGswTorus64PolynomialDecompH cb/poc_CircuitBootstrapping.cpp:492-515
//   2l forward transforms      IntPolynomial_ifft           cb/tgsw_functions.cpp:438
//   2l x 2 spectral MACs       tLweFFTAddMulRTo             cb/tlwe_functions.cpp:318-325
//   2 backward transforms      tLweFromFFTConvert           cb/tlwe_functions.cpp:299-305
//   ACC += result              tLweAddTo                    cb/tlwe_functions.cpp:163-170
// i.e. tfhe_MuxRotate_FFT (cb/lwe_functions.cpp:328-333) inside the loop of tfhe_blindRotate_FFT (:337-361), plus
// modulus switch, test-vector rotation and sample extraction on either side (tfhe_bootstrap_woKS_FFT :399-430;
// circuitBootstrapWoKS cb/poc_CircuitBootstrapping.cpp:530-659).
//
// The spectra of the 2l digit polynomials never leave registers: each lane multiply-accumulates its 16 spectrum
// slots against the bootstrapping-key spectra (read with coalesced 16-byte loads, L2 resident) into two register
// accumulators, which the two backward transforms then consume.  N=1024 groups are single warps, so the only
// synchronisation inside a CMUX is __syncwarp around the transpose.
// The reference's `if (barai==0) continue` (:350) is kept (group-uniform branch).
#include "engine.h"
#include "tree_fft.cuh"
#include "bk_pipe.cuh"
#include <cstdlib>

#ifndef BR_SHARE_TW64
#define BR_SHARE_TW64 0       // 1: Torus64 with stash keeps the per-lane twiddles ONCE per lane quarter, depth 9 included (measured slower: 196 vs 187 ms)
#endif
#ifndef BR_PRE0
#define BR_PRE0 1
#endif
#ifndef BR_FASTDIGIT
#define BR_FASTDIGIT 1
#endif

namespace tfhe_b200 {

template <typename Torus> struct TorusTraits;
template <> struct TorusTraits<int32_t> { typedef uint32_t U; static constexpr int W = 32; };
template <> struct TorusTraits<int64_t> { typedef uint64_t U; static constexpr int W = 64; };

// decomposition offset: Torus32 library form has no rounding bit (cb/tgsw_functions.cpp:30-36);
// the Torus64 PoC form carries one (cb/poc_CircuitBootstrapping.cpp:349-350).  SURVEY A.5.
__device__ __forceinline__ uint32_t decomp_offset(uint32_t, int l, int Bgbit) {
    uint32_t t = 0;
    for (int i = 0; i < l; i++) t += 1u << (32 - (i + 1) * Bgbit);
    return t * (uint32_t)((1 << Bgbit) / 2);
}
__device__ __forceinline__ uint64_t decomp_offset(uint64_t, int l, int Bgbit) {
    uint64_t t = 0;
    for (int i = 0; i <= l; i++) t |= 1ull << (63 - i * Bgbit);
    return t;
}
__device__ __forceinline__ int32_t to_torus(double x, int32_t) { return double_to_torus32(x); }
__device__ __forceinline__ int64_t to_torus(double x, int64_t) { return double_to_torus64(x); }

// modSwitchFromTorus32 (cb/numeric_functions.cpp:54-60) for Msize = 2^log2M
__device__ __forceinline__ int modswitch32(int32_t x, int log2Msize) {
    const uint64_t half = 1ull << (63 - log2Msize);           // interv/2, interv = 2^(64-log2Msize)
    const uint64_t phase64 = ((uint64_t)(uint32_t)x << 32) + half;
    return (int)(phase64 >> (64 - log2Msize));
}

// ---------------------------------------------------------------------------------------------
// Spectral accumulators in TENSOR MEMORY.  The two accumulators of a CMUX (2 x 16 complex doubles per lane = 128 32-bit
// registers) would pin half the register file and cap the SM at 8 warps.  They live in TMEM instead: lane i of warp w
// owns TMEM lane 32*(w%4)+i, columns [(w/4)*128, +128) -- R0 in the first 64 columns, R1 in the next 64 -- and are
// read-modify-written 16 columns at a time around each multiply-accumulate (tcgen05.ld/st.32x32b, SASS LDTM/STTM;
// no MMA is involved).  profiles/tmem_probe_r1.txt: an accumulator round trip sustains ~445 B/clk/SM next to LDS and DFMA.
// ---------------------------------------------------------------------------------------------
#define TFHE_TLD16(r, addr)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
                 : "r"(addr) : "memory")
#define TFHE_TST16(r, addr)                                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"       \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(addr) : "memory")
#define TFHE_TLD8(r, addr)                                                                                                \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                   \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr) : "memory")
#define TFHE_TST8(r, addr)                                                                                                \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"                                   \
                 ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(addr) : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// R (in TMEM at taddr, 64 columns) (+)= v (.) b over this lane's 16 spectrum slots; FIRST: plain product, nothing to load.
// b(i) yields the key value of slot i (shared memory or registers).
// Ordering rule used throughout: a tcgen05.ld of columns this thread has stored needs tcgen05.wait::st in between.  The wait sits
// in front of the LOADS (here, load_tmem, the stash reload), not behind the stores, so a store's latency overlaps with whatever
// comes next (the following decomposition and transform) instead of being waited out on the spot.
// PIPE: the next chunk's load is in flight during this chunk's FMAs.  Pays for N = 2048 (+1.4 %), costs 4 % for N = 1024, where the
// compiler overlaps the depth-8 butterflies with the plain loop (profiles/r1_notes.md) -- so it is a template choice.
template <bool FIRST, bool PIPE = false, typename BFn>
__device__ __forceinline__ void mac_tmem(const uint32_t taddr, const cplx (&v)[16], BFn b) {
    if (!FIRST) tmem_wait_st();
    uint32_t rr[PIPE ? 2 : 1][16];
    if (PIPE && !FIRST) TFHE_TLD16(rr[0], taddr);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t (&r)[16] = rr[PIPE ? (c & 1) : 0];
        if (!FIRST) {
            if (PIPE) { tmem_wait_ld(); if (c < 3) TFHE_TLD16(rr[PIPE ? ((c + 1) & 1) : 0], taddr + 16 * (c + 1)); }
            else { TFHE_TLD16(r, taddr + 16 * c); tmem_wait_ld(); }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cplx R = FIRST ? make_double2(0.0, 0.0)
                           : make_double2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
            cfma(R, v[4 * c + i], b(4 * c + i));
            r[4 * i] = (uint32_t)__double2loint(R.x); r[4 * i + 1] = (uint32_t)__double2hiint(R.x);
            r[4 * i + 2] = (uint32_t)__double2loint(R.y); r[4 * i + 3] = (uint32_t)__double2hiint(R.y);
        }
        TFHE_TST16(r, taddr + 16 * c);
    }
}
// same, key values from tensor memory (KeyPipe): kaddr = the 64 key columns of this polynomial
template <bool FIRST>
__device__ __forceinline__ void mac_tmem_keytm(const uint32_t taddr, const uint32_t kaddr, const cplx (&v)[16]) {
    if (!FIRST) tmem_wait_st();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t r[16], kq[16];
        TFHE_TLD16(kq, kaddr + 16 * c);
        if (!FIRST) { TFHE_TLD16(r, taddr + 16 * c); }
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cplx R = FIRST ? make_double2(0.0, 0.0)
                           : make_double2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
            const cplx b = make_double2(__hiloint2double((int)kq[4 * i + 1], (int)kq[4 * i]), __hiloint2double((int)kq[4 * i + 3], (int)kq[4 * i + 2]));
            cfma(R, v[4 * c + i], b);
            r[4 * i] = (uint32_t)__double2loint(R.x); r[4 * i + 1] = (uint32_t)__double2hiint(R.x);
            r[4 * i + 2] = (uint32_t)__double2loint(R.y); r[4 * i + 3] = (uint32_t)__double2hiint(R.y);
        }
        TFHE_TST16(r, taddr + 16 * c);
    }
}
// both accumulators (128 columns) with eight loads in flight
__device__ __forceinline__ void load_tmem2(cplx (&R0)[16], cplx (&R1)[16], const uint32_t taddr) {
    uint32_t r[8][16];
    tmem_wait_st();
#pragma unroll
    for (int c = 0; c < 8; c++) TFHE_TLD16(r[c], taddr + 16 * c);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            R0[4 * c + i] = make_double2(__hiloint2double((int)r[c][4 * i + 1], (int)r[c][4 * i]), __hiloint2double((int)r[c][4 * i + 3], (int)r[c][4 * i + 2]));
            R1[4 * c + i] = make_double2(__hiloint2double((int)r[4 + c][4 * i + 1], (int)r[4 + c][4 * i]), __hiloint2double((int)r[4 + c][4 * i + 3], (int)r[4 + c][4 * i + 2]));
        }
    }
}
__device__ __forceinline__ void load_tmem(cplx (&R)[16], const uint32_t taddr) {
    uint32_t r[4][16];
    tmem_wait_st();
#pragma unroll
    for (int c = 0; c < 4; c++) TFHE_TLD16(r[c], taddr + 16 * c);          // all four in flight, one wait
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            R[4 * c + i] = make_double2(__hiloint2double((int)r[c][4 * i + 1], (int)r[c][4 * i]), __hiloint2double((int)r[c][4 * i + 3], (int)r[c][4 * i + 2]));
    }
}

// forward transform of one digit polynomial + its two multiply-accumulates.
//   KM_REGS2: both key polynomials are prefetched into registers with ld.global.nc right after depths 4-7 (8 warps per SM)
//   KM_REGS1: one key polynomial's worth of registers: BK[p][0] is prefetched the same way, and as the first multiply-accumulate
//             consumes it, chunk by chunk, the freed registers take BK[p][1] (12 warps per SM at 168 registers)
//   KM_TMEM : they wait in tensor memory (KeyPipe, bk_pipe.cuh), no key registers at all (12 warps per SM)
enum { KM_REGS2 = 0, KM_TMEM = 1, KM_REGS1 = 2 };
template <int LOGM, bool FIRST, int KM, bool NOTT9 = false, bool PRE0 = false>
__device__ __forceinline__ void forward_and_mac(cplx (&v)[16], const uint32_t tacc, const cplx* __restrict__ bkp,
                                                cplx* __restrict__ buf, KeyPipe& kp,
                                                const cplx* __restrict__ tw, const int t, const int bar_id, const uint32_t ttw) {
    typedef TreePlan<LOGM> P;
    if constexpr (KM == KM_REGS2) {
        Tw8Regs q;
        tree_forward_a<LOGM, PRE0>(v, buf, tw, t, bar_id, [&]() { tw8_issue(q, ttw); });      // depths 4-7 twiddles ride behind the transpose
        tree_forward_b_tm(v, q);
    } else {
        tree_forward_a<LOGM>(v, buf, tw, t, bar_id);
        if (KM == KM_TMEM && (t & 31) == 0) kp.poll();
        tree_forward_b<LOGM>(v, tw, t);
    }
    TL(4);
    if (KM == KM_REGS1) {
        const cplx* __restrict__ g0 = bkp + t;
        asm volatile("" : "+l"(g0) : "d"(v[0].x), "d"(v[15].y));          // do not hoist the loads above the pass
        cplx kb[16];
#pragma unroll
        for (int i = 0; i < 16; i++) kb[i] = __ldg(g0 + i * P::T);
        tree_forward_c<LOGM, KM == KM_REGS2, KM == KM_REGS2 && !(NOTT9)>(v, tw, t, ttw);
        TL(6);
        // R0 += v * BK[p][0]; each consumed chunk's registers are refilled with the same slots of BK[p][1]
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t r[16];
            if (!FIRST) { TFHE_TLD16(r, tacc + 16 * c); tmem_wait_ld(); }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                cplx R = FIRST ? make_double2(0.0, 0.0)
                               : make_double2(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
                cfma(R, v[4 * c + i], kb[4 * c + i]);
                r[4 * i] = (uint32_t)__double2loint(R.x); r[4 * i + 1] = (uint32_t)__double2hiint(R.x);
                r[4 * i + 2] = (uint32_t)__double2loint(R.y); r[4 * i + 3] = (uint32_t)__double2hiint(R.y);
            }
            TFHE_TST16(r, tacc + 16 * c);
#pragma unroll
            for (int i = 0; i < 4; i++) kb[4 * c + i] = __ldg(g0 + P::M + (4 * c + i) * P::T);
        }
        TL(7);
        mac_tmem<FIRST>(tacc + 64, v, [&](int i) { return kb[i]; });
        TL(8);
    } else if (KM == KM_REGS2) {
        // its L2 round trip hides behind the exchange stage
        const cplx* __restrict__ g1 = bkp + P::M + t;
        asm volatile("" : "+l"(g1) : "d"(v[0].x), "d"(v[15].y));          // do not hoist the loads above the pass
        cplx b1[16], b0r[16];
#pragma unroll
        for (int i = 0; i < 16; i++) b1[i] = __ldg(g1 + i * P::T);
#pragma unroll
        for (int i = 0; i < 16; i++) b0r[i] = __ldg(g1 - P::M + i * P::T);
        TL(5);
        tree_forward_c<LOGM, KM == KM_REGS2, KM == KM_REGS2 && !(NOTT9)>(v, tw, t, ttw);
        TL(6);
        // (loading both accumulators' chunks together, or the next chunk during the FMAs, was 5 % slower each time: the compiler
        //  overlaps the depth-8 butterflies with this loop as it stands -- profiles/r1_notes.md)
        mac_tmem<FIRST, LOGM == 10>(tacc, v, [&](int i) { return b0r[i]; });
        TL(7);
        mac_tmem<FIRST, LOGM == 10>(tacc + 64, v, [&](int i) { return b1[i]; });
        TL(8);
    } else {
        tree_forward_c<LOGM, KM == KM_REGS2, KM == KM_REGS2 && !(NOTT9)>(v, tw, t, ttw);
        TL(6);
        kp.acquire(t & 31);
        TL(16);
        mac_tmem_keytm<FIRST>(tacc, kp.tkey, v);
        TL(7);
        mac_tmem_keytm<FIRST>(tacc + 64, kp.tkey + 64, v);
        TL(8);
        kp.release(t & 31);              // the key loads of both polynomials have been waited for
        TL(17);
    }
}

template <typename Torus> struct StashWords { static constexpr int PER_C = 8 * (int)(sizeof(Torus) / 4); };   // words per c (4 complex)
// ---------------------------------------------------------------------------------------------
// Two digit polynomials at once (gadget levels lev, lev+1 of one accumulator polynomial): forward transforms side by side
// (tree_forward2) and ONE pass over the spectral accumulators for both,
//     R0 += v (.) BK[p][0] + u (.) BK[p+1][0],   R1 += v (.) BK[p][1] + u (.) BK[p+1][1]
// so each tensor-memory round trip carries 16 FMAs per slot instead of 8 and happens half as often.  The four key values of a
// slot ride in a rolling register window KW slots deep: the first KW slots are requested before depths 4-7 (their L2 round trip
// hides behind two transforms' worth of butterflies), and every slot that has been consumed is refilled with the slot KW ahead.
// 4 * KW * 4 registers instead of the 128 a whole key pair would pin, which is what makes room for the second data set.
// ---------------------------------------------------------------------------------------------
template <int KW> struct KeyWindow { cplx k[KW][4]; };
template <int LOGM, int KW>
__device__ __forceinline__ void keywin_load(KeyWindow<KW>& w, const cplx* __restrict__ kb, const int slot) {
    typedef TreePlan<LOGM> P;
#pragma unroll
    for (int j = 0; j < 4; j++) w.k[slot % KW][j] = __ldg(kb + (size_t)j * P::M + slot * P::T);
}
template <int LOGM, int KW>
__device__ __forceinline__ void mac2_tmem(const bool first, const uint32_t tacc, const cplx (&v)[16], const cplx (&u)[16],
                                          KeyWindow<KW>& w, const cplx* __restrict__ kb) {
    if (!first) tmem_wait_st();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t r0[16], r1[16];
        if (!first) { TFHE_TLD16(r0, tacc + 16 * c); TFHE_TLD16(r1, tacc + 64 + 16 * c); tmem_wait_ld(); }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int s = 4 * c + i;
            cplx R0 = first ? make_double2(0.0, 0.0)
                            : make_double2(__hiloint2double((int)r0[4 * i + 1], (int)r0[4 * i]), __hiloint2double((int)r0[4 * i + 3], (int)r0[4 * i + 2]));
            cplx R1 = first ? make_double2(0.0, 0.0)
                            : make_double2(__hiloint2double((int)r1[4 * i + 1], (int)r1[4 * i]), __hiloint2double((int)r1[4 * i + 3], (int)r1[4 * i + 2]));
            cfma(R0, v[s], w.k[s % KW][0]); cfma(R1, v[s], w.k[s % KW][1]);
            cfma(R0, u[s], w.k[s % KW][2]); cfma(R1, u[s], w.k[s % KW][3]);
            r0[4 * i] = (uint32_t)__double2loint(R0.x); r0[4 * i + 1] = (uint32_t)__double2hiint(R0.x);
            r0[4 * i + 2] = (uint32_t)__double2loint(R0.y); r0[4 * i + 3] = (uint32_t)__double2hiint(R0.y);
            r1[4 * i] = (uint32_t)__double2loint(R1.x); r1[4 * i + 1] = (uint32_t)__double2hiint(R1.x);
            r1[4 * i + 2] = (uint32_t)__double2loint(R1.y); r1[4 * i + 3] = (uint32_t)__double2hiint(R1.y);
            if (s + KW < 16) {
                const cplx* __restrict__ kn = kb;
                asm volatile("" : "+l"(kn) : "d"(R1.y));          // the refill is requested here, not hoisted to the top
                keywin_load<LOGM, KW>(w, kn, s + KW);
            }
        }
        TFHE_TST16(r0, tacc + 16 * c);
        TFHE_TST16(r1, tacc + 64 + 16 * c);
    }
}

// torus value -> the digit of gadget level with shift sh, as a double
template <typename U> __device__ __forceinline__ double digit_of(const U x, const int sh, const uint32_t mask, const int half) {
    return (double)((int)((uint32_t)(x >> sh) & mask) - half);
}

// One CMUX, digit polynomials taken two at a time (see above).  Same contract as cmux_step.  l even: l/2 pairs per accumulator
// polynomial; l odd: the last level goes through the single-polynomial path.  The rotated differences u = (X^a - 1) ACC_q + offset
// are cut into both digits of the first pair as they are formed; later pairs (l > 2) reload them from the tensor-memory stash.
template <int LOGM, typename Torus, int KW, bool PLAIN = false>
__device__ __forceinline__ void cmux_step2(Torus* __restrict__ acc, const int a, const cplx* __restrict__ bk,
                                           const int l, const int Bgbit, cplx* __restrict__ buf, const uint32_t tacc,
                                           const cplx* __restrict__ tw, const int t, const int bar_id, const uint32_t ttw) {
    typedef TreePlan<LOGM> P;
    typedef typename TorusTraits<Torus>::U U;
    constexpr int M = P::M, N = P::N, T = P::T, W = TorusTraits<Torus>::W;
    constexpr int WPC = StashWords<Torus>::PER_C;
    constexpr bool NOTT9 = sizeof(Torus) == 8;                    // Torus64: depth-9 twiddles in shared memory (stash takes their columns)
    const U offset = decomp_offset((U)0, l, Bgbit);
    const uint32_t mask = (1u << Bgbit) - 1u;
    const int half = 1 << (Bgbit - 1);
    const bool stash = l > 2;
    KeyPipe kp_unused{};

#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        const Torus* __restrict__ aq = acc + q * N;
#pragma unroll 1
        for (int lev = 0; lev < l; lev += 2) {
            const int p = q * l + lev;
            const cplx* __restrict__ kb = bk + (size_t)(p * 2) * M + t;
            if (lev + 1 < l) {
                const int sh0 = W - (lev + 1) * Bgbit, sh1 = sh0 - Bgbit;
                cplx v[16], u[16];
                if (lev == 0) {
                    int a2 = a, tj = t;
                    asm volatile("" : "+r"(a2), "+r"(tj));          // keep the 32 rotated addresses from being hoisted out of the loop
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t w[WPC];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int j = tj + T * (4 * c + i);
                            const U ure = (U)(PLAIN ? aq[j] : rot_minus_one<Torus, N>(aq, j, a2)) + offset;
                            const U uim = (U)(PLAIN ? aq[j + M] : rot_minus_one<Torus, N>(aq, j + M, a2)) + offset;
                            v[4 * c + i] = make_double2(digit_of(ure, sh0, mask, half), digit_of(uim, sh0, mask, half));
                            u[4 * c + i] = make_double2(digit_of(ure, sh1, mask, half), digit_of(uim, sh1, mask, half));
                            if (WPC == 8) { w[(2 * i) % WPC] = (uint32_t)ure; w[(2 * i + 1) % WPC] = (uint32_t)uim; }
                            else { w[(4 * i) % WPC] = (uint32_t)ure; w[(4 * i + 1) % WPC] = (uint32_t)((uint64_t)ure >> 32);
                                   w[(4 * i + 2) % WPC] = (uint32_t)uim; w[(4 * i + 3) % WPC] = (uint32_t)((uint64_t)uim >> 32); }
                        }
                        if (stash) {
                            if constexpr (WPC == 8) { TFHE_TST8(w, tacc + 128 + WPC * c); }
                            else                    { TFHE_TST16(w, tacc + 128 + WPC * c); }
                        }
                    }
                } else {
                    uint32_t w[4][WPC];
                    tmem_wait_st();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if constexpr (WPC == 8) { TFHE_TLD8(w[c], tacc + 128 + WPC * c); }
                        else                    { TFHE_TLD16(w[c], tacc + 128 + WPC * c); }
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            U ure, uim;
                            if (WPC == 8) { ure = (U)w[c][(2 * i) % WPC]; uim = (U)w[c][(2 * i + 1) % WPC]; }
                            else { ure = (U)(((uint64_t)w[c][(4 * i + 1) % WPC] << 32) | w[c][(4 * i) % WPC]);
                                   uim = (U)(((uint64_t)w[c][(4 * i + 3) % WPC] << 32) | w[c][(4 * i + 2) % WPC]); }
                            v[4 * c + i] = make_double2(digit_of(ure, sh0, mask, half), digit_of(uim, sh0, mask, half));
                            u[4 * c + i] = make_double2(digit_of(ure, sh1, mask, half), digit_of(uim, sh1, mask, half));
                        }
                }
                KeyWindow<KW> kw;
#pragma unroll
                for (int s = 0; s < KW; s++) keywin_load<LOGM, KW>(kw, kb, s);
                tree_forward2<LOGM, !NOTT9>(v, u, buf, tw, t, bar_id, ttw);
                mac2_tmem<LOGM, KW>(p == 0, tacc, v, u, kw, kb);
            } else {
                // odd l: the last level alone (reloaded from the stash when there is one)
                const int sh = W - (lev + 1) * Bgbit;
                cplx v[16];
                if (lev > 0) {
                    uint32_t w[4][WPC];
                    tmem_wait_st();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if constexpr (WPC == 8) { TFHE_TLD8(w[c], tacc + 128 + WPC * c); }
                        else                    { TFHE_TLD16(w[c], tacc + 128 + WPC * c); }
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            U ure, uim;
                            if (WPC == 8) { ure = (U)w[c][(2 * i) % WPC]; uim = (U)w[c][(2 * i + 1) % WPC]; }
                            else { ure = (U)(((uint64_t)w[c][(4 * i + 1) % WPC] << 32) | w[c][(4 * i) % WPC]);
                                   uim = (U)(((uint64_t)w[c][(4 * i + 3) % WPC] << 32) | w[c][(4 * i + 2) % WPC]); }
                            v[4 * c + i] = make_double2(digit_of(ure, sh, mask, half), digit_of(uim, sh, mask, half));
                        }
                } else {
                    int a2 = a, tj = t;
                    asm volatile("" : "+r"(a2), "+r"(tj));
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int j = tj + T * i;
                        const U ure = (U)(PLAIN ? aq[j] : rot_minus_one<Torus, N>(aq, j, a2)) + offset;
                        const U uim = (U)(PLAIN ? aq[j + M] : rot_minus_one<Torus, N>(aq, j + M, a2)) + offset;
                        v[i] = make_double2(digit_of(ure, sh, mask, half), digit_of(uim, sh, mask, half));
                    }
                }
                if (p == 0) forward_and_mac<LOGM, true, KM_REGS2, NOTT9>(v, tacc, bk, buf, kp_unused, tw, t, bar_id, ttw);
                else        forward_and_mac<LOGM, false, KM_REGS2, NOTT9>(v, tacc, bk + (size_t)(p * 2) * M, buf, kp_unused, tw, t, bar_id, ttw);
            }
        }
    }
    {
        cplx R0[16], R1[16];
        load_tmem2(R0, R1, tacc);
        tree_backward2<LOGM, true, !NOTT9>(R0, R1, buf, tw, t, bar_id, ttw);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = t + T * m;
            acc[j] = (Torus)((PLAIN ? (U)0 : (U)acc[j]) + (U)to_torus(R0[m].x, (Torus)0));
            acc[j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[j + M]) + (U)to_torus(R0[m].y, (Torus)0));
            acc[N + j] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j]) + (U)to_torus(R1[m].x, (Torus)0));
            acc[N + j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j + M]) + (U)to_torus(R1[m].y, (Torus)0));
        }
    }
    lanes_sync<T>(bar_id);      // accumulator writes visible before the next step's rotated reads
}

// ---------------------------------------------------------------------------------------------
// CEILING PROBE (not a product path; TFHE_B200_BR_VARIANT=m, wrong results by construction): what a warp sustains when it does
// nothing but the FP64 work of a CMUX -- two digit polynomials at a time, rotated differences "already there" in its tensor-memory
// columns, key values "already there" in tensor memory too (224..479), results left in tensor memory -- i.e. the compute half of a
// producer / consumer split in which helper warps would do the integer and key-fetch work.  One such warp per SM sub-partition
// (4 per SM).  profiles/r2_notes.md has the measurement and what it says about that design.
// ---------------------------------------------------------------------------------------------
template <int LOGM>
__device__ __forceinline__ void cmux_step_mock(const int l, const int Bgbit, cplx* __restrict__ buf, const uint32_t tacc, const uint32_t tkey,
                                               const uint32_t kstride, const cplx* __restrict__ tw, const int t, const int bar_id, const uint32_t ttw) {
    const uint32_t mask = (1u << Bgbit) - 1u;
    const int half = 1 << (Bgbit - 1);
#pragma unroll 1
    for (int q = 0; q < 2; q++) {
        const int sh0 = 32 - Bgbit, sh1 = sh0 - Bgbit;
        cplx v[16], u[16];
        {
            uint32_t w[4][8];
            tmem_wait_st();
#pragma unroll
            for (int c = 0; c < 4; c++) TFHE_TLD8(w[c], tacc + 128 + 8 * c);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t ure = w[c][2 * i], uim = w[c][2 * i + 1];
                    v[4 * c + i] = make_double2(digit_of(ure, sh0, mask, half), digit_of(uim, sh0, mask, half));
                    u[4 * c + i] = make_double2(digit_of(ure, sh1, mask, half), digit_of(uim, sh1, mask, half));
                }
        }
        tree_forward2<LOGM, true>(v, u, buf, tw, t, bar_id, ttw);
        if (q) tmem_wait_st();
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t r0[16], r1[16], k[4][16];
#pragma unroll
            for (int x = 0; x < 4; x++) TFHE_TLD16(k[x], tkey + kstride * x + 16 * c);
            if (q) { TFHE_TLD16(r0, tacc + 16 * c); TFHE_TLD16(r1, tacc + 64 + 16 * c); }
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int s = 4 * c + i;
                cplx R0 = !q ? make_double2(0.0, 0.0) : make_double2(__hiloint2double((int)r0[4 * i + 1], (int)r0[4 * i]), __hiloint2double((int)r0[4 * i + 3], (int)r0[4 * i + 2]));
                cplx R1 = !q ? make_double2(0.0, 0.0) : make_double2(__hiloint2double((int)r1[4 * i + 1], (int)r1[4 * i]), __hiloint2double((int)r1[4 * i + 3], (int)r1[4 * i + 2]));
                cplx kk[4];
#pragma unroll
                for (int x = 0; x < 4; x++) kk[x] = make_double2(__hiloint2double((int)k[x][4 * i + 1], (int)k[x][4 * i]), __hiloint2double((int)k[x][4 * i + 3], (int)k[x][4 * i + 2]));
                cfma(R0, v[s], kk[0]); cfma(R1, v[s], kk[1]); cfma(R0, u[s], kk[2]); cfma(R1, u[s], kk[3]);
                r0[4 * i] = (uint32_t)__double2loint(R0.x); r0[4 * i + 1] = (uint32_t)__double2hiint(R0.x);
                r0[4 * i + 2] = (uint32_t)__double2loint(R0.y); r0[4 * i + 3] = (uint32_t)__double2hiint(R0.y);
                r1[4 * i] = (uint32_t)__double2loint(R1.x); r1[4 * i + 1] = (uint32_t)__double2hiint(R1.x);
                r1[4 * i + 2] = (uint32_t)__double2loint(R1.y); r1[4 * i + 3] = (uint32_t)__double2hiint(R1.y);
            }
            TFHE_TST16(r0, tacc + 16 * c);
            TFHE_TST16(r1, tacc + 64 + 16 * c);
        }
    }
    {
        cplx R0[16], R1[16];
        load_tmem2(R0, R1, tacc);
        tree_backward2<LOGM, true, true>(R0, R1, buf, tw, t, bar_id, ttw);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t w0[8], w1[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                w0[2 * i] = (uint32_t)double_to_torus32(R0[4 * c + i].x); w0[2 * i + 1] = (uint32_t)double_to_torus32(R0[4 * c + i].y);
                w1[2 * i] = (uint32_t)double_to_torus32(R1[4 * c + i].x); w1[2 * i + 1] = (uint32_t)double_to_torus32(R1[4 * c + i].y);
            }
            TFHE_TST8(w0, tacc + 8 * c);
            TFHE_TST8(w1, tacc + 64 + 8 * c);
        }
    }
}

// One CMUX: ACC <- ACC + BK_i (x) ((X^a - 1) ACC).   acc: shared [2][N].  bk: BK_i = [2l][2][M] spectra (scaled 2/N).
// STASH: the rotated difference u = (X^a - 1) ACC_q + offset of a coefficient is formed ONCE per q (level 0: two shared-memory
// reads, index and sign arithmetic) and parked in this lane's tensor-memory columns [128, 128 + 32 words); levels 1.. only
// reload it and cut their digit.  Otherwise every level re-reads the accumulator.
// PLAIN: the external product alone, ACC <- BK (x) ACC (tGswFFTExternMulToTLwe, cb/tgsw_functions.cpp:424-449): no rotation
// on the way in, no accumulation on the way out.
// The depth-0 butterfly of the forward transform multiplies its `hi` input by w(0,0) = (1 + i) / sqrt 2:
//     w (x + i y) = ((x - y) + i (x + y)) / sqrt 2.
// x and y are small integer digits here, so x - y and x + y are formed BEFORE the conversion to double, on the integer pipe, and the
// butterfly shrinks from 6 to 4 FMAs (pass16's PRE0 form): -64 FP64 instructions per CMUX at l = 2.  Registers 8..15 of a lane (c >= 2)
// are the `hi` inputs of depth 0.
template <bool PRE0> __device__ __forceinline__ cplx digit_pair(const int c, const int dre, const int dim) {
    if (PRE0 && c >= 2) return make_double2((double)(dre - dim), (double)(dre + dim));
    return make_double2((double)dre, (double)dim);
}
template <int LOGM, typename Torus, bool STASH, int KM, bool PLAIN = false>
__device__ __forceinline__ void cmux_step(Torus* __restrict__ acc, const int a, const cplx* __restrict__ bk,
                                          const int l, const int Bgbit, cplx* __restrict__ buf, const uint32_t tacc,
                                          KeyPipe& kp, const cplx* __restrict__ tw, const int t, const int bar_id, const uint32_t ttw = 0) {
    typedef TreePlan<LOGM> P;
    typedef typename TorusTraits<Torus>::U U;
    constexpr int M = P::M, N = P::N, T = P::T, W = TorusTraits<Torus>::W;
    constexpr int WPC = StashWords<Torus>::PER_C;
    // Torus32 digits without mask and bias (FAST32): the field of level lev sits at bits [sh, sh + Bgbit) of u = x + offset and the
    // digit is field - Bg/2 = the field with its top bit flipped, sign-extended.  Adding 2^(sh+Bgbit-1) flips that bit (the carry leaves
    // the field upwards), so with 2^31 folded into the offset level 0 is ONE arithmetic shift, and level lev >= 1 is
    // (int32)(u * 2^(lev Bgbit) + 2^31) >> (32 - Bgbit): the multiply-add moves the field to the top (dropping the 2^31 folded in for level
    // 0) and flips its top bit.  Same values as ((u >> sh) & mask) - half, two integer instructions fewer per coefficient.
    constexpr bool FAST32 = sizeof(Torus) == 4 && BR_FASTDIGIT;
    // digit pairs pre-combined for the depth-0 butterfly (digit_pair): N = 2048 only -- 188.8 -> 186.3 ms per 4,096 circuit bootstraps;
    // the N = 1024 kernel got SLOWER with 64 FP64 instructions fewer per CMUX (356.1 vs 351.1 ms, profiles/r2_notes.md)
    constexpr bool PRE0 = KM == KM_REGS2 && BR_PRE0 && sizeof(Torus) == 8;
    const U offset = (U)(decomp_offset((U)0, l, Bgbit) + (FAST32 ? (U)0x80000000u : (U)0));
    const uint32_t mask = (1u << Bgbit) - 1u;
    const int half = 1 << (Bgbit - 1);
    const bool stash = STASH && l > 1;

#pragma unroll 1
    for (int p = 0; p < 2 * l; p++) {
        const int q = p >= l, lev = p - q * l;
        const Torus* __restrict__ aq = acc + q * N;
        const int sh = W - (lev + 1) * Bgbit;
        cplx v[16];
        TL(0);
        if (KM == KM_TMEM && (t & 31) == 0) kp.poll();                 // keep the key stream moving (bk_pipe.cuh)
        if (STASH && lev > 0) {
            uint32_t w[4][WPC];
            tmem_wait_st();
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if constexpr (WPC == 8) { TFHE_TLD8(w[c], tacc + 128 + WPC * c); }
                else                    { TFHE_TLD16(w[c], tacc + 128 + WPC * c); }
            }
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    U ure, uim;
                    if (WPC == 8) { ure = (U)w[c][2 * i]; uim = (U)w[c][2 * i + 1]; }
                    else { ure = (U)(((uint64_t)w[c][(4 * i + 1) % WPC] << 32) | w[c][(4 * i) % WPC]);
                           uim = (U)(((uint64_t)w[c][(4 * i + 3) % WPC] << 32) | w[c][(4 * i + 2) % WPC]); }
                    if constexpr (FAST32) {
                        const uint32_t mul = 1u << (lev * Bgbit);
                        v[4 * c + i] = digit_pair<PRE0>(c, ((int32_t)((uint32_t)ure * mul + 0x80000000u) >> (32 - Bgbit)), ((int32_t)((uint32_t)uim * mul + 0x80000000u) >> (32 - Bgbit)));
                    } else
                    v[4 * c + i] = digit_pair<PRE0>(c, ((int)((uint32_t)(ure >> sh) & mask) - half), ((int)((uint32_t)(uim >> sh) & mask) - half));
                }
        } else {
            int a2 = a;
            asm volatile("" : "+r"(a2));          // keep the 32 rotated addresses from being hoisted out of the loop
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t w[WPC];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = t + T * (4 * c + i);
                    const U ure = (U)(PLAIN ? aq[j] : rot_minus_one<Torus, N>(aq, j, a2)) + offset;
                    const U uim = (U)(PLAIN ? aq[j + M] : rot_minus_one<Torus, N>(aq, j + M, a2)) + offset;
                    if constexpr (FAST32 && STASH) {          // with a stash this branch only ever sees level 0
                        v[4 * c + i] = digit_pair<PRE0>(c, ((int32_t)(uint32_t)ure >> (32 - Bgbit)), ((int32_t)(uint32_t)uim >> (32 - Bgbit)));
                    } else if constexpr (FAST32) {
                        const uint32_t mul = 1u << (lev * Bgbit), add = lev ? 0x80000000u : 0u;
                        v[4 * c + i] = digit_pair<PRE0>(c, ((int32_t)((uint32_t)ure * mul + add) >> (32 - Bgbit)), ((int32_t)((uint32_t)uim * mul + add) >> (32 - Bgbit)));
                    } else
                    v[4 * c + i] = digit_pair<PRE0>(c, ((int)((uint32_t)(ure >> sh) & mask) - half), ((int)((uint32_t)(uim >> sh) & mask) - half));
                    if (STASH) {
                        if (WPC == 8) { w[(2 * i) % WPC] = (uint32_t)ure; w[(2 * i + 1) % WPC] = (uint32_t)uim; }
                        else { w[(4 * i) % WPC] = (uint32_t)ure; w[(4 * i + 1) % WPC] = (uint32_t)((uint64_t)ure >> 32);
                               w[(4 * i + 2) % WPC] = (uint32_t)uim; w[(4 * i + 3) % WPC] = (uint32_t)((uint64_t)uim >> 32); }
                    }
                }
                if (stash) {
                    if constexpr (WPC == 8) { TFHE_TST8(w, tacc + 128 + WPC * c); }
                    else                    { TFHE_TST16(w, tacc + 128 + WPC * c); }
                }
            }
            // (no wait::st here: the reload at the next level waits)
        }
        TL(1);
        // depth-9 twiddles (N = 2048) come from shared memory only where the Torus64 stash has taken their tensor-memory columns and the
        // twiddles are not shared per lane quarter
        constexpr bool NOTT9 = STASH && sizeof(Torus) == 8 && !(BR_SHARE_TW64 && KM == KM_REGS2);
        if (p == 0) forward_and_mac<LOGM, true, KM, NOTT9, PRE0>(v, tacc, bk, buf, kp, tw, t, bar_id, ttw);
        else        forward_and_mac<LOGM, false, KM, NOTT9, PRE0>(v, tacc, bk + (size_t)(p * 2) * M, buf, kp, tw, t, bar_id, ttw);
    }
    // every lane has finished reading the accumulator once it passes the first sync inside the backward transform
    if constexpr (KM == KM_REGS2) {
        // both polynomials together: no key values are in flight here, so there is room for two data sets (tree_fft.cuh)
        cplx R0[16], R1[16];
        TL(0);
        load_tmem2(R0, R1, tacc);
        TL(9);
        tree_backward2<LOGM, true, !(STASH && sizeof(Torus) == 8 && !BR_SHARE_TW64)>(R0, R1, buf, tw, t, bar_id, ttw);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = t + T * m;
            acc[j] = (Torus)((PLAIN ? (U)0 : (U)acc[j]) + (U)to_torus(R0[m].x, (Torus)0));
            acc[j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[j + M]) + (U)to_torus(R0[m].y, (Torus)0));
            acc[N + j] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j]) + (U)to_torus(R1[m].x, (Torus)0));
            acc[N + j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j + M]) + (U)to_torus(R1[m].y, (Torus)0));
        }
        TL(14);
    } else {
    {
        cplx R[16];
        TL(0);
        load_tmem(R, tacc);
        TL(9);
        tree_backward<LOGM, KM == KM_REGS2>(R, buf, tw, t, bar_id, ttw);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = t + T * m;
            acc[j] = (Torus)((PLAIN ? (U)0 : (U)acc[j]) + (U)to_torus(R[m].x, (Torus)0));
            acc[j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[j + M]) + (U)to_torus(R[m].y, (Torus)0));
        }
        TL(14);
    }
    {
        cplx R[16];
        load_tmem(R, tacc + 64);
        TL(9);
        tree_backward<LOGM, KM == KM_REGS2>(R, buf, tw, t, bar_id, ttw);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const int j = t + T * m;
            acc[N + j] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j]) + (U)to_torus(R[m].x, (Torus)0));
            acc[N + j + M] = (Torus)((PLAIN ? (U)0 : (U)acc[N + j + M]) + (U)to_torus(R[m].y, (Torus)0));
        }
        TL(14);
    }
    }
    lanes_sync<T>(bar_id);      // accumulator writes visible before the next step's rotated reads
    TL(15);
}

// Shared memory: twiddles | CTA control (KeyPipeShared, TMEM base) | key staging (KeyPipe only) | per group: transpose buffer, ACC
// Tensor memory : per warp R0 (64 columns) | R1 (64) | stash (STASH only), warps of a lane quarter side by side; the key columns
//                 of the KeyPipe configuration come after the last warp's window.
template <int LOGM, typename Torus, int GROUPS, bool STASH, int KM> struct BRSmem {
    typedef TreePlan<LOGM> P;
    static constexpr size_t TW_BYTES = sizeof(cplx) * ((P::TW_TOTAL + 7) & ~7);        // 128-byte multiple
    static constexpr size_t CTRL_BYTES = 128;
    static constexpr size_t CHUNK_BYTES = sizeof(cplx) * 2 * P::M;                      // BK_i[p][0..1]
    static constexpr size_t STAGE_BYTES = KM == KM_TMEM ? CHUNK_BYTES : 0;
    static constexpr size_t BUF_BYTES = (sizeof(cplx) * P::BUF + 127) & ~(size_t)127;  // transpose buffer
    static constexpr size_t ACC_BYTES = sizeof(Torus) * 2 * P::N;
    static constexpr size_t GROUP_BYTES = BUF_BYTES + ACC_BYTES;
    static constexpr size_t GROUPS_OFF = TW_BYTES + CTRL_BYTES + STAGE_BYTES;
    static constexpr size_t TOTAL = GROUPS_OFF + GROUPS * GROUP_BYTES;
    static constexpr int WARPS = GROUPS * P::T / 32;
    static_assert(TOTAL <= 232448, "shared memory budget (227 KB) exceeded");
    static constexpr bool TWT = KM == KM_REGS2;                                         // per-lane twiddles in tensor memory (tree_fft.cuh)
    static constexpr int TW_COL = 128 + (STASH ? 4 * StashWords<Torus>::PER_C : 0);
    // depth-9 twiddles (N = 2048) stay in shared memory when the 64-column stash of Torus64 needs their place
    // SHARE: warps w and w + 4 sit in the same lane quarter and their lanes have the same t, hence the same twiddles -- one copy per
    // quarter behind the two warps' windows.  Used where the windows are full otherwise (Torus64: R 128 | stash 64, twice, + 96 = 480).
    static constexpr bool SHARE = BR_SHARE_TW64 && TWT && STASH && sizeof(Torus) == 8 && WARPS == 8;
    static constexpr bool TT9 = TWT && P::NS > 1 && (!(STASH && sizeof(Torus) == 8) || SHARE);
    static constexpr int TW_COLS = TWT ? 32 * (2 + (TT9 ? 1 : 0)) : 0;
    static constexpr int TMEM_COLS = TW_COL + (SHARE ? 0 : TW_COLS);                    // per-warp window
    static constexpr int SHARED_TW_COL = 2 * TW_COL;                                     // SHARE: the quarter's twiddles
    static constexpr int KEY_COL = (WARPS + 3) / 4 * TMEM_COLS;
    static_assert(KEY_COL + (KM == KM_TMEM ? 128 : 0) + (SHARE ? TW_COLS : 0) <= 512, "tensor memory columns exceeded");
    static_assert(KM != KM_TMEM || P::T == 32, "KeyPipe: one warp per accumulator");
};

// rotation amount i of sample ct (i == n: the b part), straight from the kernel's inputs -- nothing is staged in shared memory
template <int LOGM, typename Torus>
__device__ __forceinline__ int fetch_bara(const BRArgs& A, const int ct, const int i) {
    const int n = A.n;
    if (A.mode == BR_LWE) {
        if (sizeof(Torus) == 4) {
            // tfhe_bootstrap_woKS_FFT :416-419 on x = (0,cconst) + ka*xa + kb*xb  (boots* linear part)
            uint32_t x = (uint32_t)A.ka * (uint32_t)__ldg(A.xa + (size_t)ct * (n + 1) + i);
            if (A.xb) x += (uint32_t)A.kb * (uint32_t)__ldg(A.xb + (size_t)ct * (n + 1) + i);
            if (i == n) x += (uint32_t)A.cconst;
            return modswitch32((int32_t)x, LOGM + 2);
        }
        // circuitBootstrapWoKS: abar[n0+1] already mod-switched (cb/poc_CircuitBootstrapping.cpp:542,581)
        return __ldg(A.bara + (size_t)ct * (n + 1) + i);
    }
    if (i == n) return A.mode == BR_TESTVEC ? __ldg(A.barb + ct) : 0;
    return __ldg(A.bara + (size_t)ct * n + i);
}

template <int LOGM, typename Torus, int GROUPS, bool STASH, int KM, int F2 = 0>
__global__ void __launch_bounds__(GROUPS * TreePlan<LOGM>::T, 1) blind_rotate_kernel(const BRArgs A) {
    typedef TreePlan<LOGM> P;
    typedef typename TorusTraits<Torus>::U U;
    typedef BRSmem<LOGM, Torus, GROUPS, STASH, KM> S;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T, warp = threadIdx.x >> 5;
    const int bar_id = 1 + g;
    KeyPipeShared* kps = reinterpret_cast<KeyPipeShared*>(smem_raw + S::TW_BYTES);
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(smem_raw + S::TW_BYTES + 64);
    unsigned char* gbase = smem_raw + S::GROUPS_OFF + (size_t)g * S::GROUP_BYTES;
    cplx* buf = reinterpret_cast<cplx*>(gbase);
    Torus* acc = reinterpret_cast<Torus*>(gbase + S::BUF_BYTES);

    // unit = (sample, test-vector index); n_mu > 1 only on the circuit-bootstrap path
    const int n_mu = A.n_mu > 0 ? A.n_mu : 1;
    const long total_units = (long)A.count * n_mu;
    const long unit = (long)blockIdx.x * GROUPS + g;

    for (int i = threadIdx.x; i < P::TW_TOTAL; i += GROUPS * T) tw[i] = A.tw[i];
    if (threadIdx.x == 0) {
        const long left = total_units - (long)blockIdx.x * GROUPS;
        mbar_init(&kps->tma_bar, 1); mbar_init(&kps->full_bar, 1);
        kps->done = 0; kps->tma_next = 1; kps->copy_next = 1;
        kps->active = (uint32_t)(left < GROUPS ? left : GROUPS) * (T / 32);
        mbar_fence_init();
    }
    if (warp == 0) {     // the whole SM's tensor memory: one CTA per SM (shared-memory bound), nobody else wants it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;
    const uint32_t tacc = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)S::TMEM_COLS;   // R0 | R1 | stash
    const uint32_t ttw = !S::TWT ? 0u : S::SHARE ? tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)S::SHARED_TW_COL : tacc + (uint32_t)S::TW_COL;
    if (S::TWT) tree_twiddles_to_tmem<LOGM, S::TT9 || LOGM == 9>(tw, t, ttw);
    KeyPipe kp{kps, smem_raw + S::TW_BYTES + S::CTRL_BYTES, reinterpret_cast<const unsigned char*>(A.bkfft),
               tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)S::KEY_COL, tmem_base + (uint32_t)S::KEY_COL,
               (uint32_t)S::CHUNK_BYTES, (uint32_t)(A.n * 2 * A.l), 0u};
    if (KM == KM_TMEM && threadIdx.x == 0) kp.prologue();

    if (unit < (long)A.count * n_mu) {                 // idle groups of the last CTA fall through to the final barrier
        const int ct = (int)(unit / n_mu), w = (int)(unit % n_mu);
        const int n = A.n;
        const int l = A.l;

        // ---- the initial accumulator
        Torus mu = (Torus)A.mu;
        if (sizeof(Torus) == 8 && A.mode == BR_LWE && A.mu_bgbit > 0) mu = (Torus)(1ull << (64 - (w + 1) * A.mu_bgbit));   // mu_w, poc:846
        if (A.mode == BR_ACCUM) {
            const Torus* src = reinterpret_cast<const Torus*>(A.accum) + (size_t)ct * 2 * N;
            for (int j = t; j < 2 * N; j += T) acc[j] = src[j];
        } else {
            // testvectbis = X^(2N - barb) * v  (cb/lwe_functions.cpp:385; defect D3 of the PoC corrected the same way)
            const int barb = fetch_bara<LOGM, Torus>(A, ct, n);
            const int rot = (2 * N - barb) & (2 * N - 1);
            const Torus* v = reinterpret_cast<const Torus*>(A.v);
            for (int j = t; j < N; j += T) {
                const int idx = (j - rot) & (2 * N - 1);
                const int k0 = idx & (N - 1);
                Torus val;
                if (A.mode == BR_TESTVEC) val = v[k0];
                else if (sizeof(Torus) == 4) val = mu;                              // [mu,...,mu]  :422
                else val = (k0 < M) ? (Torus)(0 - (U)(mu / 2)) : (Torus)(mu / 2);   // -mu/2 | +mu/2  poc:552-553
                acc[j] = 0;                                                          // tLweNoiselessTrivial
                acc[N + j] = (idx & N) ? (Torus)(0 - (U)val) : val;
            }
        }
        lanes_sync<T>(bar_id);
        TL(-1);

        // ---- n CMUX steps (tfhe_blindRotate_FFT :348-354); a step with bara == 0 is skipped like the reference does (:350).
        // The next rotation amount is fetched one step ahead.
        const size_t bk_stride = (size_t)2 * l * 2 * M;
        int a_next = fetch_bara<LOGM, Torus>(A, ct, 0);
        for (int i = 0; i < n; i++) {
            const int a = a_next;
            if (i + 1 < n) a_next = fetch_bara<LOGM, Torus>(A, ct, i + 1);
            if (a == 0) {
                if (KM == KM_TMEM) for (int p = 0; p < 2 * l; p++) { kp.acquire(t & 31); kp.release(t & 31); }     // stay aligned with the key stream
                continue;
            }
            if constexpr (F2 == 9) cmux_step_mock<LOGM>(l, A.Bgbit, buf, tacc, tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (GROUPS == 4 ? 224u : 448u),
                                                        GROUPS == 4 ? 64u : 0u, tw, t, bar_id, ttw);
            else if constexpr (F2 > 0) cmux_step2<LOGM, Torus, F2>(acc, a, A.bkfft + (size_t)i * bk_stride, l, A.Bgbit, buf, tacc, tw, t, bar_id, ttw);
            else cmux_step<LOGM, Torus, STASH, KM>(acc, a, A.bkfft + (size_t)i * bk_stride, l, A.Bgbit, buf, tacc, kp, tw, t, bar_id, ttw);
        }

        TL(99);
        // ---- epilogue
        if (A.mode == BR_ACCUM) {
            Torus* dst = reinterpret_cast<Torus*>(A.accum) + (size_t)ct * 2 * N;
            for (int j = t; j < 2 * N; j += T) dst[j] = acc[j];
        } else {
            // tLweExtractLweSampleIndex(index 0) cb/tlwe_functions.cpp:351-363 ; PoC adds mu/2 to b (:648)
            Torus* out = reinterpret_cast<Torus*>(A.out) + (size_t)unit * A.out_stride;
            for (int j = t; j < N; j += T) out[j] = (j == 0) ? acc[0] : (Torus)(0 - (U)acc[N - j]);
            if (t == 0) {
                U b = (U)acc[N];
                if (sizeof(Torus) == 8 && A.mode == BR_LWE) b += (U)(mu / 2);
                out[N] = (Torus)b;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------
// One external product per sample: accum_b <- G_b (x) accum_b  (tGswFFTExternMulToTLwe, cb/tgsw_functions.cpp:424-449).
// Same lane-group machinery as a blind-rotation step (decomposition, 2l forward transforms, multiply-accumulates in tensor
// memory, 2 backward transforms); the TGSW spectra are per sample (circuit-bootstrapped selectors of a LUT)
// or shared (stride 0).
// ---------------------------------------------------------------------------------------------
template <int LOGM, typename Torus, int GROUPS>
__global__ void __launch_bounds__(GROUPS * TreePlan<LOGM>::T, 1) extern_mul_kernel(const BRArgs A) {
    typedef TreePlan<LOGM> P;
    typedef BRSmem<LOGM, Torus, GROUPS, true, KM_REGS2> S;
    constexpr int N = P::N, T = P::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    const int g = threadIdx.x / T, t = threadIdx.x % T, warp = threadIdx.x >> 5;
    const int bar_id = 1 + g;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(smem_raw + S::TW_BYTES + 64);
    unsigned char* gbase = smem_raw + S::GROUPS_OFF + (size_t)g * S::GROUP_BYTES;
    cplx* buf = reinterpret_cast<cplx*>(gbase);
    Torus* acc = reinterpret_cast<Torus*>(gbase + S::BUF_BYTES);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += GROUPS * T) tw[i] = A.tw[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;
    const uint32_t tacc = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)S::TMEM_COLS;
    const long unit = (long)blockIdx.x * GROUPS + g;
    if (unit < A.count) {
        Torus* io = reinterpret_cast<Torus*>(A.accum) + (size_t)unit * 2 * N;
        for (int j = t; j < 2 * N; j += T) acc[j] = io[j];
        lanes_sync<T>(bar_id);
        KeyPipe kp{};       // unused on the register-prefetch path
        const uint32_t ttw = tacc + (uint32_t)S::TW_COL;
        tree_twiddles_to_tmem<LOGM>(tw, t, ttw);
        cmux_step<LOGM, Torus, true, KM_REGS2, true>(acc, 1, A.bkfft + (size_t)(unit / (A.units_per_gsw > 0 ? A.units_per_gsw : 1)) * A.bk_sample_stride, A.l, A.Bgbit, buf, tacc, kp, tw, t, bar_id, ttw);
        for (int j = t; j < 2 * N; j += T) io[j] = acc[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}
cudaError_t launch_extern_mul32(const BRArgs& a, cudaStream_t s) {
    constexpr int G = 8;
    typedef BRSmem<9, int32_t, G, true, KM_REGS2> S;
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(extern_mul_kernel<9, int32_t, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL);
        if (e != cudaSuccess) return e;
        attr_done.done();
    }
    if (a.count <= 0) return cudaSuccess;
    extern_mul_kernel<9, int32_t, G><<<(a.count + G - 1) / G, G * TreePlan<9>::T, S::TOTAL, s>>>(a);
    return cudaGetLastError();
}

#ifdef BR_TIMELINE
}  // namespace tfhe_b200
extern "C" __attribute__((visibility("default"))) int tfhe_b200_dev_timeline(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, tfhe_b200::g_tl_acc, sizeof(long long) * 32 * 32);
}
namespace tfhe_b200 {
#endif
static PerDeviceOnce g_inited;
// Configurations (profiles/r1_notes.md has the sweep):
//   default:  8 warps per SM, key prefetched into registers, stash                    <9, int32,  8, true,  KM_REGS2> 174 k/s
//   half   : 12 warps per SM, one key polynomial's worth of registers (KM_REGS1), stash <9, int32, 12, true,  KM_REGS1> 164 k/s
//            (the LSU pipe, not the warp count, is the wall: 12 warps with the key through the LSU gain nothing over 8)
//   keytm  : 12 warps per SM, key through tensor memory (KeyPipe), no stash           <9, int32, 12, false, KM_TMEM>  163 k/s
//            (free-running 12 warps with the key "already there" measured 201 k/s; the CTA-wide lockstep on the single
//             tensor-memory key buffer and the 2 200-cycle copy issue per chunk give that back -- kept selectable for round 2)
constexpr int G32 = 8, G32_KP = 12, G64 = 4;     // accumulators per CTA (N=1024: one warp each; N=2048: two warps each)

template <int LOGM, typename Torus, int GROUPS, bool STASH, int KM, int F2 = 0> static cudaError_t br_attr() {
    return cudaFuncSetAttribute(blind_rotate_kernel<LOGM, Torus, GROUPS, STASH, KM, F2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)BRSmem<LOGM, Torus, GROUPS, STASH, KM>::TOTAL);
}
template <int LOGM, typename Torus, int GROUPS, bool STASH, int KM, int F2 = 0> static cudaError_t br_launch(const BRArgs& a, long units, cudaStream_t s) {
    const int grid = (int)((units + GROUPS - 1) / GROUPS);
    blind_rotate_kernel<LOGM, Torus, GROUPS, STASH, KM, F2><<<grid, GROUPS * TreePlan<LOGM>::T, BRSmem<LOGM, Torus, GROUPS, STASH, KM>::TOTAL, s>>>(a);
    return cudaGetLastError();
}
// The measured alternatives (keytm, half, nostash, pairs, the ceiling probe) are instantiated only in development builds
// (-DBR_EXPERIMENTS=1, tools/build_alt.sh): the product library carries the two kernels it runs.  Asking for a variant the library
// was built without is an error, not a silent fallback.
#ifndef BR_EXPERIMENTS
#define BR_EXPERIMENTS 0
#endif
cudaError_t blind_rotate_init() {
    cudaError_t e;
    if ((e = br_attr<9, int32_t, G32, true, KM_REGS2>()) != cudaSuccess) return e;
    if ((e = br_attr<10, int64_t, G64, true, KM_REGS2>()) != cudaSuccess) return e;
#if BR_EXPERIMENTS
    if ((e = br_attr<9, int32_t, G32_KP, false, KM_TMEM>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, G32_KP, true, KM_REGS1>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, G32, false, KM_REGS2>()) != cudaSuccess) return e;
    if ((e = br_attr<10, int64_t, G64, false, KM_REGS2>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, 4, true, KM_REGS2, 9>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, G32, true, KM_REGS2, 9>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, G32, true, KM_REGS2, 2>()) != cudaSuccess) return e;
    if ((e = br_attr<9, int32_t, G32, true, KM_REGS2, 3>()) != cudaSuccess) return e;
    if ((e = br_attr<10, int64_t, G64, true, KM_REGS2, 3>()) != cudaSuccess) return e;
#endif
    g_inited.done();
    return cudaSuccess;
}

cudaError_t launch_blind_rotate32(const BRArgs& a, cudaStream_t s) {
    if (g_inited.need()) { cudaError_t e = blind_rotate_init(); if (e != cudaSuccess) return e; }
    if (a.count <= 0) return cudaSuccess;
    // development knob: TFHE_B200_BR_VARIANT = keytm | half | nostash | 2 | 3 | m | M selects the measured alternatives
    static const char* variant = getenv("TFHE_B200_BR_VARIANT");
    if (variant && variant[0] && variant[0] != 's' && variant[0] != 'd') {
#if BR_EXPERIMENTS
        if (variant[0] == 'k') return br_launch<9, int32_t, G32_KP, false, KM_TMEM>(a, a.count, s);
        if (variant[0] == 'h') return br_launch<9, int32_t, G32_KP, true, KM_REGS1>(a, a.count, s);
        if (variant[0] == 'n') return br_launch<9, int32_t, G32, false, KM_REGS2>(a, a.count, s);
        // "pairs": two digit polynomials per pass (cmux_step2), rolling key window of 2 / 3 slots.  Measured 380 / 382 ms against 354 ms
        // for the default (profiles/r2_notes.md): what the side-by-side transforms gain (short_scoreboard 10.3 -> 5.9 %) the exposed
        // key latency in the multiply-accumulate gives back (long_scoreboard 1.4 -> 7.6 %), and 255 registers leave a few spills.
        if (variant[0] == '2') return br_launch<9, int32_t, G32, true, KM_REGS2, 2>(a, a.count, s);
        if (variant[0] == '3') return br_launch<9, int32_t, G32, true, KM_REGS2, 3>(a, a.count, s);
        if (variant[0] == 'm') return br_launch<9, int32_t, 4, true, KM_REGS2, 9>(a, a.count, s);           // ceiling probe, 4 warps per SM
        if (variant[0] == 'M') return br_launch<9, int32_t, G32, true, KM_REGS2, 9>(a, a.count, s);         // ceiling probe, 8 warps per SM
        // (a variant that deferred the multiply-accumulates of polynomial p into the transpose of p+1 -- spectrum parked in tensor memory --
        //  was correct and 22 % slower: profiles/r2_notes.md, commit "Experiment (negative): multiply-accumulates ... deferred")
#endif
        return cudaErrorNotSupported;        // unknown variant, or a library built without -DBR_EXPERIMENTS=1
    }
    return br_launch<9, int32_t, G32, true, KM_REGS2>(a, a.count, s);
}
cudaError_t launch_blind_rotate64(const BRArgs& a, cudaStream_t s) {
    if (g_inited.need()) { cudaError_t e = blind_rotate_init(); if (e != cudaSuccess) return e; }
    if (a.count <= 0) return cudaSuccess;
    const long units = (long)a.count * (a.n_mu > 0 ? a.n_mu : 1);
    // Torus64: the stash takes 64 columns, so the depth-9 twiddles stay in shared memory (R 128 | stash 64 | twiddles 64 = 256 columns).
    // With the waits deferred the stash pays here too (189 vs 201 ms per 4096 circuit bootstraps; before that it cost 3 %).
#if BR_EXPERIMENTS
    static const char* variant = getenv("TFHE_B200_BR_VARIANT");
    if (variant && variant[0] == 'n') return br_launch<10, int64_t, G64, false, KM_REGS2>(a, units, s);      // "nostash"
    if (variant && variant[0] == '3') return br_launch<10, int64_t, G64, true, KM_REGS2, 3>(a, units, s);    // "pairs": 196 vs 188 ms
#endif
    return br_launch<10, int64_t, G64, true, KM_REGS2>(a, units, s);
}

// ---------------------------------------------------------------------------------------------
// Standalone transforms (one lane group per polynomial, 256 threads per CTA)
// ---------------------------------------------------------------------------------------------
template <int LOGM> struct TrCfg {
    typedef TreePlan<LOGM> P;
    static constexpr int GROUPS = 256 / P::T;
    static constexpr size_t SMEM = sizeof(cplx) * (((P::TW_TOTAL + 1) & ~1) + GROUPS * P::BUF);
};

template <int LOGM, typename Torus>
__global__ void __launch_bounds__(256) poly_to_spectrum_kernel(cplx* __restrict__ out, const Torus* __restrict__ in,
                                                               const cplx* __restrict__ twg, int count, double scale) {
    typedef TreePlan<LOGM> P;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += 256) tw[i] = twg[i];
    __syncthreads();
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    cplx* buf = tw + ((P::TW_TOTAL + 1) & ~1) + g * P::BUF;
    const long poly = (long)blockIdx.x * TrCfg<LOGM>::GROUPS + g;
    if (poly >= count) return;
    const Torus* src = in + (size_t)poly * N;
    cplx v[16];
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const int j = t + T * m;
        // execute_reverse_int: exact int->double (:33-46); execute_reverse_torus64: C cast, round to 53 bits (:166-170)
        v[m] = make_double2((double)src[j], (double)src[j + M]);
    }
    tree_forward<LOGM>(v, buf, tw, t, 1 + g);
    // the stored spectrum is the TRUE one: divide the per-slot unit factor of the select-free exchange out (TreePlan::TG)
    cplx* dst = out + (size_t)poly * M + t;
    const cplx* gf = twg + P::TG + t;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const cplx gi = gf[i * T];          // v * conj(g) * scale
        dst[i * T] = make_double2((v[i].x * gi.x + v[i].y * gi.y) * scale, (v[i].y * gi.x - v[i].x * gi.y) * scale);
    }
}

template <int LOGM, typename Torus>
__global__ void __launch_bounds__(256) spectrum_to_torus_kernel(Torus* __restrict__ out, const cplx* __restrict__ in,
                                                                const cplx* __restrict__ twg, int count, double scale) {
    typedef TreePlan<LOGM> P;
    constexpr int M = P::M, N = P::N, T = P::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* tw = reinterpret_cast<cplx*>(smem_raw);
    for (int i = threadIdx.x; i < P::TW_TOTAL; i += 256) tw[i] = twg[i];
    __syncthreads();
    const int g = threadIdx.x / T, t = threadIdx.x % T;
    cplx* buf = tw + ((P::TW_TOTAL + 1) & ~1) + g * P::BUF;
    const long poly = (long)blockIdx.x * TrCfg<LOGM>::GROUPS + g;
    if (poly >= count) return;
    const cplx* src = in + (size_t)poly * M + t;
    cplx v[16];
    const cplx* gf = twg + P::TG + t;
#pragma unroll
    for (int i = 0; i < 16; i++) {          // 2/N pre-scale (:78-100), times the slot's unit factor g the backward tree expects
        const cplx x = src[i * T], gi = gf[i * T];
        v[i] = make_double2((x.x * gi.x - x.y * gi.y) * scale, (x.x * gi.y + x.y * gi.x) * scale);
    }
    tree_backward<LOGM>(v, buf, tw, t, 1 + g);
    Torus* dst = out + (size_t)poly * N;
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const int j = t + T * m;
        dst[j] = to_torus(v[m].x, (Torus)0);
        dst[j + M] = to_torus(v[m].y, (Torus)0);
    }
}

template <int LOGM, typename Torus>
static cudaError_t launch_p2s(cplx* out, const Torus* in, const cplx* tw, int count, double scale, cudaStream_t s) {
    auto kern = poly_to_spectrum_kernel<LOGM, Torus>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TrCfg<LOGM>::SMEM);
    if (e != cudaSuccess) return e;
    if (count <= 0) return cudaSuccess;
    const int G = TrCfg<LOGM>::GROUPS;
    kern<<<(count + G - 1) / G, 256, TrCfg<LOGM>::SMEM, s>>>(out, in, tw, count, scale);
    return cudaGetLastError();
}
template <int LOGM, typename Torus>
static cudaError_t launch_s2t(Torus* out, const cplx* in, const cplx* tw, int count, double scale, cudaStream_t s) {
    auto kern = spectrum_to_torus_kernel<LOGM, Torus>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TrCfg<LOGM>::SMEM);
    if (e != cudaSuccess) return e;
    if (count <= 0) return cudaSuccess;
    const int G = TrCfg<LOGM>::GROUPS;
    kern<<<(count + G - 1) / G, 256, TrCfg<LOGM>::SMEM, s>>>(out, in, tw, count, scale);
    return cudaGetLastError();
}

cudaError_t launch_poly_to_spectrum32(cplx* out, const int32_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_p2s<9, int32_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_p2s<10, int32_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_poly_to_spectrum64(cplx* out, const int64_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_p2s<9, int64_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_p2s<10, int64_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_spectrum_to_torus32(int32_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_s2t<9, int32_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_s2t<10, int32_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}
cudaError_t launch_spectrum_to_torus64(int64_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s) {
    if (N == 1024) return launch_s2t<9, int64_t>(out, in, tw, count, scale, s);
    if (N == 2048) return launch_s2t<10, int64_t>(out, in, tw, count, scale, s);
    return cudaErrorInvalidValue;
}

// res += a (.) b, complex, slot by slot (LagrangeHalfCPolynomialAddMulASM, cb/spqlios/lagrangehalfc_impl_fma.s:78-135)
__global__ void spectrum_addmul_kernel(cplx* __restrict__ res, const cplx* __restrict__ a, const cplx* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        cplx r = res[i];
        cfma(r, a[i], b[i]);
        res[i] = r;
    }
}
cudaError_t launch_spectrum_addmul(cplx* res, const cplx* a, const cplx* b, size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    spectrum_addmul_kernel<<<grid, 256, 0, s>>>(res, a, b, n);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
