// ks_kernels.cu -- batched key switching as a tiled gather-accumulate on the CUDA cores.
// (Selected with TFHE_B200_KS=cuda.  The default path is the tensor-core kernel of ks_tc_kernels.cu, 5x faster and bit-identical;
//  this is the form the design started from and the yardstick the tensor-core kernel was measured against.)
//
// Replaces lweKeySwitch / lweKeySwitchTranslate_fromArray (cb/lwe_functions.cpp:136-171), preKeySwitch
// (cb/poc_CircuitBootstrapping.cpp:437-465) and circuitPrivKS (:667-698):
//     result = (0,b) - sum_{i<rows, j<t} key[i][j][ d_ij ],   d_ij = ((a_i + prec_offset) >> (W-(j+1)basebit)) & (base-1), d_ij != 0
// The reference walks the table once per sample (12.3 MB of rows per gate, 147 MB per private key switch).
// Here a CTA owns a tile of 32 (KSCfg::TILE) samples x 512 output columns: every (i,j) block of base-1 candidate rows
// arrives ONCE per CTA (TMA bulk copy into a shared-memory ring) and each sample of the tile subtracts the row its
// digit selects, so table traffic per sample drops by ~the tile height.
//
// Device key layout: int32 [cols_pad/512][rows][t][base-1][512]  (d = 0 rows are never read by the reference either).
#include "engine.h"
#include "bk_pipe.cuh"
#include <type_traits>
#include <cstdlib>

namespace tfhe_b200 {

#ifndef KS_S_DEF
#define KS_S_DEF 8
#endif
#ifndef KS_CTAS
#define KS_CTAS 2
#endif
#ifndef KS_VOTE
#define KS_VOTE 0             // 1: digit tests through a warp vote (a uniform predicate by construction)
#endif
#ifndef KS_NSTAGE4
#define KS_NSTAGE4 10         // ring depth (blocks) of the base-4 instance
#endif
#ifndef KS_LEAVE_DEP
#define KS_LEAVE_DEP 1        // count-out ordered by operand dependencies instead of a memory barrier (see leave())
#endif
#ifndef KS_NSTAGE16
#define KS_NSTAGE16 6         // ring depth of the paired (base-16) instance
#endif
#ifndef KS_WARPS16
#define KS_WARPS16 16
#endif
#ifndef KS_CTAS16
#define KS_CTAS16 1
#endif
#ifndef KS_NSTAGE8
#define KS_NSTAGE8 6          // base-8 instance (private key switch of the circuit bootstrap): ring depth, warps per CTA, CTAs per SM
#endif
#ifndef KS_WARPS8
#define KS_WARPS8 8
#endif
#ifndef KS_CTAS8
#define KS_CTAS8 2
#endif
#ifndef KS_UNIFORM
#define KS_UNIFORM 1          // digits through a warp reduction into uniform registers (see the loop); 0 = round-1 form
#endif
constexpr int KS_S = KS_S_DEF;    // samples per thread
constexpr int KS_ICHUNK = 32;     // input coefficients staged per refill of a warp's digit source

// BASEBIT = 4 is not a parameter set of the reference: it is the PAIRED form of a base-4 key (ks_repack_pair_kernel below), two
// consecutive base-4 digits of a coefficient taken as one base-16 digit whose row is the sum of the two rows.  Half as many
// blocks and row subtractions per sample, no compare tree (each sample reads the one row its digit selects); the blocks are
// 30 KB, so that instance runs 16 warps (64 samples) on a 6-deep ring, one CTA per SM.
template <int BASEBIT> struct KSCfg {
    static constexpr int BASE = 1 << BASEBIT;
    static constexpr int STAGE_INTS = (BASE - 1) * 512;            // one (i,j) block: the base-1 candidate rows of this CTA's 512 columns
    static constexpr int STAGE_BYTES = STAGE_INTS * 4;
    static constexpr int NSTAGE = BASEBIT == 4 ? KS_NSTAGE16 : BASEBIT == 3 ? KS_NSTAGE8 : (BASEBIT == 2 ? KS_NSTAGE4 : 16);
    static constexpr int WARPS = BASEBIT == 4 ? KS_WARPS16 : BASEBIT == 3 ? KS_WARPS8 : 8;    // sample groups x 2 column halves
    static constexpr int CTAS = BASEBIT == 4 ? KS_CTAS16 : BASEBIT == 3 ? KS_CTAS8 : KS_CTAS;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int TILE = (WARPS / 2) * KS_S;                // samples per CTA
    // rows are copied to registers and selected by a warp-uniform branch while base-1 < KS_S; for base 8 there are as many
    // rows as samples per thread, so each sample reads the row it selects straight from the ring instead.
    static constexpr bool ROWS_IN_REGS = BASEBIT <= 2;
};
template <typename U, int BASEBIT> constexpr size_t ks_smem_bytes() {
    return (size_t)KSCfg<BASEBIT>::NSTAGE * KSCfg<BASEBIT>::STAGE_BYTES + KSCfg<BASEBIT>::NSTAGE * sizeof(uint64_t) +
           (size_t)KSCfg<BASEBIT>::NSTAGE * 32 * sizeof(uint32_t) + (size_t)KSCfg<BASEBIT>::WARPS * KS_S * (KS_ICHUNK + 1) * sizeof(U) + 128;
}

__device__ __forceinline__ void sub8(int4& a0, int4& a1, const int4& r0, const int4& r1) {
    a0.x -= r0.x; a0.y -= r0.y; a0.z -= r0.z; a0.w -= r0.w;
    a1.x -= r1.x; a1.y -= r1.y; a1.z -= r1.z; a1.w -= r1.w;
}

__device__ __forceinline__ uint32_t smem_inc_acq_rel(uint32_t* p) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
    return old;
}

// The key is one linear stream per 512-column group ([group][i][j][d][512] int32), walked identically by every CTA of that
// group.  It is pulled into a shared-memory ring with 1-D TMA bulk copies, NSTAGE blocks ahead of its use.  There is no
// producer warp and no CTA-wide barrier in the loop: each warp counts itself out of a slot when it has read it (an atomic
// whose result is only looked at after the block's arithmetic, so its round trip is hidden), and the LAST of the 8 warps
// to leave re-arms the slot's mbarrier and issues the copy of the block NSTAGE further down the stream.
// (Letting a fixed warp issue the refill after waiting on an `empty` mbarrier was 2x slower: the issuer waits for the
// slowest warp and every other warp then starves behind the late refill -- profiles/r1_notes.md.)
// Warp w owns sample group w>>1 (8 samples, so a sample's digit is warp-uniform and selecting its row is a uniform branch /
// a uniform address) and column half w&1; lane l accumulates columns [4l,4l+4) and [128+4l,128+4l+4) of that half -- two
// conflict-free LDS.128 per row.
template <typename TorusIn, int BASEBIT>
__global__ void __launch_bounds__(KSCfg<BASEBIT>::THREADS, KSCfg<BASEBIT>::CTAS) keyswitch_kernel(const KSArgs A) {
    typedef typename std::conditional<sizeof(TorusIn) == 4, uint32_t, uint64_t>::type U;
    typedef KSCfg<BASEBIT> C;
    constexpr int W = sizeof(TorusIn) * 8;
    constexpr int BASE = C::BASE;
    extern __shared__ __align__(128) unsigned char ks_smem[];
    int4* ring = reinterpret_cast<int4*>(ks_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(ks_smem + (size_t)C::NSTAGE * C::STAGE_BYTES);
    uint32_t* left = reinterpret_cast<uint32_t*>(full + C::NSTAGE);          // [slot][lane]: warps that have left each slot, one copy per lane
    U* abar_all = reinterpret_cast<U*>(left + C::NSTAGE * 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = A.rows_in * A.t;
    const int32_t* kstream = A.key + (size_t)blockIdx.z * A.key_z_stride + (size_t)blockIdx.y * nblk * C::STAGE_INTS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NSTAGE; s++) { mbar_init(full + s, 1); for (int l = 0; l < 32; l++) left[s * 32 + l] = 0; }
        mbar_fence_init();
        for (int s = 0; s < C::NSTAGE && s < nblk; s++) {
            mbar_expect_tx(full + s, C::STAGE_BYTES);
            tma_load_1d(reinterpret_cast<unsigned char*>(ring) + (size_t)s * C::STAGE_BYTES, kstream + (size_t)s * C::STAGE_INTS, C::STAGE_BYTES, full + s);
        }
    }
    __syncthreads();

    const int sg = warp >> 1, half = warp & 1;
    const int s0 = blockIdx.x * C::TILE + sg * KS_S;              // this warp's first sample
    const TorusIn* in = reinterpret_cast<const TorusIn*>(A.in);
    const U prec_offset = (U)1 << (W - (1 + BASEBIT * A.t));      // cb/lwe_functions.cpp:141 ; poc:444,674
    U (*abar)[KS_ICHUNK + 1] = reinterpret_cast<U (*)[KS_ICHUNK + 1]>(abar_all + (size_t)warp * KS_S * (KS_ICHUNK + 1));
    const int lofs = half * 64 + lane;                            // int4 index of this lane's first column quad inside a row

    int4 acc0[KS_S], acc1[KS_S];
#pragma unroll
    for (int s = 0; s < KS_S; s++) { acc0[s] = make_int4(0, 0, 0, 0); acc1[s] = make_int4(0, 0, 0, 0); }

    // leave(): count this warp out of a slot (after its reads of the slot have been issued); refill_if_last(): the warp
    // that counted out last refills the slot with the block NSTAGE further on.
    // Count-out without divergence: every lane increments ITS OWN copy of the slot's counter (32 addresses, one conflict-free
    // instruction), so all lanes of a warp see the same count and nobody branches.  With `if (lane == 0) atomicAdd` the other 31
    // lanes waited at the reconvergence point for lane 0's whole atomic round trip -- 18 % of the kernel's stall samples in round 1's
    // ncu source view (ptxas also wraps a single-lane atom.add in its vote / popc / shuffle aggregation sequence).  Here the result is a
    // scoreboarded register nobody looks at until refill_if_last, after the block's arithmetic.  atom.inc wraps the counter to 0 on
    // the last warp by itself, so there is no reset store either.
    // KS_LEAVE_DEP (default): the count-out is a RELAXED atomic that takes the values this lane has just loaded from the slot as (unused)
    // operands.  What the refill must not overtake is the slot's reads, and a read is over when its destination register is written --
    // exactly what the operand dependency waits for.  The acq_rel form compiled to MEMBAR.ALL.CTA + ATOMS behind a WARPSYNC: every warp
    // drained its whole memory pipeline once per block (profiles/r2_notes.md).  The refilling thread still issues fence.proxy.async
    // between its own count-out and the bulk copy.
    auto leave = [&](int slot, int dep) -> uint32_t {
        uint32_t tok;
#if KS_LEAVE_DEP
        // `dep` is computed from every value loaded from the slot; the volatile move below consumes it, a warp issues in order, so the
        // atomic cannot issue before those loads have written their registers.
        int sink;
        asm volatile("mov.b32 %0, %1;" : "=r"(sink) : "r"(dep));
        asm volatile("atom.relaxed.cta.shared::cta.inc.u32 %0, [%1], %2;"
                     : "=r"(tok) : "r"(smem_u32(left + slot * 32 + lane)), "r"(C::WARPS - 1) : "memory");
#else
        __syncwarp();
        asm volatile("atom.acq_rel.cta.shared::cta.inc.u32 %0, [%1], %2;" : "=r"(tok) : "r"(smem_u32(left + slot * 32 + lane)), "r"(C::WARPS - 1) : "memory");
#endif
        return tok;
    };
    auto refill_if_last = [&](uint32_t tok, int slot, int k) {
        if (lane == 0 && tok == C::WARPS - 1) {                   // lane 0 of the last warp out (the counter has wrapped to 0)
            const int kn = k + C::NSTAGE;
            if (kn < nblk) {
                fence_proxy_async_smem();
                mbar_expect_tx(full + slot, C::STAGE_BYTES);
                tma_load_1d(reinterpret_cast<unsigned char*>(ring) + (size_t)slot * C::STAGE_BYTES, kstream + (size_t)kn * C::STAGE_INTS,
                            C::STAGE_BYTES, full + slot);
            }
        }
    };

    int slot = 0, k = 0; uint32_t ph = 0;
    for (int i0 = 0; i0 < A.rows_in; i0 += KS_ICHUNK) {
        __syncwarp();
#pragma unroll
        for (int s = 0; s < KS_S; s++) {                          // lane = coefficient index: 128/256-byte coalesced rows
            const int smp = s0 + s, i = i0 + lane;
            U v = 0;
            if (smp < A.count && i < A.rows_in) v = (U)in[(size_t)smp * A.in_stride + i] + prec_offset;
            abar[s][lane] = v;
        }
        __syncwarp();
        const int iend = min(KS_ICHUNK, A.rows_in - i0);
        for (int ii = 0; ii < iend; ii++) {
            // The digits of a sample are the same in every lane (all lanes read the same abar entry), but a value loaded from shared
            // memory is a per-lane value to the compiler: the digit tests became vector compares, divergence brackets (BSSY / BSYNC)
            // and branch-resolve stalls -- 53 % of the kernel's stall samples sat on that machinery (profiles/r2_notes.md).  A warp
            // reduction (REDUX writes a UNIFORM register) tells ptxas what we know: everything from here to the branch runs on the
            // uniform datapath.  Only the top t * basebit <= 32 bits of a coefficient carry digits.
            uint32_t a[KS_S];
#pragma unroll
            for (int s = 0; s < KS_S; s++) a[s] = KS_UNIFORM == 1 ? __reduce_or_sync(0xffffffffu, (uint32_t)(abar[s][ii] >> (W - 32)))
                                                                  : (uint32_t)(abar[s][ii] >> (W - 32));
            for (int j = 0; j < A.t; j++, k++) {
                const int sh = 32 - (j + 1) * BASEBIT;
                int dg[KS_S];
#pragma unroll
                for (int s = 0; s < KS_S; s++) {
                    dg[s] = (int)((a[s] >> sh) & (uint32_t)(BASE - 1));
                    if (KS_UNIFORM == 2) dg[s] = (int)__reduce_or_sync(0xffffffffu, (uint32_t)dg[s]);      // the reduction right in front of the branch
                }
                const int4* st = ring + (size_t)slot * (C::STAGE_INTS / 4) + lofs;
                while (!mbar_try_wait(full + slot, ph)) {}
                if constexpr (C::ROWS_IN_REGS) {
                    int4 r0[BASE - 1], r1[BASE - 1];
#pragma unroll
                    for (int d = 0; d < BASE - 1; d++) { r0[d] = st[d * 128]; r1[d] = st[d * 128 + 32]; }
                    // rows are in registers: the slot can be refilled already 
                    int dep = 0;
#pragma unroll
                    for (int d = 0; d < BASE - 1; d++) dep ^= r0[d].x ^ r1[d].w;
                    const uint32_t tok = leave(slot, dep);
#pragma unroll
                    for (int s = 0; s < KS_S; s++) {
#pragma unroll
                        for (int d = 0; d < BASE - 1; d++)
                            if (KS_VOTE ? __any_sync(0xffffffffu, dg[s] == d + 1) : (dg[s] == d + 1)) sub8(acc0[s], acc1[s], r0[d], r1[d]);      // (predicated subtractions instead: 44.7 vs 36.6 ms)
                    }
                    refill_if_last(tok, slot, k);
                } else {
                    int dep = 0;
#pragma unroll
                    for (int s = 0; s < KS_S; s++) {
                        if (KS_VOTE ? __any_sync(0xffffffffu, dg[s] != 0) : (dg[s] != 0)) {
                            const int4* rp = st + (dg[s] - 1) * 128;
                            sub8(acc0[s], acc1[s], rp[0], rp[32]);
                        }
                        dep ^= acc0[s].x ^ acc1[s].w;             // (only as an ordering operand of the count-out below)
                    }
                    refill_if_last(leave(slot, dep), slot, k);
                }
                if (++slot == C::NSTAGE) { slot = 0; ph ^= 1; }
            }
        }
    }
    // result starts as the noiseless trivial sample (0,b) (cb/lwe_functions.cpp:169) or 0 (poc:677-681)
    const int colbase = blockIdx.y * 512 + half * 256 + lane * 4;
#pragma unroll
    for (int s = 0; s < KS_S; s++) {
        const int smp = s0 + s;
        if (smp >= A.count) continue;
        int32_t* orow = A.out + (size_t)blockIdx.z * A.out_z_stride + (size_t)(smp / A.group) * A.out_stride + (size_t)(smp % A.group) * A.out_inner;
        const int v[8] = {acc0[s].x, acc0[s].y, acc0[s].z, acc0[s].w, acc1[s].x, acc1[s].y, acc1[s].z, acc1[s].w};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int col = colbase + (c & 3) + (c >> 2) * 128;
            if (col < A.cols) {
                int32_t x = v[c];
                if (col == A.b_col) x += (int32_t)in[(size_t)smp * A.in_stride + A.b_index];
                orow[col] = x;
            }
        }
    }
}

template <typename TorusIn, int BASEBIT>
static cudaError_t launch_ks_b(const KSArgs& a, cudaStream_t s) {
    typedef typename std::conditional<sizeof(TorusIn) == 4, uint32_t, uint64_t>::type U;
    constexpr size_t smem = ks_smem_bytes<U, BASEBIT>();
    static PerDeviceOnce attr_done;              // per instantiation and device
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_kernel<TorusIn, BASEBIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(keyswitch_kernel<TorusIn, BASEBIT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_done.done();
    }
    dim3 grid((a.count + KSCfg<BASEBIT>::TILE - 1) / KSCfg<BASEBIT>::TILE, a.cols_pad / 512, a.nz > 0 ? a.nz : 1);
    keyswitch_kernel<TorusIn, BASEBIT><<<grid, KSCfg<BASEBIT>::THREADS, smem, s>>>(a);
    return cudaGetLastError();
}
template <typename TorusIn>
static cudaError_t launch_ks(KSArgs a, cudaStream_t s) {
    if (a.count <= 0) return cudaSuccess;
    if (a.cols_pad % 512) return cudaErrorInvalidValue;
    if (a.group <= 0) { a.group = 1; a.out_inner = 0; }
    switch (a.basebit) {
        case 1: return launch_ks_b<TorusIn, 1>(a, s);
        case 2: return launch_ks_b<TorusIn, 2>(a, s);
        case 3: return launch_ks_b<TorusIn, 3>(a, s);
        case 4: if (sizeof(TorusIn) == 4) return launch_ks_b<int32_t, 4>(a, s); return cudaErrorInvalidValue;      // paired base-4 key (gate path)
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t launch_keyswitch32_rows(const KSArgs& a, cudaStream_t s) { return launch_ks<int32_t>(a, s); }
cudaError_t launch_keyswitch64_rows(const KSArgs& a, cudaStream_t s) { return launch_ks<int64_t>(a, s); }

// raw [rows][t][base][cols] -> [cols_pad/512][rows][t][base-1][512]  (d = 0 dropped, zero padded).  src holds the blocks
// [blk0, blk0 + nblk) of nblk_total (a slice of input rows): large keys are staged through a small temporary.
__global__ void ks_repack_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, size_t nblk_total, size_t blk0, size_t nblk,
                                 int base, int cols, int cols_pad) {
    const size_t per_group = nblk * (size_t)(base - 1) * 512;
    const size_t per_group_total = nblk_total * (size_t)(base - 1) * 512;
    const size_t total = per_group * (size_t)(cols_pad / 512);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t g = e / per_group, r = e % per_group;
        const int c = (int)(g * 512 + r % 512);
        const size_t ro = r / 512;
        const size_t ij = ro / (base - 1); const int d = (int)(ro % (base - 1)) + 1;
        dst[g * per_group_total + blk0 * (size_t)(base - 1) * 512 + r] = c < cols ? src[(ij * base + d) * (size_t)cols + c] : 0;
    }
}
// raw base-4 key [rows][t][4][cols] -> PAIRED layout [cols_pad/512][rows][t/2][15][512]:
//     row D (1..15) of block (i, jj) = raw[i][2jj][D >> 2] + raw[i][2jj+1][D & 3]   (a zero digit contributes nothing)
// The key switch subtracts one row per (i, j) with a non-zero digit; subtracting the sum of two rows at once is the same integer
// arithmetic mod 2^32 in a different order (cb/lwe_functions.cpp:143-151), so results stay bit-identical.
__global__ void ks_repack_pair_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, size_t rows, int t, int cols, int cols_pad) {
    const size_t nblk = rows * (size_t)(t / 2);
    const size_t per_group = nblk * 15 * 512;
    const size_t total = per_group * (size_t)(cols_pad / 512);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t g = e / per_group, r = e % per_group;
        const int c = (int)(g * 512 + r % 512);
        const size_t ro = r / 512;
        const size_t blk = ro / 15; const int D = (int)(ro % 15) + 1;
        const size_t i = blk / (t / 2), jj = blk % (t / 2);
        const int dh = D >> 2, dl = D & 3;
        uint32_t v = 0;
        if (c < cols) {
            if (dh) v += (uint32_t)src[((i * t + 2 * jj) * 4 + dh) * (size_t)cols + c];
            if (dl) v += (uint32_t)src[((i * t + 2 * jj + 1) * 4 + dl) * (size_t)cols + c];
        }
        dst[e] = (int32_t)v;
    }
}
cudaError_t launch_ks_repack_pair(int32_t* dst, const int32_t* src, int rows, int t, int cols, int cols_pad, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    if (t % 2) return cudaErrorInvalidValue;
    ks_repack_pair_kernel<<<148 * 8, 256, 0, s>>>(dst, src, (size_t)rows, t, cols, cols_pad);
    return cudaGetLastError();
}
cudaError_t launch_ks_repack_rows_rows(int32_t* dst, const int32_t* src, int rows_total, int row0, int rows, int t, int base, int cols, int cols_pad,
                                       cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    ks_repack_kernel<<<148 * 8, 256, 0, s>>>(dst, src, (size_t)rows_total * t, (size_t)row0 * t, (size_t)rows * t, base, cols, cols_pad);
    return cudaGetLastError();
}
// (launch_ks_repack, launch_ks_repack_rows, launch_keyswitch32/64: the dispatching entry points live in ks_tc_kernels.cu)

}  // namespace tfhe_b200
