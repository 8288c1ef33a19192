// ks_tc_kernels.cu -- key switching on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory).
//
// Same function as ks_kernels.cu (lweKeySwitch cb/lwe_functions.cpp:136-171, preKeySwitch cb/poc_CircuitBootstrapping.cpp:437-465,
// circuitPrivKS :667-698):
//     result = (0,b) - sum_{i<rows, j<t} key[i][j][ d_ij ],      d_ij the base-2^basebit digits of a_i + prec_offset, d_ij != 0
// restated as an exact integer matrix product.  Per sample the digits are a ONE-HOT row: X[s][(i,j,d)] = 1 iff d_ij = d.  The key
// rows, 32-bit integers, are split into their four bytes K = K0 + 2^8 K1 + 2^16 K2 + 2^24 K3, so
//     sum_{i,j} key[i][j][d_ij][c] = sum_b 2^(8b) (X Kb)[s][c]          (mod 2^32)
// and each X Kb is a u8 x u8 -> s32 product: exact (entries <= 255, at most rows*t <= 2^23/255 terms per sum -- checked at launch),
// and the recombination mod 2^32 is the same wrap-around arithmetic the reference's additions perform, so results are bit-identical.
// The one-hot form wastes base-1 of every base multiplies, and still wins by a wide margin: the CUDA-core kernel is bound by the
// shared-memory bandwidth that feeds the selected rows to the integer adders (one 128-byte wavefront per 32 additions), the tensor core
// takes 128 samples x 256 columns (two byte planes) x 32 key rows per instruction, 128 cycles each when the pipe is kept fed
// (tools/imma_probe.cu, profiles/r2_imma_probe3.txt, profiles/r2_notes.md).
//
// One CTA = 128 samples (the M dimension = the 128 lanes of tensor memory) x 128 output columns, four s32 accumulator tiles of
// 128 columns (one per key byte) = all 512 columns of tensor memory.  A STEP covers 32 rows of the one-hot matrix = 32 / base
// consecutive (i, j) blocks (a group of `base` rows per block: base-1 candidates and one padding row that no digit selects).
//   producer warps (4 per group, TC_GROUPS groups taking turns, TC_SPT steps per turn): thread = sample: input coefficients arrive in
//                coalesced 32 x 32 tiles transposed through shared memory, a step's digits are cut from a bit stream, its 32-byte one-hot
//                row goes into the A ring (K-major, no swizzle); at the end read the four tiles back, recombine, negate, add b, store
//                (each group its share of the columns)
//   MMA warps  : TC_MMAW issuing threads taking turns step by step: the two MMAs of a step (same A, the two plane pairs of the key as B),
//                committed to the slot's `free` barrier
//   TMA warp   : one thread streams the key: one 16 KB bulk copy per step into the B ring
// Key image in global memory: [column group][step][plane pair][8192 B], every 8 KB block already in the shared-memory image the MMA wants:
// 256 rows (byte plane 2h of the 128 columns, then byte plane 2h+1) x 32 key rows of the step, K-major,
//   byte (row r, k) at (k / 16) * 4096 + (r / 8) * 128 + (r % 8) * 16 + k % 16
// i.e. 8 x 16-byte core matrices, LBO (K direction) 4096, SBO (row direction) 128 -- layouts verified by tools/imma_probe.cu.
#include "engine.h"
#include "bk_pipe.cuh"
#include <type_traits>
#include <cstdlib>

namespace tfhe_b200 {

#ifndef TC_STAGES_DEF
#define TC_STAGES_DEF 8
#endif
constexpr int TC_STAGES = TC_STAGES_DEF;  // ring depth (steps)
constexpr int TC_B_BYTES = 16384;         // key bytes per step: 2 plane pairs x 8 KB
constexpr int TC_A_BYTES = 4096;          // one-hot image per step: 128 samples x 32 B
#ifndef TC_GROUPS_DEF
#define TC_GROUPS_DEF 2
#endif
constexpr int TC_GROUPS = TC_GROUPS_DEF;  // producer warp sets taking turns step by step (1, 2 or 4)
#ifndef TC_MMAW_DEF
#define TC_MMAW_DEF 2
#endif
// MMA-issuing warps taking turns step by step.  tcgen05.commit holds its issuing thread for ~370 cycles, during which the thread queues
// nothing and the tensor pipe runs dry: one issuer committing every step gets 544 cycles per step, two get 272, three reach the pipe's
// 256 (tools/imma_probe.cu, profiles/r2_imma_probe3.txt).  A commit covers the committing thread's MMAs only, so each slot is still
// released by the thread that consumed it; the integer accumulation does not care in which order the tensor core takes the streams
// (the accumulators are zeroed up front instead of relying on a first non-accumulating MMA).
constexpr int TC_MMAW = TC_MMAW_DEF;
#ifndef TC_SPT_DEF
#define TC_SPT_DEF 4
#endif
constexpr int TC_SPT = TC_SPT_DEF;        // consecutive steps a producer group writes per turn
// A ring slot must always be served by the SAME producer group and the SAME issuing warp: each of them walks its own steps in order, so
// it is never more than one phase ahead of a slot's mbarriers.  (With 3 issuers on 8 slots an issuer could poll a barrier two uses
// ahead, where the parity test aliases -- measured as sporadic launch failures.)
static_assert(TC_STAGES % (TC_GROUPS * TC_SPT) == 0 && TC_STAGES % TC_MMAW == 0, "ring slots must map to fixed producer groups / MMA issuers");
constexpr int TC_THREADS = 128 * TC_GROUPS + 32 * TC_MMAW + 32;        // 4 producer / epilogue warps per group, the MMA warps, the TMA warp
#ifndef TC_CLUSTER_DEF
#define TC_CLUSTER_DEF 1
#endif
// CTAs per cluster (1 or 2).  2: two sample tiles of the same column group form a cluster and share ONE key stream -- each CTA fetches half
// of every step's 16 KB and multicasts it into both shared memories (TMA .multicast::cluster), the step's MMA commit releases the slot in
// both CTAs -- so the L2 -> SM traffic of the key halves.
constexpr int TC_CLUSTER = TC_CLUSTER_DEF;
constexpr size_t TC_SMEM_BASE = (size_t)TC_STAGES * (TC_B_BYTES + TC_A_BYTES) + 1024 /*alignment slack*/ + 512 /*barriers*/;
constexpr size_t TC_SMEM = TC_SMEM_BASE + (size_t)4 * TC_GROUPS * 32 * 33 * 4;     // + the producers' input tiles

// K-major, no swizzle, sm_100 descriptor version; lbo = distance between the two 16-byte K chunks (rows * 16), SBO = 128 (8-row groups)
__host__ __device__ inline uint64_t tc_desc(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// One MMA covers TWO byte planes: the B tile is 256 rows (plane 2h in rows 0-127, plane 2h+1 in rows 128-255), N = 256.  An i8 MMA
// costs ~76 cycles + 0.66 cycles per column of N (tools/imma_probe.cu: 161 cycles at N = 128, 246 at N = 256), so two wide MMAs per step
// (492 cycles) beat four narrow ones (646).
// cute::UMMA::InstrDescriptor: c_format[4,6)=2 (s32), a/b_format = 0 (u8), a/b_major = 0 (K), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4
constexpr uint32_t TC_IDESC = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {      // arrives on `bar` (same offset) in every CTA of the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_1d_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
#define TC_TLD16(r, addr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
                 : "r"(addr) : "memory")

#ifndef TC_DBG
#define TC_DBG 0              // development ablations (wrong results): 1 = no MMAs, 2 = no key copies, 4 = no digit work, 8 = no input loads
#endif
template <typename TorusIn, int BASEBIT>
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1) keyswitch_tc_kernel(const KSArgs A) {
    typedef typename std::conditional<sizeof(TorusIn) == 4, uint32_t, uint64_t>::type U;
    constexpr int W = sizeof(TorusIn) * 8;
    constexpr int BASE = 1 << BASEBIT;
    constexpr int Q = 32 / BASE;                                   // (i, j) blocks per step
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ringB = smem;
    unsigned char* ringA = smem + (size_t)TC_STAGES * TC_B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ringA + (size_t)TC_STAGES * TC_A_BYTES);
    uint64_t* full = bars;                       // key bytes of the step have landed (TMA, byte count)
    uint64_t* ready = bars + TC_STAGES;          // the four producer warps whose turn it is have written the step's one-hot rows
    uint64_t* freeb = bars + 2 * TC_STAGES;      // the step's MMAs have completed: both ring slots may be overwritten
    uint64_t* done = bars + 3 * TC_STAGES;       // all MMAs have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);
    unsigned char* tiles = reinterpret_cast<unsigned char*>(bars) + 512;       // per producer warp: 32 samples x 32 coefficients (+1 pad) x 4 B

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = A.rows_in * A.t;
    const int nsteps = (nblk + Q - 1) / Q;
    // CTA order.  gridDim.y > 1: sample tiles vary fastest -- the CTAs of a wave walk the same key stream together (a key image that does
    // not fit the L2 is then read from HBM once).  gridDim.y == 1: column groups vary fastest -- the CTAs that share a sample tile run
    // together and its input rows come from HBM once (key image L2-resident: the gate key switch read its 268 MB of input four times
    // before, 1.5 GB of DRAM traffic per launch against 0.46 GB algorithmic).
    const int ncg = A.cols_pad / 128;
    const int cgrp = gridDim.y > 1 ? (int)blockIdx.y : (int)(blockIdx.x % ncg);
    const long tile_idx = gridDim.y > 1 ? (long)blockIdx.x : (long)(blockIdx.x / ncg);
    const unsigned char* kstream = reinterpret_cast<const unsigned char*>(A.key) + (size_t)blockIdx.z * A.key_z_stride * sizeof(int32_t) +
                                   (size_t)cgrp * nsteps * TC_B_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full + s, 1); mbar_init(ready + s, 4); mbar_init(freeb + s, TC_CLUSTER); }
        mbar_init(done, TC_MMAW);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TC_CLUSTER > 1) cluster_sync_all();              // the peer's barriers exist before anything is multicast at them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t crank = TC_CLUSTER > 1 ? cluster_ctarank() : 0u;
    if (warp < 4) {                                       // zero the four accumulator tiles (every MMA accumulates)
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; i++) z[i] = 0u;
        const uint32_t tq = tmem + (((uint32_t)warp * 32u) << 16);
#pragma unroll 1
        for (int c = 0; c < 512; c += 16)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
                         ::"r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]),
                           "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(tq + c) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 4 * TC_GROUPS + TC_MMAW) {
        // ---- key stream
        if (lane == 0) {
            for (int st = 0; st < nsteps; st++) {
                const int slot = st % TC_STAGES, use = st / TC_STAGES;
                if (use > 0) tc_wait(freeb + slot, (uint32_t)(use - 1) & 1u);
                if (TC_DBG & 2) { mbar_arrive(full + slot); continue; }
                mbar_expect_tx(full + slot, TC_B_BYTES);
                if (TC_CLUSTER > 1) {
                    constexpr uint32_t PART = TC_B_BYTES / TC_CLUSTER;
                    tma_load_1d_multicast(ringB + (size_t)slot * TC_B_BYTES + crank * PART, kstream + (size_t)st * TC_B_BYTES + crank * PART, PART,
                                          full + slot, (uint16_t)((1u << TC_CLUSTER) - 1u));
                } else
                tma_load_1d(ringB + (size_t)slot * TC_B_BYTES, kstream + (size_t)st * TC_B_BYTES, TC_B_BYTES, full + slot);
            }
        }
    } else if (warp >= 4 * TC_GROUPS) {
        // ---- MMA issue: issuer k takes the steps st = k (mod TC_MMAW)
        if (lane == 0) {
            for (int st = warp - 4 * TC_GROUPS; st < nsteps; st += TC_MMAW) {
                const int slot = st % TC_STAGES; const uint32_t ph = (uint32_t)(st / TC_STAGES) & 1u;
                tc_wait(full + slot, ph);
                tc_wait(ready + slot, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = tc_desc(smem_u32(ringA + (size_t)slot * TC_A_BYTES), 2048);
                const uint32_t sb = smem_u32(ringB + (size_t)slot * TC_B_BYTES);
#pragma unroll
                for (int h = 0; h < 2; h++) if (!(TC_DBG & 1)) tc_mma_i8(tmem + 256u * h, da, tc_desc(sb + 8192u * h, 4096), 1u);
                if (TC_CLUSTER > 1) tc_commit_multicast(freeb + slot, (uint16_t)((1u << TC_CLUSTER) - 1u));
                else tc_commit(freeb + slot);
            }
            tc_commit(done);
        }
    } else {
        // ---- one-hot rows: TC_GROUPS threads per sample, thread (m, g) builds the rows of sample m for the steps st = g (mod TC_GROUPS).
        //      What a producer warp spends per step is not arithmetic but a chain of small latencies (slot wait, branches on the block
        //      counters, store, proxy fence, warp sync, arrive): ~900 cycles per step measured -- with one warp set doing every step that
        //      chain, not the tensor core (492 cycles per step), set the pace, and splitting each ROW between two threads changed nothing
        //      (profiles/r2_notes.md).  Taking turns step by step gives every warp TC_GROUPS step times per row.
        const int m = threadIdx.x & 127, g = threadIdx.x >> 7;
        const long smp = tile_idx * 128 + m;
        const bool live = smp < A.count;
        const TorusIn* in = reinterpret_cast<const TorusIn*>(A.in) + (size_t)(live ? smp : 0) * A.in_stride;
        const U prec_offset = (U)1 << (W - (1 + BASEBIT * A.t));      // cb/lwe_functions.cpp:141 ; poc:444,674
        const uint32_t row_off = (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u;
        // The digits of a sample form one BIT STREAM: coefficient i contributes its top t * basebit bits (of a_i + prec_offset), digit 0
        // first, and a step consumes the next Q * basebit bits of it.  The stream runs through a 64-bit buffer (valid bits at the top),
        // refilled one coefficient at a time, so a step's digits come out by constant shifts -- no per-block counters and branches
        // (those, ~25 cycles of branch latency per block, were what bounded the base-4 instances: 8 blocks per step).
        // Input: a thread walking its own sample row touches one 32-byte sector per load and 32 different lines per warp instruction --
        // those loads alone cost a third of the gate key switch and a quarter of the private one (ablations in profiles/r2_notes.md).
        // Each warp reads its 32 samples in tiles of 32 coefficients with lane = coefficient (one coalesced row segment per
        // instruction), keeps the top 32 bits of a + prec_offset, and transposes the tile through shared memory; the next tile waits in
        // registers while the current one is used.
        const int cbits = BASEBIT * A.t;                              // bits per coefficient
        constexpr int SBITS = Q * BASEBIT;                            // bits per step
        int i_next = 0;                                               // next coefficient to enter the buffer
        uint64_t bitbuf = 0; int nbits = 0;
        uint32_t* tile = reinterpret_cast<uint32_t*>(tiles) + (size_t)warp * (32 * 33);
        const long s0w = tile_idx * 128 + (m & ~31);                  // first sample of this warp
        const TorusIn* inw = reinterpret_cast<const TorusIn*>(A.in);
        uint32_t nxt[32];                                             // next tile: coefficient 32 c + lane of the warp's 32 samples
        auto fetch_tile = [&](int c) {
            const int i = 32 * c + lane;
#pragma unroll
            for (int r = 0; r < 32; r++) {
                U v = (U)0;
                if (!(TC_DBG & 8) && s0w + r < A.count && i < A.rows_in) v = (U)inw[(size_t)(s0w + r) * A.in_stride + i];
                nxt[r] = (uint32_t)((v + prec_offset) >> (W - 32));
            }
        };
        auto store_tile = [&]() {
            __syncwarp();                                             // everybody has read the old tile
#pragma unroll
            for (int r = 0; r < 32; r++) tile[r * 33 + lane] = nxt[r];
            __syncwarp();
        };
        fetch_tile(0); store_tile(); fetch_tile(1);
        auto refill = [&]() {
            if (i_next && (i_next & 31) == 0) { store_tile(); fetch_tile((i_next >> 5) + 1); }       // (warp-uniform: all lanes refill together)
            uint32_t a32 = 0u;                                        // past the last coefficient: zero digits (the tail of the last step)
            if (live && i_next < A.rows_in) a32 = tile[(m & 31) * 33 + (i_next & 31)];
            i_next++;
            bitbuf |= (uint64_t)(a32 >> (32 - cbits)) << (64 - nbits - cbits);
            nbits += cbits;
        };
        auto drop = [&](int n) {                                      // skip n bits of the stream (the other groups' steps)
            while (n > 0) {
                if (nbits == 0) refill();
                const int k = n < nbits ? n : nbits;
                bitbuf <<= k; nbits -= k; n -= k;
            }
        };
        // A group takes TC_SPT consecutive steps per turn: the slot wait, the stores, ONE proxy fence, one warp sync and the arrives are a
        // chain of latencies (~500 cycles) that is paid per turn, not per step.
        drop(g * TC_SPT * SBITS);                                     // group g starts at step g * TC_SPT
        const uint32_t arow0 = smem_u32(ringA) + row_off;
        for (int st0 = g * TC_SPT; st0 < nsteps; st0 += TC_GROUPS * TC_SPT) {
            uint32_t w[TC_SPT][8];
#pragma unroll
            for (int e = 0; e < TC_SPT; e++) {
                if (!(TC_DBG & 4)) { while (nbits < SBITS) refill(); }
#pragma unroll
                for (int x = 0; x < 8; x++) w[e][x] = 0u;
                const uint32_t top = (uint32_t)(bitbuf >> 32);        // SBITS <= 16: the step's digits sit in the top word
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const uint32_t d = (top >> (32 - (q + 1) * BASEBIT)) & (uint32_t)(BASE - 1);
                    // one-hot byte d-1 of this block's group of BASE bytes (nothing for d = 0)
                    if constexpr (BASE == 4) w[e][q] = (1u << (8 * d)) >> 8;
                    else if constexpr (BASE == 8) {
                        const uint64_t v = d ? (uint64_t)1 << (8 * (d - 1)) : (uint64_t)0;
                        w[e][2 * q] = (uint32_t)v; w[e][2 * q + 1] = (uint32_t)(v >> 32);
                    } else {                                        // BASE == 2: two bytes per block, candidate byte first
                        w[e][q >> 1] |= d << (16 * (q & 1));
                    }
                }
                bitbuf <<= SBITS; nbits -= SBITS;
            }
            if (!(TC_DBG & 4)) drop((TC_GROUPS - 1) * TC_SPT * SBITS);
#pragma unroll
            for (int e = 0; e < TC_SPT; e++) {
                const int st = st0 + e;
                if (st < nsteps) {
                    const int slot = st % TC_STAGES, use = st / TC_STAGES;
                    if (use > 0) tc_wait(freeb + slot, (uint32_t)(use - 1) & 1u);
                    const uint32_t arow = arow0 + (uint32_t)slot * TC_A_BYTES;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(arow), "r"(w[e][0]), "r"(w[e][1]), "r"(w[e][2]), "r"(w[e][3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(arow + 2048u), "r"(w[e][4]), "r"(w[e][5]), "r"(w[e][6]), "r"(w[e][7]) : "memory");
                }
            }
            fence_proxy_async_smem();                               // generic-proxy stores -> visible to the tensor core's (async proxy) reads
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int e = 0; e < TC_SPT; e++) if (st0 + e < nsteps) mbar_arrive(ready + (st0 + e) % TC_STAGES);
            }
        }
        // ---- epilogue: D[sample][column] of byte plane p sits in lane = sample, column 128 p + column
        tc_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tl = tmem + (((uint32_t)(warp & 3) * 32u) << 16);        // warps w and w + 4 reach the same lane quarter
        int32_t* orow = nullptr;
        if (live) orow = A.out + (size_t)blockIdx.z * A.out_z_stride + (size_t)(smp / A.group) * A.out_stride + (size_t)(smp % A.group) * A.out_inner;
        const int colbase = cgrp * 128;
#pragma unroll 1
        for (int c = (128 / TC_GROUPS) * g; c < (128 / TC_GROUPS) * (g + 1); c += 16) {       // each of the sample's threads stores its share of the columns
            uint32_t p0[16], p1[16], p2[16], p3[16];
            TC_TLD16(p0, tl + c); TC_TLD16(p1, tl + 128 + c); TC_TLD16(p2, tl + 256 + c); TC_TLD16(p3, tl + 384 + c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
#pragma unroll
                for (int x = 0; x < 16; x++) {
                    const int col = colbase + c + x;
                    if (col < A.cols) {
                        uint32_t v = 0u - (p0[x] + (p1[x] << 8) + (p2[x] << 16) + (p3[x] << 24));
                        if (col == A.b_col) v += (uint32_t)in[A.b_index];       // starts as the noiseless trivial sample (0,b) (cb/lwe_functions.cpp:169)
                        orow[col] = (int32_t)v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    if (TC_CLUSTER > 1) cluster_sync_all();              // nobody leaves while the peer may still signal its barriers
}

template <typename TorusIn, int BASEBIT>
static cudaError_t launch_ks_tc_b(const KSArgs& a, cudaStream_t s) {
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_tc_kernel<TorusIn, BASEBIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
        if (e != cudaSuccess) return e;
        attr_done.done();
    }
    const int tiles = ((a.count + 127) / 128 + TC_CLUSTER - 1) / TC_CLUSTER * TC_CLUSTER, ncg = a.cols_pad / 128;
    // key image per launch small enough to stay in the L2 (and more than one column group, no clusters): column groups fastest
    const bool cg_fast = TC_CLUSTER == 1 && ncg > 1 && ks_key_bytes(a.rows_in, a.t, a.basebit, a.cols_pad) * (size_t)(a.nz > 0 ? a.nz : 1) <= ((size_t)96 << 20) &&
                         (long)tiles * ncg <= 0x7fffffffL;
    dim3 grid(cg_fast ? tiles * ncg : tiles, cg_fast ? 1 : ncg, a.nz > 0 ? a.nz : 1);
    keyswitch_tc_kernel<TorusIn, BASEBIT><<<grid, TC_THREADS, TC_SMEM, s>>>(a);
    return cudaGetLastError();
}
template <typename TorusIn>
static cudaError_t launch_ks_tc(KSArgs a, cudaStream_t s) {
    if (a.count <= 0) return cudaSuccess;
    if (a.cols_pad % 128) return cudaErrorInvalidValue;
    if ((long)a.rows_in * a.t * 255 >= (1l << 31)) return cudaErrorInvalidValue;       // s32 accumulators must not wrap
    if (a.group <= 0) { a.group = 1; a.out_inner = 0; }
    switch (a.basebit) {
        case 1: return launch_ks_tc_b<TorusIn, 1>(a, s);
        case 2: return launch_ks_tc_b<TorusIn, 2>(a, s);
        case 3: return launch_ks_tc_b<TorusIn, 3>(a, s);
        default: return cudaErrorInvalidValue;
    }
}

// raw [rows][t][base][cols] (a slice [row0, row0 + rows) of rows_total input rows) -> byte-plane images, see the header.
// One thread per (block, candidate d, column): reads one key word, writes its four bytes into the four planes.  dst must have been
// zeroed (padding rows, padding columns and the tail of the last step stay zero).
__global__ void ks_tc_repack_kernel(unsigned char* __restrict__ dst, const int32_t* __restrict__ src, size_t rows_total, size_t row0, size_t rows,
                                    int t, int base, int cols, int cols_pad) {
    const int Q = 32 / base;
    const size_t nsteps = (rows_total * (size_t)t + Q - 1) / Q;
    const size_t total = rows * (size_t)t * (size_t)(base - 1) * (size_t)cols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % cols);
        size_t r = e / cols;
        const int d = (int)(r % (base - 1)) + 1; r /= (base - 1);
        const size_t lblk = r;                                        // block inside the slice: (i - row0) * t + j
        const uint32_t v = (uint32_t)src[(lblk * base + d) * (size_t)cols + c];
        const size_t blk = row0 * (size_t)t + lblk;
        const size_t step = blk / Q; const int q = (int)(blk % Q);
        const int k = q * base + d - 1;
        const int n = c & 127; const size_t cg = (size_t)c >> 7;
        // byte plane pl of column n = row (pl & 1) * 128 + n of pair tile pl >> 1; a pair tile is [K chunk 2][row group 32][8 rows][16 B]
        const size_t off = (size_t)(k >> 4) * 4096 + (size_t)(n >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 15);
        unsigned char* img = dst + (cg * nsteps + step) * (size_t)TC_B_BYTES + off;
        img[0] = (unsigned char)v; img[2048] = (unsigned char)(v >> 8); img[8192] = (unsigned char)(v >> 16); img[8192 + 2048] = (unsigned char)(v >> 24);
    }
}

// ------------------------------------------------------------------ packing choice and the dispatching entry points
int ks_packing() {
    static const int mode = [] {
        const char* e = getenv("TFHE_B200_KS");
        return (e && (e[0] == 'c' || e[0] == 'C')) ? (int)KS_PACK_ROWS : (int)KS_PACK_TC;       // "cuda": the CUDA-core kernels of ks_kernels.cu
    }();
    return mode;
}
size_t ks_key_bytes(int rows, int t, int basebit, int cols_pad) {
    const int base = 1 << basebit;
    if (ks_packing() == KS_PACK_TC) {
        const int Q = 32 / base;
        const size_t nsteps = ((size_t)rows * t + Q - 1) / Q;
        return (size_t)(cols_pad / 128) * nsteps * TC_B_BYTES;
    }
    return (size_t)rows * t * (base - 1) * (size_t)cols_pad * sizeof(int32_t);
}
cudaError_t launch_ks_repack_rows_rows(int32_t* dst, const int32_t* src, int rows_total, int row0, int rows, int t, int base, int cols, int cols_pad, cudaStream_t s);
cudaError_t launch_keyswitch32_rows(const KSArgs& a, cudaStream_t s);
cudaError_t launch_keyswitch64_rows(const KSArgs& a, cudaStream_t s);

cudaError_t launch_ks_repack_rows(int32_t* dst, const int32_t* src, int rows_total, int row0, int rows, int t, int base, int cols, int cols_pad,
                                  cudaStream_t s) {
    if (ks_packing() != KS_PACK_TC) return launch_ks_repack_rows_rows(dst, src, rows_total, row0, rows, t, base, cols, cols_pad, s);
    if (rows <= 0) return cudaSuccess;
    if (cols_pad % 128 || 32 % base) return cudaErrorInvalidValue;
    int basebit = 0; while ((1 << basebit) < base) basebit++;
    if (row0 == 0) {
        cudaError_t e = cudaMemsetAsync(dst, 0, ks_key_bytes(rows_total, t, basebit, cols_pad), s);
        if (e != cudaSuccess) return e;
    }
    ks_tc_repack_kernel<<<148 * 16, 256, 0, s>>>(reinterpret_cast<unsigned char*>(dst), src, (size_t)rows_total, (size_t)row0, (size_t)rows, t, base, cols, cols_pad);
    return cudaGetLastError();
}
cudaError_t launch_ks_repack(int32_t* dst, const int32_t* src, int rows, int t, int base, int cols, int cols_pad, cudaStream_t s) {
    return launch_ks_repack_rows(dst, src, rows, 0, rows, t, base, cols, cols_pad, s);
}
cudaError_t launch_keyswitch32(const KSArgs& a, cudaStream_t s) {
    if (ks_packing() == KS_PACK_TC && a.basebit <= 3) return launch_ks_tc<int32_t>(a, s);
    return launch_keyswitch32_rows(a, s);
}
cudaError_t launch_keyswitch64(const KSArgs& a, cudaStream_t s) {
    if (ks_packing() == KS_PACK_TC && a.basebit <= 3) return launch_ks_tc<int64_t>(a, s);
    return launch_keyswitch64_rows(a, s);
}

}  // namespace tfhe_b200
