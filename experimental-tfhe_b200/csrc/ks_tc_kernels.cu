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
// takes 128 samples x 128 columns x 32 key rows per instruction (tools/imma_probe.cu: 120 cycles per 128x128x32 MMA, profiles/r2_notes.md).
//
// One CTA = 128 samples (the M dimension = the 128 lanes of tensor memory) x 128 output columns, four s32 accumulator tiles of
// 128 columns (one per key byte) = all 512 columns of tensor memory.  A STEP covers 32 rows of the one-hot matrix = 32 / base
// consecutive (i, j) blocks (a group of `base` rows per block: base-1 candidates and one padding row that no digit selects).
//   warps 0-3  : sample s = thread: read a_i, cut the digits, write the step's 32-byte one-hot row into the A ring (K-major, no swizzle);
//                at the end read the four tiles back, recombine, negate, add b, store
//   warp 4     : one thread issues the four MMAs of a step (same A, the four byte planes of the key as B) and commits them to the
//                slot's `free` barrier
//   warp 5     : one thread streams the key: one 16 KB bulk copy (TMA) per step into the B ring
// Key image in global memory: [column group][step][plane][4096 B], every 4 KB block already in the shared-memory image the MMA wants
//   byte (n, k) at (k / 16) * 2048 + (n / 8) * 128 + (n % 8) * 16 + k % 16        (n = column in the group, k = row of the step)
// i.e. 8 x 16-byte core matrices, LBO (K direction) 2048, SBO (column direction) 128 -- verified by tools/imma_probe.cu.
#include "engine.h"
#include "bk_pipe.cuh"
#include <type_traits>
#include <cstdlib>

namespace tfhe_b200 {

constexpr int TC_STAGES = 8;              // ring depth (steps)
constexpr int TC_B_BYTES = 16384;         // key bytes per step: 4 planes x 4 KB
constexpr int TC_A_BYTES = 4096;          // one-hot image per step: 128 samples x 32 B
constexpr int TC_THREADS = 192;
constexpr size_t TC_SMEM = (size_t)TC_STAGES * (TC_B_BYTES + TC_A_BYTES) + 1024 /*alignment slack*/ + 256 /*barriers*/;

__host__ __device__ inline uint64_t tc_desc(uint32_t saddr) {          // K-major, no swizzle, LBO 2048 / SBO 128, sm_100 descriptor version
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// cute::UMMA::InstrDescriptor: c_format[4,6)=2 (s32), a/b_format = 0 (u8), a/b_major = 0 (K), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4
constexpr uint32_t TC_IDESC = (2u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
#define TC_TLD16(r, addr)                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                     \
                 : "r"(addr) : "memory")

template <typename TorusIn, int BASEBIT>
__global__ void __launch_bounds__(TC_THREADS, 1) keyswitch_tc_kernel(const KSArgs A) {
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
    uint64_t* ready = bars + TC_STAGES;          // the four producer warps have written the step's one-hot rows
    uint64_t* freeb = bars + 2 * TC_STAGES;      // the step's MMAs have completed: both ring slots may be overwritten
    uint64_t* done = bars + 3 * TC_STAGES;       // all MMAs have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = A.rows_in * A.t;
    const int nsteps = (nblk + Q - 1) / Q;
    const unsigned char* kstream = reinterpret_cast<const unsigned char*>(A.key) + (size_t)blockIdx.z * A.key_z_stride * sizeof(int32_t) +
                                   (size_t)blockIdx.y * nsteps * TC_B_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full + s, 1); mbar_init(ready + s, 4); mbar_init(freeb + s, 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 5) {
        // ---- key stream
        if (lane == 0) {
            for (int st = 0; st < nsteps; st++) {
                const int slot = st % TC_STAGES, use = st / TC_STAGES;
                if (use > 0) tc_wait(freeb + slot, (uint32_t)(use - 1) & 1u);
                mbar_expect_tx(full + slot, TC_B_BYTES);
                tma_load_1d(ringB + (size_t)slot * TC_B_BYTES, kstream + (size_t)st * TC_B_BYTES, TC_B_BYTES, full + slot);
            }
        }
    } else if (warp == 4) {
        // ---- MMA issue
        if (lane == 0) {
            for (int st = 0; st < nsteps; st++) {
                const int slot = st % TC_STAGES; const uint32_t ph = (uint32_t)(st / TC_STAGES) & 1u;
                tc_wait(full + slot, ph);
                tc_wait(ready + slot, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = tc_desc(smem_u32(ringA + (size_t)slot * TC_A_BYTES));
                const uint32_t sb = smem_u32(ringB + (size_t)slot * TC_B_BYTES);
#pragma unroll
                for (int p = 0; p < 4; p++) tc_mma_i8(tmem + 128u * p, da, tc_desc(sb + 4096u * p), st > 0 ? 1u : 0u);
                tc_commit(freeb + slot);
            }
            tc_commit(done);
        }
    } else {
        // ---- one-hot rows: thread = sample
        const int m = threadIdx.x;
        const long smp = (long)blockIdx.x * 128 + m;
        const bool live = smp < A.count;
        const TorusIn* in = reinterpret_cast<const TorusIn*>(A.in) + (size_t)(live ? smp : 0) * A.in_stride;
        const U prec_offset = (U)1 << (W - (1 + BASEBIT * A.t));      // cb/lwe_functions.cpp:141 ; poc:444,674
        const uint32_t row_off = (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u;
        // digits of coefficient i live in the top 32 bits of a_i + prec_offset; the value for the NEXT coefficient is requested one
        // coefficient ahead so its latency hides behind a whole coefficient's worth of steps
        int i_cur = 0, j_cur = 0;
        uint32_t a_cur = live ? (uint32_t)(((U)in[0] + prec_offset) >> (W - 32)) : 0u;
        U raw_next = (live && A.rows_in > 1) ? (U)in[1] : (U)0;
        for (int st = 0; st < nsteps; st++) {
            const int slot = st % TC_STAGES, use = st / TC_STAGES;
            uint32_t w[8];
#pragma unroll
            for (int x = 0; x < 8; x++) w[x] = 0u;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                uint32_t d = (a_cur >> (32 - (j_cur + 1) * BASEBIT)) & (uint32_t)(BASE - 1);
                if (i_cur >= A.rows_in) d = 0u;                     // tail of the last step
                // one-hot byte d-1 of this block's group of BASE bytes (nothing for d = 0)
                if constexpr (BASE == 4) w[q] = (1u << (8 * d)) >> 8;
                else if constexpr (BASE == 8) {
                    const uint64_t v = d ? (uint64_t)1 << (8 * (d - 1)) : (uint64_t)0;
                    w[2 * q] = (uint32_t)v; w[2 * q + 1] = (uint32_t)(v >> 32);
                } else {                                            // BASE == 2: two bytes per block, candidate byte first
                    w[q >> 1] |= d << (16 * (q & 1));
                }
                if (++j_cur == A.t) {                               // next coefficient
                    j_cur = 0; i_cur++;
                    a_cur = live ? (uint32_t)((raw_next + prec_offset) >> (W - 32)) : 0u;
                    if (live && i_cur + 1 < A.rows_in) raw_next = (U)in[i_cur + 1];
                }
            }
            if (use > 0) tc_wait(freeb + slot, (uint32_t)(use - 1) & 1u);
            unsigned char* arow = ringA + (size_t)slot * TC_A_BYTES + row_off;
            *reinterpret_cast<uint4*>(arow) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(arow + 2048) = make_uint4(w[4], w[5], w[6], w[7]);
            fence_proxy_async_smem();                               // generic-proxy stores -> visible to the tensor core's (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + slot);
        }
        // ---- epilogue: D[sample][column] of byte plane p sits in lane = sample, column 128 p + column
        tc_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tl = tmem + (((uint32_t)warp * 32u) << 16);
        int32_t* orow = nullptr;
        if (live) orow = A.out + (size_t)blockIdx.z * A.out_z_stride + (size_t)(smp / A.group) * A.out_stride + (size_t)(smp % A.group) * A.out_inner;
        const int colbase = blockIdx.y * 128;
#pragma unroll 1
        for (int c = 0; c < 128; c += 16) {
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
}

template <typename TorusIn, int BASEBIT>
static cudaError_t launch_ks_tc_b(const KSArgs& a, cudaStream_t s) {
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_tc_kernel<TorusIn, BASEBIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
        if (e != cudaSuccess) return e;
        attr_done.done();
    }
    dim3 grid((a.count + 127) / 128, a.cols_pad / 128, a.nz > 0 ? a.nz : 1);
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
        const size_t off = (size_t)(k >> 4) * 2048 + (size_t)(n >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 15);
        unsigned char* img = dst + (cg * nsteps + step) * (size_t)TC_B_BYTES + off;
        img[0] = (unsigned char)v; img[4096] = (unsigned char)(v >> 8); img[8192] = (unsigned char)(v >> 16); img[12288] = (unsigned char)(v >> 24);
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
