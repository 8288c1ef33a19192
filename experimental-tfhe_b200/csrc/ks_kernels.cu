// ks_kernels.cu -- batched key switching as a tiled gather-accumulate.
//
// Replaces lweKeySwitch / lweKeySwitchTranslate_fromArray (cb/lwe_functions.cpp:136-171), preKeySwitch
// (cb/poc_CircuitBootstrapping.cpp:437-465) and circuitPrivKS (:667-698):
//     result = (0,b) - sum_{i<rows, j<t} key[i][j][ d_ij ],   d_ij = ((a_i + prec_offset) >> (W-(j+1)basebit)) & (base-1), d_ij != 0
// The reference walks the table once per sample (12.3 MB of rows per gate, 147 MB per private key switch).
// Here a CTA owns a tile of KS_TILE (32) samples x 512 output columns: for every (i,j) it fetches the base-1
// candidate rows ONCE (coalesced 16-byte loads) and each sample of the tile subtracts the row its digit
// selects, so table traffic per sample drops by ~KS_TILE*(base-1)/base / (base-1).  Digits are uniform across
// a warp (all lanes of a warp work on the same samples), so the selection is a uniform branch.
//
// Device key layout: int32 [rows][t][base-1][cols_pad]  (d = 0 rows are never read by the reference either).
#include "engine.h"
#include <type_traits>
#include <cstdlib>

namespace tfhe_b200 {

constexpr int KS_TILE = 32;     // samples per CTA
constexpr int KS_S = 8;         // samples per thread
constexpr int KS_ICHUNK = 32;   // input coefficients staged per shared-memory refill

// Thread layout: 64 column groups (8 consecutive int32 columns = two int4 each) x 4 sample groups (8 samples each).
// A warp is 32 column groups of ONE sample group, so a sample's digit is warp-uniform and picking the row is a uniform branch.
// 8 columns per thread (instead of 4) halves the per-add overhead of digit extraction and branching (profiles/r1_notes.md).
template <typename TorusIn, int BASEBIT>
__global__ void __launch_bounds__(256, 2) keyswitch_kernel(const KSArgs A) {
    typedef typename std::conditional<sizeof(TorusIn) == 4, uint32_t, uint64_t>::type U;
    constexpr int W = sizeof(TorusIn) * 8;
    constexpr int BASE = 1 << BASEBIT;
    __shared__ U abar[KS_TILE][KS_ICHUNK + 1];

    const int cg = threadIdx.x & 63;            // column group: columns [8 cg, 8 cg + 8) of this CTA's 512-column slice
    const int sg = threadIdx.x >> 6;            // sample group
    const int s0 = blockIdx.x * KS_TILE;        // first sample of the tile
    const int col0 = blockIdx.y * 512 + cg * 8;
    const TorusIn* in = reinterpret_cast<const TorusIn*>(A.in);
    const U prec_offset = (U)1 << (W - (1 + BASEBIT * A.t));      // cb/lwe_functions.cpp:141 ; poc:444,674

    int4 acc0[KS_S], acc1[KS_S];
#pragma unroll
    for (int s = 0; s < KS_S; s++) { acc0[s] = make_int4(0, 0, 0, 0); acc1[s] = make_int4(0, 0, 0, 0); }

    const size_t rs4 = (size_t)A.cols_pad / 4;                    // one key row, in int4
    const int4* key4 = reinterpret_cast<const int4*>(A.key + col0);

    for (int i0 = 0; i0 < A.rows_in; i0 += KS_ICHUNK) {
        __syncthreads();
        for (int e = threadIdx.x; e < KS_ICHUNK * KS_TILE; e += 256) {
            const int ii = e % KS_ICHUNK, s = e / KS_ICHUNK;      // consecutive threads read consecutive coefficients
            const int smp = s0 + s, i = i0 + ii;
            U v = 0;
            if (smp < A.count && i < A.rows_in) v = (U)in[(size_t)smp * A.in_stride + i] + prec_offset;
            abar[s][ii] = v;
        }
        __syncthreads();
        const int iend = min(KS_ICHUNK, A.rows_in - i0);
        for (int ii = 0; ii < iend; ii++) {
            U a[KS_S];
#pragma unroll
            for (int s = 0; s < KS_S; s++) a[s] = abar[sg * KS_S + s][ii];
            const int4* krow = key4 + ((size_t)(i0 + ii) * A.t) * (BASE - 1) * rs4;
            for (int j = 0; j < A.t; j++) {
                const int sh = W - (j + 1) * BASEBIT;
                int4 r0[BASE - 1], r1[BASE - 1];
#pragma unroll
                for (int d = 0; d < BASE - 1; d++) {
                    const int4* rp = krow + ((size_t)j * (BASE - 1) + d) * rs4;
                    r0[d] = __ldg(rp); r1[d] = __ldg(rp + 1);
                }
#pragma unroll
                for (int s = 0; s < KS_S; s++) {
                    const int dg = (int)((a[s] >> sh) & (U)(BASE - 1));
#pragma unroll
                    for (int d = 0; d < BASE - 1; d++) {
                        if (dg == d + 1) {
                            acc0[s].x -= r0[d].x; acc0[s].y -= r0[d].y; acc0[s].z -= r0[d].z; acc0[s].w -= r0[d].w;
                            acc1[s].x -= r1[d].x; acc1[s].y -= r1[d].y; acc1[s].z -= r1[d].z; acc1[s].w -= r1[d].w;
                        }
                    }
                }
            }
        }
    }
    // result starts as the noiseless trivial sample (0,b) (cb/lwe_functions.cpp:169) or 0 (poc:677-681)
#pragma unroll
    for (int s = 0; s < KS_S; s++) {
        const int smp = s0 + sg * KS_S + s;
        if (smp >= A.count) continue;
        const int v[8] = {acc0[s].x, acc0[s].y, acc0[s].z, acc0[s].w, acc1[s].x, acc1[s].y, acc1[s].z, acc1[s].w};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int col = col0 + c;
            if (col < A.cols) {
                int32_t x = v[c];
                if (col == A.b_col) x += (int32_t)in[(size_t)smp * A.in_stride + A.b_index];
                A.out[(size_t)smp * A.out_stride + col] = x;
            }
        }
    }
}

template <typename TorusIn>
static cudaError_t launch_ks(const KSArgs& a, cudaStream_t s) {
    if (a.count <= 0) return cudaSuccess;
    if (a.cols_pad % 512) return cudaErrorInvalidValue;
    dim3 grid((a.count + KS_TILE - 1) / KS_TILE, a.cols_pad / 512);
    switch (a.basebit) {
        case 1: keyswitch_kernel<TorusIn, 1><<<grid, 256, 0, s>>>(a); break;
        case 2: keyswitch_kernel<TorusIn, 2><<<grid, 256, 0, s>>>(a); break;
        case 3: keyswitch_kernel<TorusIn, 3><<<grid, 256, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_keyswitch32(const KSArgs& a, cudaStream_t s) { return launch_ks<int32_t>(a, s); }
cudaError_t launch_keyswitch64(const KSArgs& a, cudaStream_t s) { return launch_ks<int64_t>(a, s); }

// raw [rows][t][base][cols] -> [rows][t][base-1][cols_pad]  (zero padded)
__global__ void ks_repack_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, size_t nrows_out, int base, int cols, int cols_pad) {
    const size_t total = nrows_out * (size_t)cols_pad;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t ro = e / cols_pad; const int c = (int)(e % cols_pad);
        const size_t ij = ro / (base - 1); const int d = (int)(ro % (base - 1)) + 1;
        dst[e] = c < cols ? src[(ij * base + d) * (size_t)cols + c] : 0;
    }
}
cudaError_t launch_ks_repack(int32_t* dst, const int32_t* src, int rows, int t, int base, int cols, int cols_pad, cudaStream_t s) {
    const size_t nrows_out = (size_t)rows * t * (base - 1);
    ks_repack_kernel<<<148 * 8, 256, 0, s>>>(dst, src, nrows_out, base, cols, cols_pad);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
