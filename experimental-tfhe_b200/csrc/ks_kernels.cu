// ks_kernels.cu -- batched key switching as a tiled gather-accumulate.
//
// Replaces lweKeySwitch / lweKeySwitchTranslate_fromArray (cb/lwe_functions.cpp:136-171), preKeySwitch
// (cb/poc_CircuitBootstrapping.cpp:437-465) and circuitPrivKS (:667-698):
//     result = (0,b) - sum_{i<rows, j<t} key[i][j][ d_ij ],   d_ij = ((a_i + prec_offset) >> (W-(j+1)basebit)) & (base-1), d_ij != 0
// The reference walks the table once per sample (12.3 MB of rows per gate, 147 MB per private key switch).
// Here a CTA owns a tile of KS_TILE samples x 512 output columns: for every (i,j) it fetches the base-1
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
constexpr int KS_HALF = 16;     // samples per thread
constexpr int KS_ICHUNK = 32;   // input coefficients staged per shared-memory refill

template <typename TorusIn, int BASEBIT, int VARIANT>
__global__ void __launch_bounds__(256, 2) keyswitch_kernel(const KSArgs A) {
    typedef typename std::conditional<sizeof(TorusIn) == 4, uint32_t, uint64_t>::type U;
    constexpr int W = sizeof(TorusIn) * 8;
    constexpr int BASE = 1 << BASEBIT;
    __shared__ U abar[KS_TILE][KS_ICHUNK + 1];

    const int cg = threadIdx.x & 127;           // column group: 4 consecutive int32 columns
    const int half = threadIdx.x >> 7;          // which 16 samples of the tile
    const int s0 = blockIdx.x * KS_TILE;        // first sample of the tile
    const int col0 = blockIdx.y * 512 + cg * 4;
    const TorusIn* in = reinterpret_cast<const TorusIn*>(A.in);
    const U prec_offset = (U)1 << (W - (1 + BASEBIT * A.t));      // cb/lwe_functions.cpp:141 ; poc:444,674

    int4 acc[KS_HALF];
#pragma unroll
    for (int s = 0; s < KS_HALF; s++) acc[s] = make_int4(0, 0, 0, 0);

    const size_t row_stride = (size_t)A.cols_pad;                 // one key row
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
            U a[KS_HALF];
#pragma unroll
            for (int s = 0; s < KS_HALF; s++) a[s] = abar[half * KS_HALF + s][ii];
            const int4* krow = key4 + ((size_t)(i0 + ii) * A.t) * (BASE - 1) * (row_stride / 4);
            for (int j = 0; j < A.t; j++) {
                const int sh = W - (j + 1) * BASEBIT;
                if (VARIANT == 0) {
                    // candidate rows in registers, every sample picks one (predicated subtracts)
                    int4 r[BASE - 1];
#pragma unroll
                    for (int d = 0; d < BASE - 1; d++) r[d] = __ldg(krow + ((size_t)j * (BASE - 1) + d) * (row_stride / 4));
#pragma unroll
                    for (int s = 0; s < KS_HALF; s++) {
                        const int dg = (int)((a[s] >> sh) & (U)(BASE - 1));
#pragma unroll
                        for (int d = 0; d < BASE - 1; d++) {
                            if (dg == d + 1) {
                                acc[s].x -= r[d].x; acc[s].y -= r[d].y; acc[s].z -= r[d].z; acc[s].w -= r[d].w;
                            }
                        }
                    }
                } else {
                    // every sample loads the row its digit selects (the base-1 rows of (i,j) stay hot in L1); digit 0 reads
                    // row 0 and is masked out, so there is no branch and no select
                    const int4* kj = krow + (size_t)j * (BASE - 1) * (row_stride / 4);
#pragma unroll
                    for (int s = 0; s < KS_HALF; s++) {
                        const int dg = (int)((a[s] >> sh) & (U)(BASE - 1));
                        const int keep = dg != 0 ? -1 : 0;
                        const int4 r = __ldg(kj + (size_t)max(dg - 1, 0) * (row_stride / 4));
                        acc[s].x -= r.x & keep; acc[s].y -= r.y & keep; acc[s].z -= r.z & keep; acc[s].w -= r.w & keep;
                    }
                }
            }
        }
    }
    // result starts as the noiseless trivial sample (0,b) (cb/lwe_functions.cpp:169) or 0 (poc:677-681)
#pragma unroll
    for (int s = 0; s < KS_HALF; s++) {
        const int smp = s0 + half * KS_HALF + s;
        if (smp >= A.count) continue;
        int v[4] = {acc[s].x, acc[s].y, acc[s].z, acc[s].w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
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
    static const int variant = getenv("TFHE_B200_KS_VARIANT") ? atoi(getenv("TFHE_B200_KS_VARIANT")) : 0;   // development knob
    if (variant == 0) {
        switch (a.basebit) {
            case 1: keyswitch_kernel<TorusIn, 1, 0><<<grid, 256, 0, s>>>(a); break;
            case 2: keyswitch_kernel<TorusIn, 2, 0><<<grid, 256, 0, s>>>(a); break;
            case 3: keyswitch_kernel<TorusIn, 3, 0><<<grid, 256, 0, s>>>(a); break;
            default: return cudaErrorInvalidValue;
        }
    } else {
        switch (a.basebit) {
            case 1: keyswitch_kernel<TorusIn, 1, 1><<<grid, 256, 0, s>>>(a); break;
            case 2: keyswitch_kernel<TorusIn, 2, 1><<<grid, 256, 0, s>>>(a); break;
            case 3: keyswitch_kernel<TorusIn, 3, 1><<<grid, 256, 0, s>>>(a); break;
            default: return cudaErrorInvalidValue;
        }
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
