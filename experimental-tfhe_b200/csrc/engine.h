// engine.h -- internal C++ interface between the C-ABI layer (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/tfhe_b200.h"

namespace tfhe_b200 {

typedef double2 cplx;

// cudaFuncSetAttribute is per DEVICE: "done" flags for the opt-in shared-memory sizes are kept per device ordinal, so a process
// that opens contexts on several GPUs configures each of them.
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need() const { int d = 0; cudaGetDevice(&d); return !((mask >> (d & 63)) & 1ull); }
    void done() { int d = 0; cudaGetDevice(&d); mask |= 1ull << (d & 63); }
};

// host-side twiddle generation (twiddles.cpp): TreePlan<LOGM> layout (tree_fft.cuh), entries are (re,im) pairs
void make_fft_tables(int logM, double* out /* 2 * TW_TOTAL doubles */);
int  fft_table_entries(int logM);
// hp tables: 2N entries of {re_lo,re_hi,im_lo,im_hi}; inverse=0 -> powomega, 1 -> powombar (hp/code.cpp:378-388)
void make_hp_tables(int n2N, int inverse, uint64_t* out /* 4 * n2N words */);

// ------------------------------------------------------------------ blind rotation (br_kernels.cu)
enum BRMode { BR_ACCUM = 0, BR_TESTVEC = 1, BR_LWE = 2, BR_EXTMUL = 3 };
struct BRArgs {
    const cplx* bkfft;    // [n][2l][2][M] spectra, pre-scaled by 2/N
    const cplx* tw;       // TreePlan table
    int n, l, Bgbit, count, mode;
    // BR_ACCUM  : accum[B][2][N] in/out, bara[B][n]
    // BR_TESTVEC: v[N], barb[B], bara[B][n] -> out[B][N+1]
    // BR_LWE    : x = (0,cconst) + ka*xa + kb*xb (LWE(n) samples, xb may be null), test vector = mu -> out[B][N+1]
    // BR_EXTMUL : accum[B][2][N] <- G_b (x) accum_b, one external product per sample, G_b = bkfft + b * bk_sample_stride
    //             (tGswFFTExternMulToTLwe, cb/tgsw_functions.cpp:424-449); n, bara unused
    void* accum;
    const int32_t* bara;
    const int32_t* barb;
    const void* v;
    const int32_t* xa;
    const int32_t* xb;
    int ka, kb;
    int32_t cconst;
    int64_t mu;           // Torus32 path uses the low 32 bits
    void* out;
    int out_stride;       // elements between consecutive outputs (N+1 by default)
    // Torus64 (circuit bootstrap) only: number of test vectors sharing one pass (mu_w = 2^(64-(w+1)*bgbit1))
    int n_mu; int mu_bgbit;
    size_t bk_sample_stride;   // BR_EXTMUL: cplx elements between the TGSW spectra of consecutive samples (0: one TGSW for all)
    int units_per_gsw;         // BR_EXTMUL: consecutive accumulators sharing one TGSW (nodes of a LUT level); 0 = 1
};
cudaError_t launch_blind_rotate32(const BRArgs& a, cudaStream_t s);     // N = 1024, Torus32
cudaError_t launch_blind_rotate64(const BRArgs& a, cudaStream_t s);     // N = 2048, Torus64 (circuitBootstrapWoKS)
cudaError_t blind_rotate_init();                                         // opt-in shared memory sizes
cudaError_t launch_extern_mul32(const BRArgs& a, cudaStream_t s);        // BR_EXTMUL, N = 1024, Torus32

// coefficient polynomials -> spectra in engine order; scale applied to the output
cudaError_t launch_poly_to_spectrum32(cplx* out, const int32_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s);
cudaError_t launch_poly_to_spectrum64(cplx* out, const int64_t* in, const cplx* tw, int N, int count, double scale, cudaStream_t s);
cudaError_t launch_spectrum_to_torus32(int32_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s);
cudaError_t launch_spectrum_to_torus64(int64_t* out, const cplx* in, const cplx* tw, int N, int count, double scale, cudaStream_t s);
cudaError_t launch_spectrum_addmul(cplx* res, const cplx* a, const cplx* b, size_t n_cplx, cudaStream_t s);

// ------------------------------------------------------------------ exact Torus64 blind rotation (exact_kernels.cu, exact_ntt.cuh)
struct ExactArgs {
    const uint64_t* key;            // [n][2l][2 q][2 limbs][N] NTT domain (Goldilocks), bit-reversed slots
    const uint64_t* psi_rev;        // [N]
    const uint64_t* psi_inv_rev;    // [N]
    uint64_t n_inv;
    int n, l, Bgbit, count, mode;   // mode: BR_ACCUM (accum[B][2][N], bara[B][n]) or BR_LWE (bara[B][n+1] -> out[B*n_mu][out_stride])
    int64_t* accum;
    const int32_t* bara;
    int64_t mu;
    int64_t* out;
    int out_stride;
    int n_mu, mu_bgbit;
};
cudaError_t launch_exact_key(uint64_t* out, const int64_t* in, const uint64_t* psi_rev, int N, size_t npoly, cudaStream_t s);
cudaError_t launch_exact_blind_rotate(const ExactArgs& a, cudaStream_t s);

// ------------------------------------------------------------------ key switching (ks_kernels.cu)
// Device key layout: int32 [rows_in][t][base-1][cols_pad], cols_pad multiple of 512 (d = 0 rows dropped).
struct KSArgs {
    const void* in;        // [B][in_stride] torus (32 or 64 bit)
    int in_stride;         // elements per input sample
    int rows_in;           // number of input coefficients consumed (N, or N2+1 for the private KS)
    int t, basebit;
    const int32_t* key;    // device layout above
    int cols, cols_pad;    // output width (n+1, or 2*N1)
    int b_col;             // >= 0: out[b_col] starts at (int32) in[b_index]; < 0: starts at 0
    int b_index;
    int32_t* out;          // [B][out_stride]
    int out_stride;
    int count;
    // optional (0 = plain): sample s writes to out + (s / group) * out_stride + (s % group) * out_inner ; grid.z = nz
    // independent keys / outputs (key + z * key_z_stride, out + z * out_z_stride) in one launch
    int group, out_inner;
    int nz; size_t key_z_stride, out_z_stride;
};
// Two packings of a key-switching key, chosen once per process (TFHE_B200_KS=cuda selects the first; default is the second):
//   KS_PACK_ROWS: int32 rows for the CUDA-core kernels (ks_kernels.cu): [cols_pad/512][rows][t][base-1][512]
//   KS_PACK_TC  : byte-plane images for the tensor-core kernel (ks_tc_kernels.cu): [cols_pad/128][step][plane][4096 B]
// launch_ks_repack* build the active packing, launch_keyswitch* run the matching kernel, ks_key_bytes sizes the buffer.
enum KSPacking { KS_PACK_ROWS = 0, KS_PACK_TC = 1 };
int ks_packing();
size_t ks_key_bytes(int rows, int t, int basebit, int cols_pad);
cudaError_t launch_keyswitch32(const KSArgs& a, cudaStream_t s);
cudaError_t launch_keyswitch64(const KSArgs& a, cudaStream_t s);
// raw [rows][t][base][cols] -> device layout
cudaError_t launch_ks_repack(int32_t* dst, const int32_t* src, int rows, int t, int base, int cols, int cols_pad, cudaStream_t s);
// paired form of a base-4 key (two digits -> one base-16 digit whose row is the sum of the two rows; ks_kernels.cu), t even
cudaError_t launch_ks_repack_pair(int32_t* dst, const int32_t* src, int rows, int t, int cols, int cols_pad, cudaStream_t s);
// the same for a slice [row0, row0 + rows) of rows_total input rows; src holds the slice only
cudaError_t launch_ks_repack_rows(int32_t* dst, const int32_t* src, int rows_total, int row0, int rows, int t, int base, int cols, int cols_pad,
                                  cudaStream_t s);

// ------------------------------------------------------------------ key generation on the device (keygen_kernels.cu)
// raw host layouts of include/tfhe_b200.h; kidx = indices of the set bits of the binary TLWE key, s = the key being encrypted
cudaError_t launch_lwe_ks_keygen(int32_t* out, const int32_t* s_in, const int32_t* s_out, int rows_in, int n_out, int t, int basebit, double stdev,
                                 uint64_t seed, uint32_t stream, cudaStream_t st);
cudaError_t launch_tlwe_gadget_keygen32(int32_t* out, const int32_t* kidx, int kweight, const int32_t* s, int n, int l, int Bgbit, double stdev,
                                        uint64_t seed, uint32_t stream, cudaStream_t st);
cudaError_t launch_tlwe_gadget_keygen64(int64_t* out, const int32_t* kidx, int kweight, const int32_t* s, int n, int l, int Bgbit, double stdev,
                                        uint64_t seed, uint32_t stream, cudaStream_t st);
cudaError_t launch_tlwe_privks_keygen(int32_t* out, const int32_t* kidx, int kweight, const int32_t* s, int rows_i, int t, int basebit, double stdev,
                                      uint64_t seed, uint32_t stream, size_t row0, size_t nrows, cudaStream_t st);

// ------------------------------------------------------------------ small elementwise (misc_kernels.cu)
cudaError_t launch_lwe_lincomb(int32_t* out, const int32_t* a, const int32_t* b, int ka, int kb, int32_t cconst,
                               int n, int count, cudaStream_t s);   // out = (0,cconst) + ka*a + kb*b   (b may be null)
cudaError_t probe_fp64(double* tflops);
cudaError_t probe_read(size_t bytes, int passes, double* gbs);
// TRLWE pair steps of a CMUX tree over flat units of `len` int32: mode 0: out[u] = in[2u+1] - in[2u]; mode 1: out[u] += in[2u]
cudaError_t launch_pair_combine(int32_t* out, const int32_t* in, int len, size_t units, int mode, cudaStream_t s);
// LUT level 0 from a plaintext table: mode 0: out[c][i] = trivial TRLWE (0, table[2i+1] - table[2i]); mode 1: out[c][i].b += table[2i]
cudaError_t launch_lut_table(int32_t* out, const int32_t* table, int N, int pairs, int count, int mode, cudaStream_t s);
cudaError_t launch_modswitch(int32_t* out, const int32_t* in, int Msize_log2, size_t total, cudaStream_t s);

// ------------------------------------------------------------------ high-precision FFT (hp_kernels.cu)
cudaError_t launch_hp_ifft(tfhe_b200_cplx96* out, const int64_t* in, const uint64_t* powomega, int N, int count, cudaStream_t s);
cudaError_t launch_hp_fft(int64_t* out, const tfhe_b200_cplx96* in, const uint64_t* powombar, int N, int count, cudaStream_t s);
cudaError_t hp_init();
cudaError_t probe_real96(double* gprod_per_s);   // sustained rate of real96 products (10^9 / s): the hp FFT's arithmetic roofline

}  // namespace tfhe_b200
