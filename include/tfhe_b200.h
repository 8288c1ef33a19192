/*
 * tfhe_b200.h -- C ABI of the B200-native TFHE bootstrapping engine.
 *
 * Drop-in boundary for the hot path of tfhe/experimental-tfhe.  Every entry point
 * names the reference function it replaces; "cb/" = circuit-bootstrapping/src/, "hp/" =
 * high-precision-anticyclic-fft/src/ of the reference tree.  The reference processes ONE sample per
 * call on host structs made of separately allocated polynomials (cb/poc_types.h:137-251); this ABI
 * processes a BATCH of samples held in flat, contiguous arrays:
 *
 *   LWE sample, dimension n      : torus[n+1]            a[0..n) then b          (cb/poc_types.h:137-158)
 *   TLWE sample, degree N, k=1   : torus[2][N]           a polynomial then b     (cb/poc_types.h:164-184)
 *   TGSW sample                  : torus[2*l][2][N]      row p = bloc*l + i      (cb/poc_types.h:206-234)
 *   bootstrapping key            : TGSW[n], COEFFICIENT domain (transformed on the device at load)
 *   LWE key-switching key        : int32[N_in][t][base][n_out+1]                 (cb/lwe_functions.cpp:96-110)
 *   private key-switching key    : int32[2][n_in+1][t][base][2][N_out]           (cb/poc_CircuitBootstrapping.cpp:408)
 *   batches                      : sample index is the slowest dimension.
 *
 * Pointers named *_dev are device pointers on the context's GPU; *_host are host pointers (pinned
 * memory gives full PCIe speed but is not required).  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  All *_batch calls on device pointers are asynchronous on
 * `stream`; *_host calls return after the results are in host memory.
 *
 * Streams and threads.  Calls on one context must come from one host thread at a time (the reference is single-threaded and not
 * re-entrant either, cb/spqlios/lagrangehalfc_impl.h:14-19), but they may use DIFFERENT streams: the context owns one set of scratch
 * buffers, and calls that use it (bootstrap_FFT, boots*, bootsMUX, CMux, LUT, CircuitBootstrapFFT, circuit_eval and the *_host
 * forms) order themselves against each other with an event, so a call queued on stream B starts its kernels after the previous
 * call's last kernel on stream A -- results are the same as on a single stream.  Calls that do not touch scratch (blindRotate,
 * bootstrap_woKS, lweKeySwitch, the transforms, hp FFT) run concurrently on their streams.  During CUDA-graph capture the event
 * ordering is off: replay a captured circuit on one stream at a time.  Every entry point selects the context's device itself.
 *
 * Errors: the reference has none (assert/abort).  Here every call returns TFHE_B200_OK or a negative
 * code and tfhe_b200_last_error() gives the message.  There is NO CPU fallback: without a usable
 * sm_100 device tfhe_b200_ctx_create fails.
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TFHE_B200_OK            0
#define TFHE_B200_ERR_CUDA     -1   /* CUDA runtime error (message has the cudaError string) */
#define TFHE_B200_ERR_PARAM    -2   /* unsupported / inconsistent parameter set or argument */
#define TFHE_B200_ERR_NOKEY    -3   /* key material for this call has not been loaded */
#define TFHE_B200_ERR_NODEVICE -4   /* no CUDA device of compute capability 10.x */

typedef struct tfhe_b200_ctx tfhe_b200_ctx;

/* One context per process and GPU; owns key storage and scratch (reference: none -- keys live in
 * `Globals`, cb/poc_types.h:267-312, temporaries are new/delete'd per call, cb/poc_CircuitBootstrapping.cpp:537-546). */
int         tfhe_b200_ctx_create(tfhe_b200_ctx** ctx, int device_ordinal);
int         tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx);
const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx);   /* ctx may be NULL: last create error */
int         tfhe_b200_sm_count(const tfhe_b200_ctx* ctx);
int         tfhe_b200_synchronize(tfhe_b200_ctx* ctx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gate bootstrapping, 32-bit torus (library-style path, cb/lwe_functions.cpp + cb/tgsw_functions.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n;          /* LweParams::n                         cb/lwe_functions.cpp:17   */
    int32_t N;          /* TLweParams::N (1024)                 cb/tlwe_functions.cpp:14  */
    int32_t k;          /* TLweParams::k, must be 1             cb/poc_types.h:10         */
    int32_t bk_l;       /* TGswParams::l                        cb/tgsw_functions.cpp:15  */
    int32_t bk_Bgbit;   /* TGswParams::Bgbit                    cb/tgsw_functions.cpp:15  */
    int32_t ks_t;       /* LweKeySwitchKey::t                   cb/lwe_functions.cpp:96   */
    int32_t ks_basebit; /* LweKeySwitchKey::basebit             cb/lwe_functions.cpp:96   */
} tfhe_b200_gate_params;

/* Replaces init_LweBootstrappingKeyFFT (cb/lwe_functions.cpp:287-316): takes the coefficient-domain
 * bootstrapping key bk[n][2l][2][N] and the key-switching key ks[N][t][base][n+1] from HOST memory,
 * transforms bk on the device (tGswToFFTConvert, cb/tgsw_functions.cpp:389-394) into the engine's
 * private spectral layout and repacks ks on the device: by default into byte-plane images for the tensor-core key switch,
 * with TFHE_B200_KS=cuda in the environment into rows for the CUDA-core kernels (one choice per process, same results). */
int tfhe_b200_gate_load_keys(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p,
                             const int32_t* bk_host, const int32_t* ks_host);
/* Multi-GPU replication: rank 0 loads keys, every other rank allocates, then the two
 * device blobs are broadcast (NCCL / cudaMemcpyPeer) by the caller.  which: 0 = bk spectra, 1 = ks. */
int tfhe_b200_gate_alloc_keys(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p);
int tfhe_b200_gate_key_blob(tfhe_b200_ctx* ctx, int which, void** dev_ptr, size_t* bytes);
/* parameters of the gate keys currently allocated / loaded (the reference reaches them through key->params) */
int tfhe_b200_gate_get_params(const tfhe_b200_ctx* ctx, tfhe_b200_gate_params* p);
/* The allocated buffers are uninitialised: gates refuse to run (TFHE_B200_ERR_NOKEY) until the caller has filled both blobs and
 * calls this. */
int tfhe_b200_gate_commit_keys(tfhe_b200_ctx* ctx);
/* Wire format of the LOADED gate keys (the reference has no serialization and rebuilds its keys on every run,
 * cb/poc_CircuitBootstrapping.cpp:342-423): a 96-byte header -- magic "TFHEB200", format version, the seven parameters, the two
 * blob sizes, an FNV-1a checksum of the payload -- followed by the bootstrapping-key spectra and the repacked key-switching key
 * exactly as they sit in device memory, so importing is two copies and no transform.  The layout is private to a format version:
 * import refuses other versions.  export: call with buf_host = NULL to get the size in *bytes, then with a buffer of that size. */
int tfhe_b200_gate_export_keys(tfhe_b200_ctx* ctx, void* buf_host, size_t* bytes);
int tfhe_b200_gate_import_keys(tfhe_b200_ctx* ctx, const void* buf_host, size_t bytes);

/* tfhe_blindRotate_FFT (cb/lwe_functions.cpp:337-361): accum[B][2][N] in/out, bara[B][n] in [0,2N). */
int tfhe_b200_blindRotate_FFT_batch(tfhe_b200_ctx* ctx, int32_t* accum_dev, const int32_t* bara_dev,
                                    int count, void* stream);
/* tfhe_blindRotateAndExtract_FFT (cb/lwe_functions.cpp:366-395): result[B][N+1]; v[N] one test
 * polynomial shared by the batch; barb[B]; bara[B][n]. */
int tfhe_b200_blindRotateAndExtract_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* v_dev,
                                              const int32_t* barb_dev, const int32_t* bara_dev,
                                              int count, void* stream);
/* tfhe_bootstrap_woKS_FFT (cb/lwe_functions.cpp:399-430): result[B][N+1], x[B][n+1]. */
int tfhe_b200_bootstrap_woKS_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu,
                                       const int32_t* x_dev, int count, void* stream);
/* lweKeySwitch (cb/lwe_functions.cpp:163-171): result[B][n+1], sample[B][N+1]. */
int tfhe_b200_lweKeySwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev,
                                 int count, void* stream);
/* tfhe_bootstrap_FFT (cb/lwe_functions.cpp:434-446): result[B][n+1], x[B][n+1]. */
int tfhe_b200_bootstrap_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu,
                                  const int32_t* x_dev, int count, void* stream);

/* boots* gates (upstream tfhe/tfhe boot-gates.cpp -- not in the reference tree; semantics in
 * SURVEY.md Appendix C): tmp = (0,c) + ka*ca + kb*cb, then tfhe_bootstrap_FFT(result, bk, 1/8, tmp). */
enum {
    TFHE_B200_NAND = 0, TFHE_B200_AND, TFHE_B200_OR, TFHE_B200_NOR, TFHE_B200_XOR, TFHE_B200_XNOR,
    TFHE_B200_ANDNY, TFHE_B200_ANDYN, TFHE_B200_ORNY, TFHE_B200_ORYN, TFHE_B200_NUM_GATES
};
int tfhe_b200_bootsGate_batch(tfhe_b200_ctx* ctx, int op, int32_t* result_dev, const int32_t* ca_dev,
                              const int32_t* cb_dev, int count, void* stream);
/* bootsNOT: negation, no bootstrapping. */
int tfhe_b200_bootsNOT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* ca_dev, int count, void* stream);
/* bootsMUX(a,b,c) = a ? b : c : two bootstraps without key switch, one key switch. */
int tfhe_b200_bootsMUX_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* a_dev, const int32_t* b_dev,
                             const int32_t* c_dev, int count, void* stream);
/* Same gate call on HOST buffers: H2D of ca/cb, the gate, D2H of result, all inside the call
 * (this is the end-to-end path a reference user would bind).  Large batches are cut into chunks of whole waves on two private
 * streams so that the copies of one chunk run under the kernels of another; pinned host buffers make the copies asynchronous. */
int tfhe_b200_bootsGate_batch_host(tfhe_b200_ctx* ctx, int op, int32_t* result_host, const int32_t* ca_host,
                                   const int32_t* cb_host, int count);

/* Gate-level circuits (BASELINE configs[2], SURVEY.md 8d "Config 3" / 8f rank 2; the reference has no circuit layer -- a
 * circuit evaluator over it would be a loop of upstream boots* calls, one sample at a time).
 * A WIRE is a batch of `count` LWE samples, wires_dev[wire][count][n+1]: `count` independent instances of the same netlist
 * run side by side.  Gates execute in the order given (the caller supplies a topological order); a run of consecutive gates
 * with the same op whose wires advance by one per gate (out+j, in0+j, in1+j, no gate of the run reading another's output)
 * is merged into ONE batched launch of run*count samples -- e.g. the 32 a_i XOR b_i of an adder.  All launches go to
 * `stream`; after a first (warm-up) call nothing is allocated, so the call can be captured into a CUDA graph.
 * op: TFHE_B200_NAND .. TFHE_B200_ORYN (in0, in1), TFHE_B200_NOT / TFHE_B200_COPY (in0), TFHE_B200_MUX (in0 ? in1 : in2). */
enum { TFHE_B200_NOT = 16, TFHE_B200_COPY = 17, TFHE_B200_MUX = 18 };
typedef struct { int32_t op, out, in0, in1, in2; } tfhe_b200_gate;
int tfhe_b200_circuit_eval_batch(tfhe_b200_ctx* ctx, const tfhe_b200_gate* gates_host, int n_gates,
                                 int32_t* wires_dev, int n_wires, int count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Negacyclic FP64 transforms (cb/spqlios/fft_processor_spqlios.cpp).  The spectral ("LagrangeHalfC")
 * layout is engine-private, exactly as the reference's is private to spqlios: N doubles per
 * polynomial, only meaningful to these functions.  Pointwise products are order-agnostic.
 * N in {1024, 2048}.
 * ---------------------------------------------------------------------------------------------- */
/* IntPolynomial_ifft / execute_reverse_int (:27-67) */
int tfhe_b200_IntPolynomial_ifft_batch(tfhe_b200_ctx* ctx, double* result_dev, const int32_t* poly_dev,
                                       int N, int count, void* stream);
/* TorusPolynomial64_ifft_lvl2 / execute_reverse_torus64 (:166-170) */
int tfhe_b200_TorusPolynomial64_ifft_batch(tfhe_b200_ctx* ctx, double* result_dev, const int64_t* poly_dev,
                                           int N, int count, void* stream);
/* TorusPolynomial_fft / execute_direct_torus32 (:77-103): scales by 2/N, truncates int32(int64(x)). */
int tfhe_b200_TorusPolynomial_fft_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* lagr_dev,
                                        int N, int count, void* stream);
/* TorusPolynomial64_fft_lvl2 / execute_direct_torus64 (:105-156) */
int tfhe_b200_TorusPolynomial64_fft_batch(tfhe_b200_ctx* ctx, int64_t* result_dev, const double* lagr_dev,
                                          int N, int count, void* stream);
/* LagrangeHalfCPolynomialAddMul (cb/spqlios/lagrangehalfc_impl_fma.s:78-135): res += a (.) b */
int tfhe_b200_LagrangeHalfCPolynomialAddMul_batch(tfhe_b200_ctx* ctx, double* res_dev, const double* a_dev,
                                                  const double* b_dev, int N, int count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TRGSW x TRLWE products at N = 1024, Torus32 (what the TRGSW outputs of a circuit bootstrap feed).
 * A TGSW sample is [2l][2][N] int32 (rows p = bloc*l + i, cb/poc_types.h:206-234) -- exactly one [u][w] block of
 * tfhe_b200_CircuitBootstrapFFT_batch's output; its spectral form is [2l][2][N doubles], engine-private layout.
 * ---------------------------------------------------------------------------------------------- */
/* tGswToFFTConvert (cb/tgsw_functions.cpp:389-394): count TGSW samples -> spectra (scaled by 2/N for the product below). */
int tfhe_b200_tGswToFFTConvert_batch(tfhe_b200_ctx* ctx, double* gswfft_dev, const int32_t* gsw_dev, int l,
                                     int count, void* stream);
/* tGswFFTExternMulToTLwe (cb/tgsw_functions.cpp:424-449): accum[b] <- G (x) accum[b], accum [count][2][N] in place.
 * per_sample != 0: G = gswfft[b]; per_sample == 0: one G for the whole batch. */
int tfhe_b200_tGswFFTExternMulToTLwe_batch(tfhe_b200_ctx* ctx, int32_t* accum_dev, const double* gswfft_dev, int per_sample,
                                           int l, int Bgbit, int count, void* stream);
/* CMux(C, d1, d0) = C (x) (d1 - d0) + d0 (the reference's commented stub, cb/poc_CircuitBootstrapping.cpp:877-879):
 * result, d1, d0 are TRLWE batches [count][2][N]; result may alias d1 or d0. */
int tfhe_b200_CMux_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* gswfft_dev, int per_sample,
                         const int32_t* d1_dev, const int32_t* d0_dev, int l, int Bgbit, int count, void* stream);
/* Vertical-packing look-up: for every sample b, result[b] = TRLWE of table[ sum_j bit_j(b) 2^j ], selected by a CMUX tree
 * over its nsel TRGSW selector bits sel[b][j] (spectral form, [count][nsel][2l][2][N doubles]); table: 2^nsel plaintext
 * polynomials [2^nsel][N] int32 shared by the batch (BASELINE configs[3] "feeding a vertical-packing LUT"). */
int tfhe_b200_LUT_vertical_packing_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* selfft_dev, int nsel,
                                         const int32_t* table_dev, int l, int Bgbit, int count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Circuit bootstrapping (cb/poc_CircuitBootstrapping.cpp), LWE32(N1) -> TRGSW32(N1)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_lvl0, N_lvl1, N_lvl2;               /* :71-73 */
    int32_t bgbit_lvl1, ell_lvl1;                 /* :74-75 */
    int32_t bgbit_lvl2, ell_lvl2;                 /* :76-77 */
    int32_t kslength_lvl10, ksbasebit_lvl10;      /* :80-81 */
    int32_t kslength_lvl21, ksbasebit_lvl21;      /* :83-84 */
} tfhe_b200_cb_params;

/* Replaces the cloud-key part of Globals::Globals (:372-419).  preKS[N1][t10][base10][n0+1],
 * bk[n0][2*l2][2][N2] (Torus64, coefficient domain), privKS[2][N2+1][t21][base21][2][N1] or NULL. */
int tfhe_b200_cb_load_keys(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, const int32_t* preKS_host,
                           const int64_t* bk_host, const int32_t* privKS_host);
/* Multi-GPU replication of the circuit-bootstrap keys, same protocol as the gate keys: rank 0 loads, every other rank
 * allocates, the device blobs are broadcast by the caller (which: 0 = bk spectra, 1 = preKS, 2 = privKS), then commit. */
int tfhe_b200_cb_alloc_keys(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, int with_privks);
int tfhe_b200_cb_key_blob(tfhe_b200_ctx* ctx, int which, void** dev_ptr, size_t* bytes);
int tfhe_b200_cb_commit_keys(tfhe_b200_ctx* ctx);
/* Wire format of the LOADED circuit-bootstrap keys (the reference rebuilds them on every run, ~100 s,
 * cb/poc_CircuitBootstrapping.cpp:342-423): a 128-byte header (magic "TFHEB200", format version, kind 2, the eleven parameters,
 * the three blob sizes, FNV-1a checksum) followed by the three device blobs.  Both directions move the 2.35 GB private key-switch
 * key without a full-size temporary.  export: buf_host = NULL returns the size in *bytes. */
int tfhe_b200_cb_export_keys(tfhe_b200_ctx* ctx, void* buf_host, size_t* bytes);
int tfhe_b200_cb_import_keys(tfhe_b200_ctx* ctx, const void* buf_host, size_t bytes);
/* Ciphertext wire format (the reference has none): 64-byte header (magic "TFHEB2CT", version, kind: 1 LWE32,
 * 2 LWE64, 3 TLWE32, 4 TGSW32; up to four dimensions, slowest first, unused = 0; payload size; FNV-1a checksum) followed by the
 * samples in the flat layouts listed at the top of this header.  Host-side only, no context needed (errors: tfhe_b200_last_error(NULL)).
 * pack: buf_host = NULL returns the size.  unpack: samples_host = NULL returns kind, dims and the payload size. */
int tfhe_b200_ciphertext_pack(int kind, const int64_t dims[4], const void* samples_host, void* buf_host, size_t* bytes);
int tfhe_b200_ciphertext_unpack(const void* buf_host, size_t bytes, int* kind, int64_t dims[4], void* samples_host, size_t* sample_bytes);
/* preKeySwitch (:437-465): result[B][n0+1], x[B][N1+1] */
int tfhe_b200_preKeySwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* x_dev, int count, void* stream);
/* preModSwitch (:472-484): result[B][n0+1] in [0, 2*N2) */
int tfhe_b200_preModSwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* x_dev, int count, void* stream);
/* The blind-rotation loop of circuitBootstrapWoKS alone (:580-642, defects D1/D2 of SURVEY Appendix B corrected) on Torus64
 * accumulators: accum[B][2][N2] in/out, bara[B][n0] in [0, 2*N2).  The Torus64 twin of tfhe_b200_blindRotate_FFT_batch. */
int tfhe_b200_blindRotate64_FFT_batch(tfhe_b200_ctx* ctx, int64_t* accum_dev, const int32_t* bara_dev, int count, void* stream);
/* Exact Torus64 path.  The reference's own answers to "the FP64 FFT keeps 53 of ~85 product bits" are its exact
 * `fake FFT' build (:285-316 -> Karatsuba, cb/poc_karatsuba.cpp:135-206) and the 128-bit FFT (hp/code.cpp:391-512).  Here: a
 * number-theoretic transform over p = 2^64 - 2^32 + 1 with the key split into two 32-bit limbs; every external product is
 * bit-identical to the schoolbook product mod 2^64.  load_exact_key takes the same coefficient-domain bk as cb_load_keys (which must
 * have been called) and builds the NTT-domain key (2x the size of the FP64 spectra); blindRotate64_exact is the exact twin of
 * blindRotate64_FFT; cb_set_exact(1) makes circuitBootstrapWoKS / CircuitBootstrapFFT run their blind rotations on this path. */
int tfhe_b200_cb_load_exact_key(tfhe_b200_ctx* ctx, const int64_t* bk_host);
int tfhe_b200_cb_set_exact(tfhe_b200_ctx* ctx, int on);
int tfhe_b200_blindRotate64_exact_batch(tfhe_b200_ctx* ctx, int64_t* accum_dev, const int32_t* bara_dev, int count, void* stream);
/* circuitBootstrapWoKS (:530-659, with the corrections D1-D3 of SURVEY Appendix B):
 * result[B][N2+1] (Torus64), abar[B][n0+1]. */
int tfhe_b200_circuitBootstrapWoKS_batch(tfhe_b200_ctx* ctx, int64_t* result_dev, int64_t mu,
                                         const int32_t* abar_dev, int count, void* stream);
/* circuitPrivKS (:667-698): result[B][2][N1], x[B][N2+1] */
int tfhe_b200_circuitPrivKS_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int u, const int64_t* x_dev,
                                  int count, void* stream);
/* tfhe_CircuitBootstrapFFT (:823-873): result[B][2][l1][2][N1] (= samples[u][w]), sample[B][N1+1] */
int tfhe_b200_CircuitBootstrapFFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev,
                                        int count, void* stream);
int tfhe_b200_CircuitBootstrapFFT_batch_host(tfhe_b200_ctx* ctx, int32_t* result_host, const int32_t* sample_host,
                                             int count);

/* ------------------------------------------------------------------------------------------------
 * Key generation on the device.  The reference builds its cloud keys in Globals::Globals
 * (cb/poc_CircuitBootstrapping.cpp:342-423) on one core in ~100 s; here each key row is one CTA (Philox4x32-10 counters, Box-Muller
 * Gaussians scaled like cb/generic_utils.h:175-189, exact wrap-around a*K for the TLWE rows) and the keys land in the context ready
 * to use.  The binary SECRET keys are drawn on the host and handed back to the caller (client side); the context does not keep them.
 * gate: lwe_key[n], tlwe_key[N]; optional raw copies bk_raw_host[n][2l][2][N], ks_raw_host[N][t][base][n+1] (NULL to skip) for
 * inspection.  cb: key_lvl0[n0], key_lvl1[N1], key_lvl2[N2+1] (last entry -1, :365-367).
 * ---------------------------------------------------------------------------------------------- */
int tfhe_b200_gate_keygen(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p, double bk_stdev, double ks_stdev, uint64_t seed,
                          int32_t* lwe_key_host, int32_t* tlwe_key_host, int32_t* bk_raw_host, int32_t* ks_raw_host);
int tfhe_b200_cb_keygen(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, double bkstdev_lvl2, double ksstdev_lvl10, double ksstdev_lvl21,
                        uint64_t seed, int32_t* key_lvl0_host, int32_t* key_lvl1_host, int32_t* key_lvl2_host, int with_privks);

/* ------------------------------------------------------------------------------------------------
 * High-precision anticyclic FFT, 128-bit fixed point (hp/code.cpp).  Real96 = value * 2^64 in a
 * wrapping 128-bit integer stored as {lo, hi} 64-bit words (little endian, == unsigned __int128).
 * Complex = {re, im}.  N in {2048, 4096}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t re_lo, re_hi, im_lo, im_hi; } tfhe_b200_cplx96;
/* iFFT (hp/code.cpp:391-443): out[B][N/2], in[B][N] */
int tfhe_b200_hp_iFFT_batch(tfhe_b200_ctx* ctx, tfhe_b200_cplx96* out_dev, const int64_t* in_dev,
                            int N, int count, void* stream);
/* FFT (hp/code.cpp:446-512): out[B][N], in[B][N/2] (not clobbered, unlike the reference) */
int tfhe_b200_hp_FFT_batch(tfhe_b200_ctx* ctx, int64_t* out_dev, const tfhe_b200_cplx96* in_dev,
                           int N, int count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Engine diagnostics (no reference counterpart; the reference's only instrumentation is clock() around loops,
 * cb/poc_CircuitBootstrapping.cpp:1008-1016).  Used by bench.py for the roofline numbers.
 * ---------------------------------------------------------------------------------------------- */
/* When enabled, every kernel launch of the batched entry points is bracketed by CUDA events on its stream.
 * categories: 0 = blind rotation, 1 = key switching, 2 = everything else. */
int tfhe_b200_profile_enable(tfhe_b200_ctx* ctx, int on);
/* Synchronises the recorded events, returns summed milliseconds and launch counts per category, then resets. */
int tfhe_b200_profile_read(tfhe_b200_ctx* ctx, double ms[3], int launches[3]);
/* Measured FP64 FMA throughput of this GPU (dependent-chain-free DFMA kernel), in TFLOP/s (2 flop per FMA). */
int tfhe_b200_probe_fp64_tflops(tfhe_b200_ctx* ctx, double* tflops);
/* Measured read bandwidth over a `bytes`-sized buffer that is re-read `passes` times (L2-resident when small), GB/s. */
/* sustained rate (10^9 per second) of 128-bit fixed-point products (hp/code.cpp:148-169 intmul_best) with operands in registers:
 * the arithmetic roofline of the high-precision FFT, measured next to it */
int tfhe_b200_probe_real96_gprods(tfhe_b200_ctx* ctx, double* gprods);
int tfhe_b200_probe_read_gbs(tfhe_b200_ctx* ctx, size_t bytes, int passes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_B200_H */
