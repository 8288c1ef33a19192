// capi.cu -- the C ABI of include/tfhe_b200.h: context, key ingestion, batched entry points.
// No CPU fallback anywhere in this file: every entry point launches sm_100a kernels or fails.
#include "engine.h"
#include "exact_ntt.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>

using namespace tfhe_b200;

struct tfhe_b200_ctx {
    int device = -1;
    int sm_count = 0;
    std::string err;
    // FFT tables per ring degree
    cplx* tw1024 = nullptr;
    cplx* tw2048 = nullptr;
    // gate keys
    bool gate_ready = false;
    tfhe_b200_gate_params gp{};
    cplx* g_bkfft = nullptr;  size_t g_bkfft_bytes = 0;
    int32_t* g_ks = nullptr;  size_t g_ks_bytes = 0;
    // circuit-bootstrap keys
    bool cb_ready = false;
    tfhe_b200_cb_params cp{};
    cplx* c_bkfft = nullptr;
    int32_t* c_preks = nullptr;
    // exact (NTT) form of the Torus64 bootstrapping key, optional (tfhe_b200_cb_load_exact_key)
    uint64_t* c_bkntt = nullptr;    // [n0][2 l2][2][2 limbs][N2]
    uint64_t* ntt_tab = nullptr;    // psi_rev[N2] | psi_inv_rev[N2]
    uint64_t ntt_n_inv = 0;
    bool cb_exact = false;          // circuitBootstrapWoKS / CircuitBootstrapFFT use the exact blind rotation
    int32_t* c_privks = nullptr;    // [2][rows][t][base-1][2*N1]
    size_t c_privks_u_stride = 0;   // int32 elements per u
    // hp tables
    uint64_t* hp_omega[2] = {nullptr, nullptr};   // N = 2048, 4096
    uint64_t* hp_ombar[2] = {nullptr, nullptr};
    // scratch (grown on demand, never per call once warm)
    void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    // Scratch is one set per context while every *_batch call is asynchronous on the CALLER's stream: users of the scratch are
    // ordered against each other through this event (recorded behind the last kernel of each call, waited for by the next call
    // when it arrives on another stream) -- see ScratchUse below.
    cudaEvent_t scratch_ev = nullptr;
    bool scratch_busy = false;
    cudaStream_t scratch_stream = nullptr;
    // two private streams for the chunked host-buffer path (copies of one chunk overlap the kernels of another)
    cudaStream_t hs[2] = {nullptr, nullptr};
    // profiling (diagnostics): events around every launch when enabled
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int cat; };
    std::vector<Span> spans;
};

// brackets one kernel launch with events on its stream when profiling is on
struct ProfScope {
    tfhe_b200_ctx* c; cudaStream_t s; int idx = -1;
    ProfScope(tfhe_b200_ctx* ctx, int cat, cudaStream_t st) : c(ctx), s(st) {
        if (!c->profiling) return;
        tfhe_b200_ctx::Span sp; sp.cat = cat;
        cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, s);
        c->spans.push_back(sp); idx = (int)c->spans.size() - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(c->spans[idx].b, s); }
};

static std::string g_create_err;
static std::mutex g_mu;

static int fail(tfhe_b200_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else { std::lock_guard<std::mutex> l(g_mu); g_create_err = msg; }
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, TFHE_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define NEED(cond, msg) do { if (!(cond)) return fail(ctx, TFHE_B200_ERR_PARAM, msg); } while (0)

// Brackets the launches of one entry point that uses ctx->scratch.  Two calls on different streams of the same context would
// otherwise race on the scratch buffers (the second call's blind rotation overwriting what the first call's key switch still
// reads); with the bracket the second call's stream waits for the first call's last kernel.  Calls on one stream cost nothing
// extra (stream order already holds).  While `stream` is being captured into a CUDA graph the bracket does nothing: a captured
// sequence is ordered by the capture itself, and the graph must not be replayed concurrently with other calls on this context.
struct ScratchUse {
    tfhe_b200_ctx* c; cudaStream_t s; bool active = false;
    ScratchUse(tfhe_b200_ctx* ctx, cudaStream_t st) : c(ctx), s(st) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
        active = cs == cudaStreamCaptureStatusNone;
        if (!active) return;
        if (!c->scratch_ev) cudaEventCreateWithFlags(&c->scratch_ev, cudaEventDisableTiming);
        if (c->scratch_busy && c->scratch_stream != s) cudaStreamWaitEvent(s, c->scratch_ev, 0);
    }
    ~ScratchUse() {
        if (!active || !c->scratch_ev) return;
        if (cudaEventRecord(c->scratch_ev, s) == cudaSuccess) { c->scratch_busy = true; c->scratch_stream = s; }
    }
};

static int ensure_scratch(tfhe_b200_ctx* ctx, int slot, size_t bytes) {
    if (ctx->scratch_bytes[slot] >= bytes) return TFHE_B200_OK;
    // growing: whatever was queued on the old buffer has to be finished before it is freed (first calls / larger batches only)
    if (ctx->scratch_busy && ctx->scratch_ev) { CU(cudaEventSynchronize(ctx->scratch_ev)); ctx->scratch_busy = false; }
    if (ctx->scratch[slot]) { CU(cudaFree(ctx->scratch[slot])); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
    size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&ctx->scratch[slot], want));
    ctx->scratch_bytes[slot] = want;
    return TFHE_B200_OK;
}

static int upload_tw(tfhe_b200_ctx* ctx, int logM, cplx** dst) {
    const int ne = fft_table_entries(logM);
    std::vector<double> h((size_t)ne * 2);
    make_fft_tables(logM, h.data());
    CU(cudaMalloc(dst, sizeof(cplx) * ne));
    CU(cudaMemcpy(*dst, h.data(), sizeof(cplx) * ne, cudaMemcpyHostToDevice));
    return TFHE_B200_OK;
}

namespace {
struct DevTmp {           // frees on scope exit
    void* p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
uint64_t splitmix64(uint64_t& x) { uint64_t z = (x += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
// uniform binary secret key on the host (client side: it is returned to the caller, never kept)
void binary_key(int32_t* out, int n, uint64_t& state) { for (int i = 0; i < n; i++) out[i] = (int32_t)(splitmix64(state) >> 63); }
int upload_key(tfhe_b200_ctx* ctx, DevTmp& bits, DevTmp& idx, int* weight, const int32_t* key, int n) {
    std::vector<int32_t> set;
    for (int i = 0; i < n; i++) if (key[i]) set.push_back(i);
    *weight = (int)set.size();
    CU(cudaMalloc(&bits.p, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
    CU(cudaMemcpy(bits.p, key, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&idx.p, sizeof(int32_t) * (set.size() + 1)));
    if (!set.empty()) CU(cudaMemcpy(idx.p, set.data(), sizeof(int32_t) * set.size(), cudaMemcpyHostToDevice));
    return TFHE_B200_OK;
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int tfhe_b200_ctx_create(tfhe_b200_ctx** out, int device) {
    tfhe_b200_ctx* ctx = nullptr;   // for the CU macro: errors go to the global slot
    if (!out) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TFHE_B200_ERR_NODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: device ordinal out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, TFHE_B200_ERR_NODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                                        ", this library carries sm_100a code only");
    CU(cudaSetDevice(device));
    tfhe_b200_ctx* c = new tfhe_b200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    ctx = c;
    int rc;
    if ((rc = upload_tw(ctx, 9, &c->tw1024)) != TFHE_B200_OK || (rc = upload_tw(ctx, 10, &c->tw2048)) != TFHE_B200_OK) {
        fail(nullptr, rc, c->err); delete c; return rc;
    }
    e = blind_rotate_init();
    if (e != cudaSuccess) { fail(nullptr, TFHE_B200_ERR_CUDA, std::string("blind_rotate_init: ") + cudaGetErrorString(e)); delete c; return TFHE_B200_ERR_CUDA; }
    e = hp_init();
    if (e != cudaSuccess) { fail(nullptr, TFHE_B200_ERR_CUDA, std::string("hp_init: ") + cudaGetErrorString(e)); delete c; return TFHE_B200_ERR_CUDA; }
    *out = c;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_OK;
    cudaSetDevice(ctx->device);
    for (int k = 0; k < 2; k++) if (ctx->hs[k]) cudaStreamDestroy(ctx->hs[k]);
    if (ctx->scratch_ev) cudaEventDestroy(ctx->scratch_ev);
    cudaFree(ctx->tw1024); cudaFree(ctx->tw2048);
    cudaFree(ctx->g_bkfft); cudaFree(ctx->g_ks);
    cudaFree(ctx->c_bkfft); cudaFree(ctx->c_preks); cudaFree(ctx->c_privks); cudaFree(ctx->c_bkntt); cudaFree(ctx->ntt_tab);
    for (int i = 0; i < 2; i++) { cudaFree(ctx->hp_omega[i]); cudaFree(ctx->hp_ombar[i]); }
    for (int i = 0; i < 4; i++) cudaFree(ctx->scratch[i]);
    delete ctx;
    return TFHE_B200_OK;
}

const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> l(g_mu);
    return g_create_err.c_str();
}
int tfhe_b200_sm_count(const tfhe_b200_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int tfhe_b200_synchronize(tfhe_b200_ctx* ctx, void* stream) {
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ gate keys */
static int check_gate_params(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p) {
    NEED(p, "gate params: null");
    NEED(p->N == 1024, "gate params: only N = 1024 is compiled (TLweParams::N)");
    NEED(p->k == 1, "gate params: k must be 1 (cb/poc_types.h:10)");
    NEED(p->n >= 1 && p->n <= 1024, "gate params: n out of range [1,1024]");
    NEED(p->bk_l >= 1 && p->bk_l * p->bk_Bgbit <= 32 && p->bk_Bgbit >= 1 && p->bk_Bgbit <= 16, "gate params: bad gadget (l, Bgbit)");
    NEED(p->ks_basebit >= 1 && p->ks_basebit <= 3, "gate params: ks_basebit must be 1..3");
    NEED(p->ks_t >= 1 && p->ks_t * p->ks_basebit <= 31, "gate params: ks_t * ks_basebit must be <= 31");
    return TFHE_B200_OK;
}
static size_t gate_cols_pad(const tfhe_b200_gate_params& p) { return (size_t)((p.n + 1 + 511) / 512) * 512; }
// Device packing of the gate key-switching key.  A base-4 key with an even number of digits is stored PAIRED: two digits = one
// base-16 digit whose row is the sum of the two rows (ks_kernels.cu) -- 2.5x the bytes, half the blocks and subtractions.
// TFHE_B200_KS_PAIR=0 keeps the plain packing (development comparison).
static bool gate_ks_paired(const tfhe_b200_gate_params& p) {
    static const char* env = getenv("TFHE_B200_KS_PAIR");
    if (env && env[0] == '0') return false;
    return ks_packing() == KS_PACK_ROWS && p.ks_basebit == 2 && p.ks_t % 2 == 0;
}
static size_t gate_ks_bytes(const tfhe_b200_gate_params& p) {
    if (ks_packing() == KS_PACK_TC) return ks_key_bytes(p.N, p.ks_t, p.ks_basebit, (int)gate_cols_pad(p));
    const size_t blocks = gate_ks_paired(p) ? (size_t)p.N * (p.ks_t / 2) * 15 : (size_t)p.N * p.ks_t * ((1 << p.ks_basebit) - 1);
    return blocks * gate_cols_pad(p) * sizeof(int32_t);
}
static int gate_ks_packing_id(const tfhe_b200_gate_params& p) { return ks_packing() == KS_PACK_TC ? 2 : (gate_ks_paired(p) ? 1 : 0); }
static cudaError_t gate_ks_repack(int32_t* dst, const int32_t* raw_dev, const tfhe_b200_gate_params& p) {
    if (gate_ks_paired(p)) return launch_ks_repack_pair(dst, raw_dev, p.N, p.ks_t, p.n + 1, (int)gate_cols_pad(p), 0);
    return launch_ks_repack(dst, raw_dev, p.N, p.ks_t, 1 << p.ks_basebit, p.n + 1, (int)gate_cols_pad(p), 0);
}

int tfhe_b200_gate_alloc_keys(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    int rc = check_gate_params(ctx, p); if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if (ctx->g_bkfft) { CU(cudaFree(ctx->g_bkfft)); ctx->g_bkfft = nullptr; }
    if (ctx->g_ks) { CU(cudaFree(ctx->g_ks)); ctx->g_ks = nullptr; }
    ctx->gate_ready = false;
    ctx->gp = *p;
    const size_t npoly = (size_t)p->n * 2 * p->bk_l * 2;
    ctx->g_bkfft_bytes = npoly * (p->N / 2) * sizeof(cplx);
    ctx->g_ks_bytes = gate_ks_bytes(*p);
    CU(cudaMalloc(&ctx->g_bkfft, ctx->g_bkfft_bytes));
    CU(cudaMalloc(&ctx->g_ks, ctx->g_ks_bytes));
    // NOT ready yet: the buffers are uninitialised until the caller has filled them (broadcast receiver) and says so with
    // tfhe_b200_gate_commit_keys
    return TFHE_B200_OK;
}

int tfhe_b200_gate_commit_keys(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (!ctx->g_bkfft || !ctx->g_ks) return fail(ctx, TFHE_B200_ERR_NOKEY, "gate_commit_keys: gate keys not allocated");
    ctx->gate_ready = true;
    return TFHE_B200_OK;
}

int tfhe_b200_gate_load_keys(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p, const int32_t* bk_host, const int32_t* ks_host) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(bk_host && ks_host, "gate_load_keys: null key pointer");
    int rc = tfhe_b200_gate_alloc_keys(ctx, p); if (rc) return rc;
    const int N = p->N, base = 1 << p->ks_basebit;
    const size_t npoly = (size_t)p->n * 2 * p->bk_l * 2;
    // bk: coefficient domain -> spectra, scaled by 2/N so the backward transform needs no scaling
    // (tGswToFFTConvert cb/tgsw_functions.cpp:389-394 ; the 2/N of execute_direct_torus32 :78-100 is folded in here)
    int32_t* tmp = nullptr;
    CU(cudaMalloc(&tmp, npoly * N * sizeof(int32_t)));
    CU(cudaMemcpy(tmp, bk_host, npoly * N * sizeof(int32_t), cudaMemcpyHostToDevice));
    cudaError_t e = launch_poly_to_spectrum32(ctx->g_bkfft, tmp, ctx->tw1024, N, (int)npoly, 2.0 / N, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    // ks: drop d = 0 rows, pad rows to 512 columns
    const size_t raw = (size_t)N * p->ks_t * base * (p->n + 1);
    CU(cudaMalloc(&tmp, raw * sizeof(int32_t)));
    CU(cudaMemcpy(tmp, ks_host, raw * sizeof(int32_t), cudaMemcpyHostToDevice));
    e = gate_ks_repack(ctx->g_ks, tmp, *p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    ctx->gate_ready = true;
    return TFHE_B200_OK;
}

int tfhe_b200_gate_get_params(const tfhe_b200_ctx* ctx, tfhe_b200_gate_params* p) {
    if (!ctx || !p) return TFHE_B200_ERR_PARAM;
    if (!ctx->g_bkfft) return TFHE_B200_ERR_NOKEY;
    *p = ctx->gp;
    return TFHE_B200_OK;
}
int tfhe_b200_gate_key_blob(tfhe_b200_ctx* ctx, int which, void** dev_ptr, size_t* bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(dev_ptr && bytes, "gate_key_blob: null output");
    if (!ctx->g_bkfft) return fail(ctx, TFHE_B200_ERR_NOKEY, "gate_key_blob: gate keys not allocated");
    if (which == 0) { *dev_ptr = ctx->g_bkfft; *bytes = ctx->g_bkfft_bytes; }
    else if (which == 1) { *dev_ptr = ctx->g_ks; *bytes = ctx->g_ks_bytes; }
    else return fail(ctx, TFHE_B200_ERR_PARAM, "gate_key_blob: which must be 0 or 1");
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ wire format of the loaded gate keys */
struct KeyBlobHeader {            // 96 bytes, little endian
    char magic[8];                // "TFHEB200"
    uint32_t version, kind;       // kind 1 = gate keys
    int32_t params[8];            // n, N, k, bk_l, bk_Bgbit, ks_t, ks_basebit, packing of the key-switching key (0 rows, 1 paired rows, 2 byte planes)
    uint64_t bk_bytes, ks_bytes, checksum;
    uint64_t reserved[3];
};
static_assert(sizeof(KeyBlobHeader) == 96, "wire header is 96 bytes");
static const uint32_t kKeyBlobVersion = 3;      // bump whenever the spectral slot order or the key-switch packing changes
static uint64_t fnv1a(const unsigned char* p, size_t n, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
int tfhe_b200_gate_export_keys(tfhe_b200_ctx* ctx, void* buf_host, size_t* bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (!ctx->gate_ready) return fail(ctx, TFHE_B200_ERR_NOKEY, "gate_export_keys: gate keys not loaded");
    NEED(bytes, "gate_export_keys: null size pointer");
    const size_t need = sizeof(KeyBlobHeader) + ctx->g_bkfft_bytes + ctx->g_ks_bytes;
    if (!buf_host) { *bytes = need; return TFHE_B200_OK; }
    NEED(*bytes >= need, "gate_export_keys: buffer too small");
    CU(cudaSetDevice(ctx->device));
    unsigned char* out = (unsigned char*)buf_host;
    CU(cudaMemcpy(out + sizeof(KeyBlobHeader), ctx->g_bkfft, ctx->g_bkfft_bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out + sizeof(KeyBlobHeader) + ctx->g_bkfft_bytes, ctx->g_ks, ctx->g_ks_bytes, cudaMemcpyDeviceToHost));
    KeyBlobHeader h{};
    memcpy(h.magic, "TFHEB200", 8);
    h.version = kKeyBlobVersion; h.kind = 1;
    const tfhe_b200_gate_params& p = ctx->gp;
    const int32_t pv[8] = {p.n, p.N, p.k, p.bk_l, p.bk_Bgbit, p.ks_t, p.ks_basebit, gate_ks_packing_id(p)};
    memcpy(h.params, pv, sizeof(pv));
    h.bk_bytes = ctx->g_bkfft_bytes; h.ks_bytes = ctx->g_ks_bytes;
    h.checksum = fnv1a(out + sizeof(KeyBlobHeader), ctx->g_bkfft_bytes + ctx->g_ks_bytes);
    memcpy(out, &h, sizeof(h));
    *bytes = need;
    return TFHE_B200_OK;
}
int tfhe_b200_gate_import_keys(tfhe_b200_ctx* ctx, const void* buf_host, size_t bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(buf_host && bytes >= sizeof(KeyBlobHeader), "gate_import_keys: buffer too small for a header");
    KeyBlobHeader h;
    memcpy(&h, buf_host, sizeof(h));
    NEED(memcmp(h.magic, "TFHEB200", 8) == 0, "gate_import_keys: bad magic");
    NEED(h.version == kKeyBlobVersion, "gate_import_keys: key blob written by another format version");
    NEED(h.kind == 1, "gate_import_keys: not a gate-key blob");
    NEED(bytes == sizeof(KeyBlobHeader) + h.bk_bytes + h.ks_bytes, "gate_import_keys: size does not match the header");
    const unsigned char* in = (const unsigned char*)buf_host + sizeof(KeyBlobHeader);
    NEED(fnv1a(in, h.bk_bytes + h.ks_bytes) == h.checksum, "gate_import_keys: checksum mismatch");
    tfhe_b200_gate_params p{h.params[0], h.params[1], h.params[2], h.params[3], h.params[4], h.params[5], h.params[6]};
    // everything about the blob is checked BEFORE the keys currently loaded are released
    int rc = check_gate_params(ctx, &p); if (rc) return rc;
    NEED(h.params[7] == gate_ks_packing_id(p), "gate_import_keys: the blob's key-switching key is in another packing (TFHE_B200_KS / TFHE_B200_KS_PAIR differ)");
    NEED((size_t)p.n * 2 * p.bk_l * 2 * (p.N / 2) * sizeof(cplx) == h.bk_bytes &&
         gate_ks_bytes(p) == h.ks_bytes,
         "gate_import_keys: blob sizes do not match the parameters");
    rc = tfhe_b200_gate_alloc_keys(ctx, &p); if (rc) return rc;
    CU(cudaMemcpy(ctx->g_bkfft, in, h.bk_bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->g_ks, in + h.bk_bytes, h.ks_bytes, cudaMemcpyHostToDevice));
    ctx->gate_ready = true;
    return TFHE_B200_OK;
}

// every entry point selects its context's device first: a process may drive several contexts (one per GPU)
#define ENTER() do { if (!ctx) return TFHE_B200_ERR_PARAM; CU(cudaSetDevice(ctx->device)); } while (0)
#define NEED_GATE() do { ENTER(); if (!ctx->gate_ready) return fail(ctx, TFHE_B200_ERR_NOKEY, "gate keys not loaded"); } while (0)

static BRArgs gate_br_args(const tfhe_b200_ctx* ctx, int count) {
    BRArgs a{};
    a.bkfft = ctx->g_bkfft; a.tw = ctx->tw1024;
    a.n = ctx->gp.n; a.l = ctx->gp.bk_l; a.Bgbit = ctx->gp.bk_Bgbit; a.count = count;
    a.out_stride = ctx->gp.N + 1;
    a.n_mu = 1; a.mu_bgbit = 0;
    return a;
}

int tfhe_b200_blindRotate_FFT_batch(tfhe_b200_ctx* ctx, int32_t* accum_dev, const int32_t* bara_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (accum_dev && bara_dev), "null buffer");
    BRArgs a = gate_br_args(ctx, count);
    a.mode = BR_ACCUM; a.accum = accum_dev; a.bara = bara_dev;
    { ProfScope ps(ctx, 0, (cudaStream_t)stream); CU(launch_blind_rotate32(a, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
int tfhe_b200_blindRotateAndExtract_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* v_dev, const int32_t* barb_dev,
                                              const int32_t* bara_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && v_dev && barb_dev && bara_dev), "null buffer");
    BRArgs a = gate_br_args(ctx, count);
    a.mode = BR_TESTVEC; a.v = v_dev; a.barb = barb_dev; a.bara = bara_dev; a.out = result_dev;
    { ProfScope ps(ctx, 0, (cudaStream_t)stream); CU(launch_blind_rotate32(a, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
static int bootstrap_woks(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu, const int32_t* xa, const int32_t* xb, int ka, int kb,
                          int32_t cconst, int count, cudaStream_t s) {
    BRArgs a = gate_br_args(ctx, count);
    a.mode = BR_LWE; a.xa = xa; a.xb = xb; a.ka = ka; a.kb = kb; a.cconst = cconst; a.mu = mu; a.out = result_dev;
    { ProfScope ps(ctx, 0, s); CU(launch_blind_rotate32(a, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_bootstrap_woKS_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu, const int32_t* x_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && x_dev), "null buffer");
    return bootstrap_woks(ctx, result_dev, mu, x_dev, nullptr, 1, 0, 0, count, (cudaStream_t)stream);
}
static int gate_keyswitch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev, int count, cudaStream_t s) {
    KSArgs k{};
    k.in = sample_dev; k.in_stride = ctx->gp.N + 1; k.rows_in = ctx->gp.N; k.t = ctx->gp.ks_t; k.basebit = ctx->gp.ks_basebit;
    if (gate_ks_paired(ctx->gp)) { k.t = ctx->gp.ks_t / 2; k.basebit = 4; }      // same digits, taken two at a time
    k.key = ctx->g_ks; k.cols = ctx->gp.n + 1; k.cols_pad = (int)gate_cols_pad(ctx->gp);
    k.b_col = ctx->gp.n; k.b_index = ctx->gp.N; k.out = result_dev; k.out_stride = ctx->gp.n + 1; k.count = count;
    { ProfScope ps(ctx, 1, s); CU(launch_keyswitch32(k, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_lweKeySwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && sample_dev), "null buffer");
    return gate_keyswitch(ctx, result_dev, sample_dev, count, (cudaStream_t)stream);
}
static int bootstrap_full(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu, const int32_t* xa, const int32_t* xb, int ka, int kb,
                          int32_t cconst, int count, cudaStream_t s) {
    const size_t ubytes = (size_t)count * (ctx->gp.N + 1) * sizeof(int32_t);
    int rc = ensure_scratch(ctx, 0, ubytes); if (rc) return rc;
    ScratchUse use(ctx, s);
    int32_t* u = (int32_t*)ctx->scratch[0];
    rc = bootstrap_woks(ctx, u, mu, xa, xb, ka, kb, cconst, count, s); if (rc) return rc;
    return gate_keyswitch(ctx, result_dev, u, count, s);
}
int tfhe_b200_bootstrap_FFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int32_t mu, const int32_t* x_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && x_dev), "null buffer");
    return bootstrap_full(ctx, result_dev, mu, x_dev, nullptr, 1, 0, 0, count, (cudaStream_t)stream);
}

/* boots* constants [UPSTREAM boot-gates.cpp, SURVEY Appendix C]: tmp = (0, c8/8) + ka*ca + kb*cb */
static const struct { int c8, ka, kb; } kGate[TFHE_B200_NUM_GATES] = {
    {1, -1, -1}, {-1, 1, 1}, {1, 1, 1}, {-1, -1, -1}, {2, 2, 2}, {-2, -2, -2}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, 1}, {1, 1, -1},
};
static const int32_t kMU = 1 << 29;   // modSwitchToTorus32(1, 8)

int tfhe_b200_bootsGate_batch(tfhe_b200_ctx* ctx, int op, int32_t* result_dev, const int32_t* ca_dev, const int32_t* cb_dev, int count, void* stream) {
    NEED_GATE(); NEED(op >= 0 && op < TFHE_B200_NUM_GATES, "bootsGate: unknown gate");
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && ca_dev && cb_dev), "null buffer");
    return bootstrap_full(ctx, result_dev, kMU, ca_dev, cb_dev, kGate[op].ka, kGate[op].kb, (int32_t)((uint32_t)kGate[op].c8 * (uint32_t)kMU),
                          count, (cudaStream_t)stream);
}
int tfhe_b200_bootsNOT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* ca_dev, int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && ca_dev), "null buffer");
    { ProfScope ps(ctx, 2, (cudaStream_t)stream); CU(launch_lwe_lincomb(result_dev, ca_dev, nullptr, -1, 0, 0, ctx->gp.n, count, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
int tfhe_b200_bootsMUX_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* a_dev, const int32_t* b_dev, const int32_t* c_dev,
                             int count, void* stream) {
    NEED_GATE(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && a_dev && b_dev && c_dev), "null buffer");
    cudaStream_t s = (cudaStream_t)stream;
    const int N = ctx->gp.N;
    const size_t ubytes = (size_t)count * (N + 1) * sizeof(int32_t);
    int rc = ensure_scratch(ctx, 0, ubytes); if (rc) return rc;
    rc = ensure_scratch(ctx, 1, ubytes); if (rc) return rc;
    ScratchUse use(ctx, s);
    int32_t* u1 = (int32_t*)ctx->scratch[0]; int32_t* u2 = (int32_t*)ctx->scratch[1];
    // AND(a,b) and AND(not a, c), both without key switch; sum + (0,1/8); one key switch
    rc = bootstrap_woks(ctx, u1, kMU, a_dev, b_dev, 1, 1, -kMU, count, s); if (rc) return rc;
    rc = bootstrap_woks(ctx, u2, kMU, a_dev, c_dev, -1, 1, -kMU, count, s); if (rc) return rc;
    { ProfScope ps(ctx, 2, s); CU(launch_lwe_lincomb(u1, u1, u2, 1, 1, kMU, N, count, s)); }
    return gate_keyswitch(ctx, result_dev, u1, count, s);
}
int tfhe_b200_bootsGate_batch_host(tfhe_b200_ctx* ctx, int op, int32_t* result_host, const int32_t* ca_host, const int32_t* cb_host, int count) {
    NEED_GATE(); NEED(op >= 0 && op < TFHE_B200_NUM_GATES, "bootsGate: unknown gate");
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_host && ca_host && cb_host), "null buffer");
    if (count == 0) return TFHE_B200_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t row = (size_t)ctx->gp.n + 1, urow = (size_t)ctx->gp.N + 1;
    const size_t bytes = (size_t)count * row * sizeof(int32_t);
    int rc = ensure_scratch(ctx, 2, 2 * bytes); if (rc) return rc;
    rc = ensure_scratch(ctx, 3, bytes); if (rc) return rc;
    rc = ensure_scratch(ctx, 0, (size_t)count * urow * sizeof(int32_t)); if (rc) return rc;
    int32_t* da = (int32_t*)ctx->scratch[2]; int32_t* db = da + (size_t)count * row; int32_t* dr = (int32_t*)ctx->scratch[3];
    int32_t* u = (int32_t*)ctx->scratch[0];
    for (int k = 0; k < 2; k++) if (!ctx->hs[k]) CU(cudaStreamCreateWithFlags(&ctx->hs[k], cudaStreamNonBlocking));
    // the private streams are not ordered against the caller's streams: wait for whatever still uses the scratch
    if (ctx->scratch_busy && ctx->scratch_ev) for (int k = 0; k < 2; k++) CU(cudaStreamWaitEvent(ctx->hs[k], ctx->scratch_ev, 0));
    // Chunks of whole waves (8 accumulators per SM), alternating between two streams: the host->device copies of chunk k+1 and the
    // device->host copy of chunk k-1 run under the kernels of chunk k, and the blind rotations of consecutive chunks fill each other's
    // last wave.  Small batches go through as one chunk.
    const int wave = 8 * ctx->sm_count;
    const int nchunk = count >= 8 * wave ? 4 : (count >= 2 * wave ? 2 : 1);
    const int per = ((count + nchunk - 1) / nchunk + wave - 1) / wave * wave;
    const int32_t cst = (int32_t)((uint32_t)kGate[op].c8 * (uint32_t)kMU);
    for (int k = 0, off = 0; off < count; k++, off += per) {
        const int c = count - off < per ? count - off : per;
        cudaStream_t s = ctx->hs[k & 1];
        const size_t o = (size_t)off * row, cb = (size_t)c * row * sizeof(int32_t);
        CU(cudaMemcpyAsync(da + o, ca_host + o, cb, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(db + o, cb_host + o, cb, cudaMemcpyHostToDevice, s));
        rc = bootstrap_woks(ctx, u + (size_t)off * urow, kMU, da + o, db + o, kGate[op].ka, kGate[op].kb, cst, c, s); if (rc) return rc;
        rc = gate_keyswitch(ctx, dr + o, u + (size_t)off * urow, c, s); if (rc) return rc;
        CU(cudaMemcpyAsync(result_host + o, dr + o, cb, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(ctx->hs[0]));
    CU(cudaStreamSynchronize(ctx->hs[1]));
    ctx->scratch_busy = false;            // both streams (and everything they waited for) have drained
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ TRGSW x TRLWE products (N = 1024, Torus32) */
static int check_gadget(tfhe_b200_ctx* ctx, int l, int Bgbit) {
    NEED(l >= 1 && Bgbit >= 1 && Bgbit <= 16 && l * Bgbit <= 32, "bad gadget (l, Bgbit)");
    return TFHE_B200_OK;
}
int tfhe_b200_tGswToFFTConvert_batch(tfhe_b200_ctx* ctx, double* gswfft_dev, const int32_t* gsw_dev, int l, int count, void* stream) {
    ENTER();
    NEED(l >= 1 && l <= 32, "tGswToFFTConvert: bad l"); NEED(count >= 0, "count < 0"); NEED(count == 0 || (gswfft_dev && gsw_dev), "null buffer");
    const long npoly = (long)count * 2 * l * 2;
    NEED(npoly <= 0x7fffffffL, "tGswToFFTConvert: too many polynomials");
    CU(launch_poly_to_spectrum32((cplx*)gswfft_dev, gsw_dev, ctx->tw1024, 1024, (int)npoly, 2.0 / 1024, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
static int extmul(tfhe_b200_ctx* ctx, int32_t* accum_dev, const double* gswfft_dev, size_t stride, int units_per_gsw, int l, int Bgbit,
                  long count, cudaStream_t s) {
    NEED(count <= 0x7fffffffL, "extern mul: batch too large");
    BRArgs a{};
    a.bkfft = (const cplx*)gswfft_dev; a.tw = ctx->tw1024; a.l = l; a.Bgbit = Bgbit; a.count = (int)count; a.mode = BR_EXTMUL;
    a.accum = accum_dev; a.bk_sample_stride = stride; a.units_per_gsw = units_per_gsw;
    { ProfScope ps(ctx, 0, s); CU(launch_extern_mul32(a, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_tGswFFTExternMulToTLwe_batch(tfhe_b200_ctx* ctx, int32_t* accum_dev, const double* gswfft_dev, int per_sample,
                                           int l, int Bgbit, int count, void* stream) {
    ENTER();
    int rc = check_gadget(ctx, l, Bgbit); if (rc) return rc;
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (accum_dev && gswfft_dev), "null buffer");
    return extmul(ctx, accum_dev, gswfft_dev, per_sample ? (size_t)2 * l * 2 * 512 : 0, 1, l, Bgbit, count, (cudaStream_t)stream);
}
int tfhe_b200_CMux_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* gswfft_dev, int per_sample, const int32_t* d1_dev,
                         const int32_t* d0_dev, int l, int Bgbit, int count, void* stream) {
    ENTER();
    int rc = check_gadget(ctx, l, Bgbit); if (rc) return rc;
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && gswfft_dev && d1_dev && d0_dev), "null buffer");
    if (count == 0) return TFHE_B200_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int len = 2 * 1024;
    const size_t bytes = (size_t)count * len * sizeof(int32_t);
    rc = ensure_scratch(ctx, 0, bytes); if (rc) return rc;
    ScratchUse use(ctx, s);
    int32_t* tmp = (int32_t*)ctx->scratch[0];
    // tmp = d1 - d0 ; tmp <- C (x) tmp ; result = tmp + d0      (flat rows of len int32; the lincomb constant is 0)
    { ProfScope ps(ctx, 2, s); CU(launch_lwe_lincomb(tmp, d1_dev, d0_dev, 1, -1, 0, len - 1, count, s)); }
    rc = extmul(ctx, tmp, gswfft_dev, per_sample ? (size_t)2 * l * 2 * 512 : 0, 1, l, Bgbit, count, s); if (rc) return rc;
    { ProfScope ps(ctx, 2, s); CU(launch_lwe_lincomb(result_dev, tmp, d0_dev, 1, 1, 0, len - 1, count, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_LUT_vertical_packing_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* selfft_dev, int nsel,
                                         const int32_t* table_dev, int l, int Bgbit, int count, void* stream) {
    ENTER();
    int rc = check_gadget(ctx, l, Bgbit); if (rc) return rc;
    NEED(nsel >= 1 && nsel <= 16, "LUT: nsel must be 1..16"); NEED(count >= 0, "count < 0");
    NEED(count == 0 || (result_dev && selfft_dev && table_dev), "null buffer");
    if (count == 0) return TFHE_B200_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int N = 1024, len = 2 * N;
    const size_t gsw_cplx = (size_t)2 * l * 2 * 512;                 // one TGSW spectrum
    const size_t lvl0_units = (size_t)count << (nsel - 1);
    NEED(lvl0_units <= 0x7fffffffUL, "LUT: count * 2^(nsel-1) too large");
    // ping-pong level buffers: level j produces count * 2^(nsel-1-j) TRLWE of len int32
    rc = ensure_scratch(ctx, 0, lvl0_units * len * sizeof(int32_t)); if (rc) return rc;
    rc = ensure_scratch(ctx, 1, (lvl0_units / 2 + 1) * len * sizeof(int32_t)); if (rc) return rc;
    ScratchUse use(ctx, s);
    int32_t* pp[2] = {(int32_t*)ctx->scratch[0], (int32_t*)ctx->scratch[1]};
    const int32_t* in = nullptr;
    for (int j = 0; j < nsel; j++) {
        const int nodes = 1 << (nsel - 1 - j);
        const size_t units = (size_t)count * nodes;
        int32_t* out = (j == nsel - 1) ? result_dev : pp[j & 1];
        // out = d1 - d0 : level 0 straight from the plaintext table (trivial TRLWE, the same for every sample), then from pairs
        if (j == 0) { ProfScope ps(ctx, 2, s); CU(launch_lut_table(out, table_dev, N, nodes, count, 0, s)); }
        else        { ProfScope ps(ctx, 2, s); CU(launch_pair_combine(out, in, len, units, 0, s)); }
        // out <- sel_j (x) out ; the TGSW of unit u is selector j of sample u / nodes
        rc = extmul(ctx, out, selfft_dev + (size_t)j * gsw_cplx * 2 /* doubles per cplx */, (size_t)nsel * gsw_cplx, nodes, l, Bgbit, (long)units, s);
        if (rc) return rc;
        // out += d0
        if (j == 0) { ProfScope ps(ctx, 2, s); CU(launch_lut_table(out, table_dev, N, nodes, count, 1, s)); }
        else        { ProfScope ps(ctx, 2, s); CU(launch_pair_combine(out, in, len, units, 1, s)); }
        in = out;
    }
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ gate-level circuits */
int tfhe_b200_circuit_eval_batch(tfhe_b200_ctx* ctx, const tfhe_b200_gate* gates, int n_gates, int32_t* wires_dev, int n_wires,
                                 int count, void* stream) {
    NEED_GATE(); NEED(n_gates >= 0 && n_wires >= 0 && count >= 0, "circuit_eval: negative size");
    NEED(n_gates == 0 || gates, "circuit_eval: null netlist"); NEED(n_gates == 0 || count == 0 || wires_dev, "circuit_eval: null wires");
    if (count == 0) return TFHE_B200_OK;
    const size_t wstride = (size_t)count * (ctx->gp.n + 1);
    auto arity = [](int op) { return op == TFHE_B200_MUX ? 3 : (op == TFHE_B200_NOT || op == TFHE_B200_COPY ? 1 : 2); };
    for (int i = 0; i < n_gates; i++) {
        const tfhe_b200_gate& g = gates[i];
        NEED((g.op >= 0 && g.op < TFHE_B200_NUM_GATES) || g.op == TFHE_B200_NOT || g.op == TFHE_B200_COPY || g.op == TFHE_B200_MUX,
             "circuit_eval: unknown op");
        const int ar = arity(g.op);
        NEED(g.out >= 0 && g.out < n_wires && g.in0 >= 0 && g.in0 < n_wires && (ar < 2 || (g.in1 >= 0 && g.in1 < n_wires)) &&
             (ar < 3 || (g.in2 >= 0 && g.in2 < n_wires)), "circuit_eval: wire index out of range");
    }
    for (int i = 0; i < n_gates;) {
        const tfhe_b200_gate& g = gates[i];
        const int ar = arity(g.op);
        // longest run of "the same gate on the next wires" that keeps the gates independent of each other
        int run = 1;
        while (i + run < n_gates) {
            const tfhe_b200_gate& h = gates[i + run];
            if (h.op != g.op || h.out != g.out + run || h.in0 != g.in0 + run || (ar >= 2 && h.in1 != g.in1 + run) ||
                (ar >= 3 && h.in2 != g.in2 + run)) break;
            auto reads_run_output = [&](int w) { return w >= g.out && w < g.out + run; };     // outputs of gates i .. i+run-1
            if (reads_run_output(h.in0) || (ar >= 2 && reads_run_output(h.in1)) || (ar >= 3 && reads_run_output(h.in2))) break;
            // and no earlier gate of the run may read what this one writes later (write-after-read inside one launch is fine:
            // every launch reads all its inputs before the key switch writes, but keep the semantics of sequential execution)
            bool war = false;
            for (int j = 0; j < run && !war; j++) {
                const tfhe_b200_gate& e = gates[i + j];
                war = e.in0 == h.out || (ar >= 2 && e.in1 == h.out) || (ar >= 3 && e.in2 == h.out);
            }
            if (war) break;
            run++;
        }
        const long total = (long)run * count;
        NEED(total <= 0x7fffffffL, "circuit_eval: run * count overflows");
        int32_t* out = wires_dev + (size_t)g.out * wstride;
        const int32_t* a = wires_dev + (size_t)g.in0 * wstride;
        const int32_t* b = ar >= 2 ? wires_dev + (size_t)g.in1 * wstride : nullptr;
        const int32_t* c = ar >= 3 ? wires_dev + (size_t)g.in2 * wstride : nullptr;
        int rc;
        if (g.op == TFHE_B200_NOT) rc = tfhe_b200_bootsNOT_batch(ctx, out, a, (int)total, stream);
        else if (g.op == TFHE_B200_COPY) {
            rc = TFHE_B200_OK;
            if (out != a) CU(cudaMemcpyAsync(out, a, (size_t)total * (ctx->gp.n + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        } else if (g.op == TFHE_B200_MUX) rc = tfhe_b200_bootsMUX_batch(ctx, out, a, b, c, (int)total, stream);
        else rc = tfhe_b200_bootsGate_batch(ctx, g.op, out, a, b, (int)total, stream);
        if (rc) return rc;
        i += run;
    }
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ standalone transforms */
static const cplx* tw_for(const tfhe_b200_ctx* ctx, int N) { return N == 1024 ? ctx->tw1024 : (N == 2048 ? ctx->tw2048 : nullptr); }

int tfhe_b200_IntPolynomial_ifft_batch(tfhe_b200_ctx* ctx, double* result_dev, const int32_t* poly_dev, int N, int count, void* stream) {
    ENTER();
    NEED(tw_for(ctx, N), "transform: N must be 1024 or 2048"); NEED(count >= 0, "count < 0");
    CU(launch_poly_to_spectrum32((cplx*)result_dev, poly_dev, tw_for(ctx, N), N, count, 1.0, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_TorusPolynomial64_ifft_batch(tfhe_b200_ctx* ctx, double* result_dev, const int64_t* poly_dev, int N, int count, void* stream) {
    ENTER();
    NEED(tw_for(ctx, N), "transform: N must be 1024 or 2048"); NEED(count >= 0, "count < 0");
    CU(launch_poly_to_spectrum64((cplx*)result_dev, poly_dev, tw_for(ctx, N), N, count, 1.0, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_TorusPolynomial_fft_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const double* lagr_dev, int N, int count, void* stream) {
    ENTER();
    NEED(tw_for(ctx, N), "transform: N must be 1024 or 2048"); NEED(count >= 0, "count < 0");
    CU(launch_spectrum_to_torus32(result_dev, (const cplx*)lagr_dev, tw_for(ctx, N), N, count, 2.0 / N, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_TorusPolynomial64_fft_batch(tfhe_b200_ctx* ctx, int64_t* result_dev, const double* lagr_dev, int N, int count, void* stream) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(tw_for(ctx, N), "transform: N must be 1024 or 2048"); NEED(count >= 0, "count < 0");
    CU(launch_spectrum_to_torus64(result_dev, (const cplx*)lagr_dev, tw_for(ctx, N), N, count, 2.0 / N, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_LagrangeHalfCPolynomialAddMul_batch(tfhe_b200_ctx* ctx, double* res_dev, const double* a_dev, const double* b_dev, int N, int count, void* stream) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(N > 0 && N % 2 == 0 && count >= 0, "AddMul: bad sizes");
    CU(launch_spectrum_addmul((cplx*)res_dev, (const cplx*)a_dev, (const cplx*)b_dev, (size_t)count * (N / 2), (cudaStream_t)stream));
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ circuit bootstrapping */
static size_t pad512(size_t c) { return (c + 511) / 512 * 512; }

static int check_cb_params(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p) {
    NEED(p, "cb params: null");
    NEED(p->N_lvl2 == 2048, "cb params: only N_lvl2 = 2048 is compiled");
    NEED(p->N_lvl1 == 1024, "cb params: only N_lvl1 = 1024 is compiled");
    NEED(p->n_lvl0 >= 1 && p->n_lvl0 < 1024, "cb params: n_lvl0 out of range");
    NEED(p->ell_lvl2 >= 1 && p->bgbit_lvl2 >= 1 && p->ell_lvl2 * p->bgbit_lvl2 < 64 && p->bgbit_lvl2 <= 16, "cb params: bad lvl2 gadget");
    NEED(p->ell_lvl1 >= 1 && p->ell_lvl1 <= 4 && p->bgbit_lvl1 >= 1 && p->ell_lvl1 * p->bgbit_lvl1 <= 32, "cb params: bad lvl1 gadget");
    NEED(p->kslength_lvl10 >= 1 && p->kslength_lvl21 >= 1, "cb params: ks length must be >= 1");
    NEED(p->ksbasebit_lvl10 >= 1 && p->ksbasebit_lvl10 <= 3 && p->ksbasebit_lvl21 >= 1 && p->ksbasebit_lvl21 <= 3, "cb params: ks basebit must be 1..3");
    NEED(p->kslength_lvl10 * p->ksbasebit_lvl10 <= 31 && p->kslength_lvl21 * p->ksbasebit_lvl21 <= 32, "cb params: ks length too large (digits must sit in the top 32 bits)");
    return TFHE_B200_OK;
}
// device sizes of the three key blobs: bk spectra, repacked preKS, repacked privKS (both u)
static void cb_blob_bytes(const tfhe_b200_cb_params& p, size_t out[3], size_t* privks_u_stride) {
    out[0] = (size_t)p.n_lvl0 * 2 * p.ell_lvl2 * 2 * (p.N_lvl2 / 2) * sizeof(cplx);
    out[1] = ks_key_bytes(p.N_lvl1, p.kslength_lvl10, p.ksbasebit_lvl10, (int)pad512(p.n_lvl0 + 1));
    const size_t us = ks_key_bytes(p.N_lvl2 + 1, p.kslength_lvl21, p.ksbasebit_lvl21, 2 * p.N_lvl1) / sizeof(int32_t);
    out[2] = 2 * us * sizeof(int32_t);
    if (privks_u_stride) *privks_u_stride = us;
}

/* Multi-GPU replication of the circuit-bootstrap keys, same protocol as the gate keys: allocate, let the caller fill the blobs
 * (which: 0 = bk spectra, 1 = preKS, 2 = privKS), commit. */
int tfhe_b200_cb_alloc_keys(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, int with_privks) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    int rc = check_cb_params(ctx, p); if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    ctx->cb_ready = false;
    cudaFree(ctx->c_bkfft); cudaFree(ctx->c_preks); cudaFree(ctx->c_privks);
    ctx->c_bkfft = nullptr; ctx->c_preks = nullptr; ctx->c_privks = nullptr;
    cudaFree(ctx->c_bkntt); ctx->c_bkntt = nullptr; ctx->cb_exact = false;       // belongs to the previous key
    ctx->cp = *p;
    size_t bytes[3];
    cb_blob_bytes(*p, bytes, &ctx->c_privks_u_stride);
    CU(cudaMalloc(&ctx->c_bkfft, bytes[0]));
    CU(cudaMalloc(&ctx->c_preks, bytes[1]));
    if (with_privks) CU(cudaMalloc(&ctx->c_privks, bytes[2]));
    return TFHE_B200_OK;
}
int tfhe_b200_cb_key_blob(tfhe_b200_ctx* ctx, int which, void** dev_ptr, size_t* bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(dev_ptr && bytes, "cb_key_blob: null output");
    NEED(which >= 0 && which <= 2, "cb_key_blob: which must be 0, 1 or 2");
    if (!ctx->c_bkfft) return fail(ctx, TFHE_B200_ERR_NOKEY, "cb_key_blob: circuit-bootstrap keys not allocated");
    size_t b[3];
    cb_blob_bytes(ctx->cp, b, nullptr);
    void* ptrs[3] = {ctx->c_bkfft, ctx->c_preks, ctx->c_privks};
    if (which == 2 && !ctx->c_privks) return fail(ctx, TFHE_B200_ERR_NOKEY, "cb_key_blob: private key-switch key not allocated");
    *dev_ptr = ptrs[which]; *bytes = b[which];
    return TFHE_B200_OK;
}
int tfhe_b200_cb_commit_keys(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (!ctx->c_bkfft || !ctx->c_preks) return fail(ctx, TFHE_B200_ERR_NOKEY, "cb_commit_keys: circuit-bootstrap keys not allocated");
    ctx->cb_ready = true;
    return TFHE_B200_OK;
}

int tfhe_b200_cb_load_keys(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, const int32_t* preKS_host, const int64_t* bk_host,
                           const int32_t* privKS_host) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(p && preKS_host && bk_host, "cb_load_keys: null pointer");
    int rc = tfhe_b200_cb_alloc_keys(ctx, p, privKS_host != nullptr); if (rc) return rc;
    const int N2 = p->N_lvl2, N1 = p->N_lvl1, n0 = p->n_lvl0;
    // bk (Torus64 coefficients) -> spectra scaled by 2/N   (cb/poc_CircuitBootstrapping.cpp:394-402)
    {
        const size_t npoly = (size_t)n0 * 2 * p->ell_lvl2 * 2;
        int64_t* tmp = nullptr;
        CU(cudaMalloc(&tmp, npoly * N2 * sizeof(int64_t)));
        CU(cudaMemcpy(tmp, bk_host, npoly * N2 * sizeof(int64_t), cudaMemcpyHostToDevice));
        cudaError_t e = launch_poly_to_spectrum64(ctx->c_bkfft, tmp, ctx->tw2048, N2, (int)npoly, 2.0 / N2, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(tmp);
        CU(e);
    }
    {   // preKS [N1][t10][base10][n0+1]
        const int base = 1 << p->ksbasebit_lvl10, t = p->kslength_lvl10;
        const size_t raw = (size_t)N1 * t * base * (n0 + 1), cp = pad512(n0 + 1);
        int32_t* tmp = nullptr;
        CU(cudaMalloc(&tmp, raw * sizeof(int32_t)));
        CU(cudaMemcpy(tmp, preKS_host, raw * sizeof(int32_t), cudaMemcpyHostToDevice));
        cudaError_t e = launch_ks_repack(ctx->c_preks, tmp, N1, t, base, n0 + 1, (int)cp, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(tmp);
        CU(e);
    }
    if (privKS_host) {   // privKS [2][N2+1][t21][base21][2][N1]
        const int base = 1 << p->ksbasebit_lvl21, t = p->kslength_lvl21, cols = 2 * N1;
        const size_t raw_u = (size_t)(N2 + 1) * t * base * cols;
        int32_t* tmp = nullptr;
        // staged in slices of input rows (64 MB at a time) instead of one 1.3 GB temporary per u
        const size_t row_raw = (size_t)t * base * cols;
        int slice = (int)((size_t)(64u << 20) / (row_raw * sizeof(int32_t)));
        if (slice < 1) slice = 1;
        CU(cudaMalloc(&tmp, (size_t)slice * row_raw * sizeof(int32_t)));
        for (int u = 0; u < 2; u++) {
            for (int r0 = 0; r0 <= N2; r0 += slice) {
                const int nr = N2 + 1 - r0 < slice ? N2 + 1 - r0 : slice;
                cudaError_t e = cudaMemcpy(tmp, privKS_host + (size_t)u * raw_u + (size_t)r0 * row_raw, (size_t)nr * row_raw * sizeof(int32_t), cudaMemcpyHostToDevice);
                if (e == cudaSuccess) e = launch_ks_repack_rows(ctx->c_privks + (size_t)u * ctx->c_privks_u_stride, tmp, N2 + 1, r0, nr, t, base, cols, cols, 0);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { cudaFree(tmp); CU(e); }
            }
        }
        cudaFree(tmp);
    }
    ctx->cb_ready = true;
    return TFHE_B200_OK;
}

/* Wire format of the LOADED circuit-bootstrap keys: the same 96-byte header as the gate keys with kind = 2 and
 * the eleven parameters spread over params[8] + reserved, then the three device blobs as they sit in memory.  The 2.35 GB private
 * key-switch key moves in 64 MB slices through a pinned bounce buffer in both directions: no full-size temporary anywhere.
 * The checksum is FNV-1a over the payload in stream order. */
struct CBKeyBlobHeader {          // 128 bytes, little endian
    char magic[8];                // "TFHEB200"
    uint32_t version, kind;       // kind 2 = circuit-bootstrap keys
    int32_t params[12];           // the eleven tfhe_b200_cb_params fields in declaration order, then with_privks
    uint64_t blob_bytes[3], checksum;
    uint64_t reserved[4];
};
static_assert(sizeof(CBKeyBlobHeader) == 128, "cb wire header is 128 bytes");
int tfhe_b200_cb_export_keys(tfhe_b200_ctx* ctx, void* buf_host, size_t* bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (!ctx->cb_ready) return fail(ctx, TFHE_B200_ERR_NOKEY, "cb_export_keys: circuit-bootstrap keys not loaded");
    NEED(bytes, "cb_export_keys: null size pointer");
    size_t b[3];
    cb_blob_bytes(ctx->cp, b, nullptr);
    if (!ctx->c_privks) b[2] = 0;
    const size_t need = sizeof(CBKeyBlobHeader) + b[0] + b[1] + b[2];
    if (!buf_host) { *bytes = need; return TFHE_B200_OK; }
    NEED(*bytes >= need, "cb_export_keys: buffer too small");
    CU(cudaSetDevice(ctx->device));
    unsigned char* out = (unsigned char*)buf_host + sizeof(CBKeyBlobHeader);
    const void* src[3] = {ctx->c_bkfft, ctx->c_preks, ctx->c_privks};
    uint64_t h64 = 1469598103934665603ull;
    const size_t SL = (size_t)64 << 20;
    for (int k = 0; k < 3; k++) {
        for (size_t off = 0; off < b[k]; off += SL) {
            const size_t n = b[k] - off < SL ? b[k] - off : SL;
            CU(cudaMemcpy(out, (const unsigned char*)src[k] + off, n, cudaMemcpyDeviceToHost));
            h64 = fnv1a(out, n, h64);
            out += n;
        }
    }
    CBKeyBlobHeader h{};
    memcpy(h.magic, "TFHEB200", 8);
    h.version = kKeyBlobVersion; h.kind = 2;
    memcpy(h.params, &ctx->cp, 11 * sizeof(int32_t));
    h.params[11] = ctx->c_privks ? 1 : 0;
    for (int k = 0; k < 3; k++) h.blob_bytes[k] = b[k];
    h.checksum = h64;
    memcpy(buf_host, &h, sizeof(h));
    *bytes = need;
    return TFHE_B200_OK;
}
int tfhe_b200_cb_import_keys(tfhe_b200_ctx* ctx, const void* buf_host, size_t bytes) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(buf_host && bytes >= sizeof(CBKeyBlobHeader), "cb_import_keys: buffer too small for a header");
    CBKeyBlobHeader h;
    memcpy(&h, buf_host, sizeof(h));
    NEED(memcmp(h.magic, "TFHEB200", 8) == 0, "cb_import_keys: bad magic");
    NEED(h.version == kKeyBlobVersion, "cb_import_keys: key blob written by another format version");
    NEED(h.kind == 2, "cb_import_keys: not a circuit-bootstrap key blob");
    tfhe_b200_cb_params p;
    memcpy(&p, h.params, 11 * sizeof(int32_t));
    const int with_privks = h.params[11] != 0;
    int rc = check_cb_params(ctx, &p); if (rc) return rc;
    size_t b[3];
    cb_blob_bytes(p, b, nullptr);
    if (!with_privks) b[2] = 0;
    NEED(h.blob_bytes[0] == b[0] && h.blob_bytes[1] == b[1] && h.blob_bytes[2] == b[2], "cb_import_keys: blob sizes do not match the parameters");
    NEED(bytes == sizeof(CBKeyBlobHeader) + b[0] + b[1] + b[2], "cb_import_keys: size does not match the header");
    const unsigned char* in = (const unsigned char*)buf_host + sizeof(CBKeyBlobHeader);
    NEED(fnv1a(in, b[0] + b[1] + b[2]) == h.checksum, "cb_import_keys: checksum mismatch");
    // everything checked: only now are the keys currently loaded released
    rc = tfhe_b200_cb_alloc_keys(ctx, &p, with_privks); if (rc) return rc;
    void* dst[3] = {ctx->c_bkfft, ctx->c_preks, ctx->c_privks};
    for (int k = 0; k < 3; k++) {
        if (b[k]) CU(cudaMemcpy(dst[k], in, b[k], cudaMemcpyHostToDevice));      // straight from the caller's buffer: no staging copy
        in += b[k];
    }
    ctx->cb_ready = true;
    return TFHE_B200_OK;
}

/* Ciphertext wire format (the reference has none): a 64-byte header -- magic "TFHEB2CT", version, kind
 * (1 LWE32, 2 LWE64, 3 TLWE32, 4 TGSW32), up to four dimensions (slowest first, e.g. {count, n+1}), payload bytes, FNV-1a checksum --
 * followed by the samples in the flat layouts of this header's first comment.  Host-side only (no device work): what a client sends
 * to the server that owns the context.  pack: buf = NULL returns the size. */
struct CtHeader { char magic[8]; uint32_t version, kind; int64_t dims[4]; uint64_t payload_bytes, checksum; };
static_assert(sizeof(CtHeader) == 64, "ciphertext wire header is 64 bytes");
static size_t ct_elem_bytes(int kind) { return kind == 2 ? 8 : 4; }
int tfhe_b200_ciphertext_pack(int kind, const int64_t dims[4], const void* samples_host, void* buf_host, size_t* bytes) {
    tfhe_b200_ctx* ctx = nullptr;
    NEED(kind >= 1 && kind <= 4 && dims && bytes, "ciphertext_pack: bad arguments");
    size_t n = ct_elem_bytes(kind);
    for (int i = 0; i < 4; i++) { NEED(dims[i] >= 0 && dims[i] < ((int64_t)1 << 40), "ciphertext_pack: bad dimension"); if (dims[i] > 0) n *= (size_t)dims[i]; }
    const size_t need = sizeof(CtHeader) + n;
    if (!buf_host) { *bytes = need; return TFHE_B200_OK; }
    NEED(samples_host && *bytes >= need, "ciphertext_pack: buffer too small");
    CtHeader h{};
    memcpy(h.magic, "TFHEB2CT", 8);
    h.version = 1; h.kind = (uint32_t)kind;
    for (int i = 0; i < 4; i++) h.dims[i] = dims[i];
    h.payload_bytes = n;
    h.checksum = fnv1a((const unsigned char*)samples_host, n);
    memcpy(buf_host, &h, sizeof(h));
    memcpy((unsigned char*)buf_host + sizeof(h), samples_host, n);
    *bytes = need;
    return TFHE_B200_OK;
}
int tfhe_b200_ciphertext_unpack(const void* buf_host, size_t bytes, int* kind, int64_t dims[4], void* samples_host, size_t* sample_bytes) {
    tfhe_b200_ctx* ctx = nullptr;
    NEED(buf_host && bytes >= sizeof(CtHeader) && sample_bytes, "ciphertext_unpack: buffer too small for a header");
    CtHeader h;
    memcpy(&h, buf_host, sizeof(h));
    NEED(memcmp(h.magic, "TFHEB2CT", 8) == 0, "ciphertext_unpack: bad magic");
    NEED(h.version == 1, "ciphertext_unpack: unknown format version");
    NEED(h.kind >= 1 && h.kind <= 4, "ciphertext_unpack: unknown kind");
    size_t n = ct_elem_bytes((int)h.kind);
    for (int i = 0; i < 4; i++) { NEED(h.dims[i] >= 0 && h.dims[i] < ((int64_t)1 << 40), "ciphertext_unpack: bad dimension"); if (h.dims[i] > 0) n *= (size_t)h.dims[i]; }
    NEED(n == h.payload_bytes && bytes == sizeof(CtHeader) + n, "ciphertext_unpack: size does not match the header");
    if (kind) *kind = (int)h.kind;
    if (dims) for (int i = 0; i < 4; i++) dims[i] = h.dims[i];
    if (!samples_host) { *sample_bytes = n; return TFHE_B200_OK; }
    NEED(*sample_bytes >= n, "ciphertext_unpack: output buffer too small");
    const unsigned char* in = (const unsigned char*)buf_host + sizeof(CtHeader);
    NEED(fnv1a(in, n) == h.checksum, "ciphertext_unpack: checksum mismatch");
    memcpy(samples_host, in, n);
    *sample_bytes = n;
    return TFHE_B200_OK;
}
#define NEED_CB() do { ENTER(); if (!ctx->cb_ready) return fail(ctx, TFHE_B200_ERR_NOKEY, "circuit-bootstrap keys not loaded"); } while (0)

int tfhe_b200_preKeySwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* x_dev, int count, void* stream) {
    NEED_CB(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && x_dev), "null buffer");
    const tfhe_b200_cb_params& p = ctx->cp;
    KSArgs k{};
    k.in = x_dev; k.in_stride = p.N_lvl1 + 1; k.rows_in = p.N_lvl1; k.t = p.kslength_lvl10; k.basebit = p.ksbasebit_lvl10;
    k.key = ctx->c_preks; k.cols = p.n_lvl0 + 1; k.cols_pad = (int)pad512(p.n_lvl0 + 1);
    k.b_col = p.n_lvl0; k.b_index = p.N_lvl1; k.out = result_dev; k.out_stride = p.n_lvl0 + 1; k.count = count;
    { ProfScope ps(ctx, 1, (cudaStream_t)stream); CU(launch_keyswitch32(k, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
int tfhe_b200_preModSwitch_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* x_dev, int count, void* stream) {
    NEED_CB(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && x_dev), "null buffer");
    { ProfScope ps(ctx, 2, (cudaStream_t)stream); CU(launch_modswitch(result_dev, x_dev, 12 /* 2*N2 = 4096 */, (size_t)count * (ctx->cp.n_lvl0 + 1), (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
static ExactArgs exact_args(const tfhe_b200_ctx* ctx, int count) {
    const tfhe_b200_cb_params& p = ctx->cp;
    ExactArgs a{};
    a.key = ctx->c_bkntt; a.psi_rev = ctx->ntt_tab; a.psi_inv_rev = ctx->ntt_tab + p.N_lvl2; a.n_inv = ctx->ntt_n_inv;
    a.n = p.n_lvl0; a.l = p.ell_lvl2; a.Bgbit = p.bgbit_lvl2; a.count = count; a.out_stride = p.N_lvl2 + 1; a.n_mu = 1;
    return a;
}
static int cb_woks(tfhe_b200_ctx* ctx, int64_t* result_dev, int64_t mu, int n_mu, int mu_bgbit, const int32_t* abar_dev, int count, cudaStream_t s) {
    const tfhe_b200_cb_params& p = ctx->cp;
    if (ctx->cb_exact) {
        ExactArgs a = exact_args(ctx, count);
        a.mode = BR_LWE; a.bara = abar_dev; a.mu = mu; a.n_mu = n_mu; a.mu_bgbit = mu_bgbit; a.out = result_dev;
        { ProfScope ps(ctx, 0, s); CU(launch_exact_blind_rotate(a, s)); }
        return TFHE_B200_OK;
    }
    BRArgs a{};
    a.bkfft = ctx->c_bkfft; a.tw = ctx->tw2048; a.n = p.n_lvl0; a.l = p.ell_lvl2; a.Bgbit = p.bgbit_lvl2; a.count = count;
    a.mode = BR_LWE; a.bara = abar_dev; a.mu = mu; a.n_mu = n_mu; a.mu_bgbit = mu_bgbit; a.out = result_dev; a.out_stride = p.N_lvl2 + 1;
    { ProfScope ps(ctx, 0, s); CU(launch_blind_rotate64(a, s)); }
    return TFHE_B200_OK;
}
// the blind-rotation loop alone on Torus64 accumulators (cb/poc_CircuitBootstrapping.cpp:580-642 with defects D1/D2 corrected):
// accum[B][2][N2] in/out, bara[B][n0] rotation amounts in [0, 2*N2)
int tfhe_b200_blindRotate64_FFT_batch(tfhe_b200_ctx* ctx, int64_t* accum_dev, const int32_t* bara_dev, int count, void* stream) {
    NEED_CB(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (accum_dev && bara_dev), "null buffer");
    const tfhe_b200_cb_params& p = ctx->cp;
    BRArgs a{};
    a.bkfft = ctx->c_bkfft; a.tw = ctx->tw2048; a.n = p.n_lvl0; a.l = p.ell_lvl2; a.Bgbit = p.bgbit_lvl2; a.count = count;
    a.mode = BR_ACCUM; a.accum = accum_dev; a.bara = bara_dev; a.n_mu = 1; a.out_stride = p.N_lvl2 + 1;
    { ProfScope ps(ctx, 0, (cudaStream_t)stream); CU(launch_blind_rotate64(a, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
/* ---- exact Torus64 path */
int tfhe_b200_cb_load_exact_key(tfhe_b200_ctx* ctx, const int64_t* bk_host) {
    NEED_CB(); NEED(bk_host, "cb_load_exact_key: null key");
    const tfhe_b200_cb_params& p = ctx->cp;
    const int N2 = p.N_lvl2;
    NEED(N2 == 2048, "cb_load_exact_key: N_lvl2 must be 2048");
    // sum over 2l digit polynomials of N terms |digit| <= Bg/2 times a 32-bit limb must stay inside (-p/2, p/2): exact_ntt.cuh
    NEED(1 + p.ell_lvl2 > 0 && (p.bgbit_lvl2 - 1) + 32 + 11 + 4 <= 62, "cb_load_exact_key: gadget too wide for the limb split");
    const size_t npoly = (size_t)p.n_lvl0 * 2 * p.ell_lvl2 * 2;
    if (!ctx->ntt_tab) {
        std::vector<uint64_t> tab((size_t)2 * N2);
        gl_make_tables(11, tab.data(), tab.data() + N2, &ctx->ntt_n_inv);
        CU(cudaMalloc(&ctx->ntt_tab, tab.size() * sizeof(uint64_t)));
        CU(cudaMemcpy(ctx->ntt_tab, tab.data(), tab.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    if (ctx->c_bkntt) { CU(cudaFree(ctx->c_bkntt)); ctx->c_bkntt = nullptr; }
    CU(cudaMalloc(&ctx->c_bkntt, npoly * 2 * N2 * sizeof(uint64_t)));
    int64_t* tmp = nullptr;
    CU(cudaMalloc(&tmp, npoly * N2 * sizeof(int64_t)));
    cudaError_t e = cudaMemcpy(tmp, bk_host, npoly * N2 * sizeof(int64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_exact_key(ctx->c_bkntt, tmp, ctx->ntt_tab, N2, npoly, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(tmp);
    CU(e);
    return TFHE_B200_OK;
}
int tfhe_b200_cb_set_exact(tfhe_b200_ctx* ctx, int on) {
    NEED_CB();
    if (on && !ctx->c_bkntt) return fail(ctx, TFHE_B200_ERR_NOKEY, "cb_set_exact: call tfhe_b200_cb_load_exact_key first");
    ctx->cb_exact = on != 0;
    return TFHE_B200_OK;
}
int tfhe_b200_blindRotate64_exact_batch(tfhe_b200_ctx* ctx, int64_t* accum_dev, const int32_t* bara_dev, int count, void* stream) {
    NEED_CB(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (accum_dev && bara_dev), "null buffer");
    if (!ctx->c_bkntt) return fail(ctx, TFHE_B200_ERR_NOKEY, "blindRotate64_exact: call tfhe_b200_cb_load_exact_key first");
    ExactArgs a = exact_args(ctx, count);
    a.mode = BR_ACCUM; a.accum = accum_dev; a.bara = bara_dev;
    { ProfScope ps(ctx, 0, (cudaStream_t)stream); CU(launch_exact_blind_rotate(a, (cudaStream_t)stream)); }
    return TFHE_B200_OK;
}
int tfhe_b200_circuitBootstrapWoKS_batch(tfhe_b200_ctx* ctx, int64_t* result_dev, int64_t mu, const int32_t* abar_dev, int count, void* stream) {
    NEED_CB(); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && abar_dev), "null buffer");
    return cb_woks(ctx, result_dev, mu, 1, 0, abar_dev, count, (cudaStream_t)stream);
}
static int cb_privks(tfhe_b200_ctx* ctx, int32_t* result_dev, int out_stride, int u, const int64_t* x_dev, int in_stride, int count, cudaStream_t s) {
    const tfhe_b200_cb_params& p = ctx->cp;
    KSArgs k{};
    k.in = x_dev; k.in_stride = in_stride; k.rows_in = p.N_lvl2 + 1; k.t = p.kslength_lvl21; k.basebit = p.ksbasebit_lvl21;
    k.key = ctx->c_privks + (size_t)u * ctx->c_privks_u_stride; k.cols = 2 * p.N_lvl1; k.cols_pad = 2 * p.N_lvl1;
    k.b_col = -1; k.b_index = 0; k.out = result_dev; k.out_stride = out_stride; k.count = count;
    { ProfScope ps(ctx, 1, s); CU(launch_keyswitch64(k, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_circuitPrivKS_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, int u, const int64_t* x_dev, int count, void* stream) {
    NEED_CB(); NEED(ctx->c_privks, "circuitPrivKS: private key-switch key was not loaded");
    NEED(u == 0 || u == 1, "circuitPrivKS: u must be 0 or 1"); NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && x_dev), "null buffer");
    return cb_privks(ctx, result_dev, 2 * ctx->cp.N_lvl1, u, x_dev, ctx->cp.N_lvl2 + 1, count, (cudaStream_t)stream);
}
// one chunk of the circuit bootstrap on stream s with explicit scratch pointers (pre / abar / boot hold `count` samples):
// preKeySwitch :832, preModSwitch :836, both mu_w in ONE pass over bk (boot[B][ell1][N2+1]; the reference makes two passes :845-847),
// then result[B][u][w][2][N1]: all four private key switches (u,w) in one launch :852-855 -- samples are the B*ell1 rows of boot,
// grid.z walks u (key and output offset)
static int cb_chunk(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev, int32_t* pre, int32_t* abar, int64_t* boot,
                    int count, cudaStream_t s) {
    const tfhe_b200_cb_params& p = ctx->cp;
    const int ell1 = p.ell_lvl1, N1 = p.N_lvl1, N2 = p.N_lvl2;
    int rc;
    if ((rc = tfhe_b200_preKeySwitch_batch(ctx, pre, sample_dev, count, s))) return rc;
    if ((rc = tfhe_b200_preModSwitch_batch(ctx, abar, pre, count, s))) return rc;
    if ((rc = cb_woks(ctx, boot, 0, ell1, p.bgbit_lvl1, abar, count, s))) return rc;
    KSArgs k{};
    k.in = boot; k.in_stride = N2 + 1; k.rows_in = N2 + 1; k.t = p.kslength_lvl21; k.basebit = p.ksbasebit_lvl21;
    k.key = ctx->c_privks; k.cols = 2 * N1; k.cols_pad = 2 * N1; k.b_col = -1; k.b_index = 0;
    k.out = result_dev; k.count = count * ell1; k.group = ell1; k.out_stride = 2 * ell1 * 2 * N1; k.out_inner = 2 * N1;
    k.nz = 2; k.key_z_stride = ctx->c_privks_u_stride; k.out_z_stride = (size_t)ell1 * 2 * N1;
    { ProfScope ps(ctx, 1, s); CU(launch_keyswitch64(k, s)); }
    return TFHE_B200_OK;
}
int tfhe_b200_CircuitBootstrapFFT_batch(tfhe_b200_ctx* ctx, int32_t* result_dev, const int32_t* sample_dev, int count, void* stream) {
    NEED_CB(); NEED(ctx->c_privks, "CircuitBootstrapFFT: private key-switch key was not loaded");
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_dev && sample_dev), "null buffer");
    const tfhe_b200_cb_params& p = ctx->cp;
    cudaStream_t s = (cudaStream_t)stream;
    const int ell1 = p.ell_lvl1, n0 = p.n_lvl0, N2 = p.N_lvl2;
    int rc;
    if ((rc = ensure_scratch(ctx, 0, (size_t)count * (n0 + 1) * sizeof(int32_t)))) return rc;
    if ((rc = ensure_scratch(ctx, 1, (size_t)count * (n0 + 1) * sizeof(int32_t)))) return rc;
    if ((rc = ensure_scratch(ctx, 2, (size_t)count * ell1 * (N2 + 1) * sizeof(int64_t)))) return rc;
    ScratchUse use(ctx, s);
    int32_t* pre = (int32_t*)ctx->scratch[0]; int32_t* abar = (int32_t*)ctx->scratch[1]; int64_t* boot = (int64_t*)ctx->scratch[2];
    return cb_chunk(ctx, result_dev, sample_dev, pre, abar, boot, count, s);
}
int tfhe_b200_CircuitBootstrapFFT_batch_host(tfhe_b200_ctx* ctx, int32_t* result_host, const int32_t* sample_host, int count) {
    NEED_CB(); NEED(ctx->c_privks, "CircuitBootstrapFFT: private key-switch key was not loaded");
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (result_host && sample_host), "null buffer");
    if (count == 0) return TFHE_B200_OK;
    const tfhe_b200_cb_params& p = ctx->cp;
    const int ell1 = p.ell_lvl1, n0 = p.n_lvl0, N2 = p.N_lvl2;
    const size_t in_row = (size_t)p.N_lvl1 + 1, out_row = (size_t)2 * ell1 * 2 * p.N_lvl1;
    const size_t in_bytes = (size_t)count * in_row * sizeof(int32_t), out_bytes = (size_t)count * out_row * sizeof(int32_t);
    int rc;
    if ((rc = ensure_scratch(ctx, 3, in_bytes + out_bytes))) return rc;
    if ((rc = ensure_scratch(ctx, 0, (size_t)count * (n0 + 1) * sizeof(int32_t)))) return rc;
    if ((rc = ensure_scratch(ctx, 1, (size_t)count * (n0 + 1) * sizeof(int32_t)))) return rc;
    if ((rc = ensure_scratch(ctx, 2, (size_t)count * ell1 * (N2 + 1) * sizeof(int64_t)))) return rc;
    int32_t* din = (int32_t*)ctx->scratch[3]; int32_t* dout = (int32_t*)((char*)ctx->scratch[3] + in_bytes);
    int32_t* pre = (int32_t*)ctx->scratch[0]; int32_t* abar = (int32_t*)ctx->scratch[1]; int64_t* boot = (int64_t*)ctx->scratch[2];
    for (int k = 0; k < 2; k++) if (!ctx->hs[k]) CU(cudaStreamCreateWithFlags(&ctx->hs[k], cudaStreamNonBlocking));
    if (ctx->scratch_busy && ctx->scratch_ev) for (int k = 0; k < 2; k++) CU(cudaStreamWaitEvent(ctx->hs[k], ctx->scratch_ev, 0));
    // Chunks of whole waves of the N = 2048 blind rotation (4 accumulators per SM, ell1 accumulators per sample) on the two private
    // streams: the 32 KB-per-sample result of one chunk goes back to the host under the kernels of the next (134 MB per 4096 samples).
    const int wave = 4 * ctx->sm_count / (ell1 > 0 ? ell1 : 1);
    const int nchunk = count >= 8 * wave ? 4 : (count >= 2 * wave ? 2 : 1);
    const int per = wave > 0 ? ((count + nchunk - 1) / nchunk + wave - 1) / wave * wave : count;
    for (int k = 0, off = 0; off < count; k++, off += per) {
        const int c = count - off < per ? count - off : per;
        cudaStream_t s = ctx->hs[k & 1];
        CU(cudaMemcpyAsync(din + (size_t)off * in_row, sample_host + (size_t)off * in_row, (size_t)c * in_row * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        rc = cb_chunk(ctx, dout + (size_t)off * out_row, din + (size_t)off * in_row, pre + (size_t)off * (n0 + 1), abar + (size_t)off * (n0 + 1),
                      boot + (size_t)off * ell1 * (N2 + 1), c, s);
        if (rc) return rc;
        CU(cudaMemcpyAsync(result_host + (size_t)off * out_row, dout + (size_t)off * out_row, (size_t)c * out_row * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(ctx->hs[0]));
    CU(cudaStreamSynchronize(ctx->hs[1]));
    ctx->scratch_busy = false;
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ key generation on the device */

int tfhe_b200_gate_keygen(tfhe_b200_ctx* ctx, const tfhe_b200_gate_params* p, double bk_stdev, double ks_stdev, uint64_t seed,
                          int32_t* lwe_key_host, int32_t* tlwe_key_host, int32_t* bk_raw_host, int32_t* ks_raw_host) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(lwe_key_host && tlwe_key_host, "gate_keygen: the secret keys need somewhere to go");
    NEED(bk_stdev >= 0 && ks_stdev >= 0, "gate_keygen: negative noise");
    int rc = tfhe_b200_gate_alloc_keys(ctx, p); if (rc) return rc;
    const int n = p->n, N = p->N, l = p->bk_l, base = 1 << p->ks_basebit;
    uint64_t st = seed ^ 0x5851F42D4C957F2Dull;
    binary_key(lwe_key_host, n, st);                       // LweKeyGen cb/lwe_functions.cpp:35-41
    binary_key(tlwe_key_host, N, st);                      // tLweKeyGen cb/tlwe_functions.cpp:60-68
    DevTmp s_lwe, i_lwe, s_tlwe, i_tlwe; int w_lwe = 0, w_tlwe = 0;
    if ((rc = upload_key(ctx, s_lwe, i_lwe, &w_lwe, lwe_key_host, n))) return rc;
    if ((rc = upload_key(ctx, s_tlwe, i_tlwe, &w_tlwe, tlwe_key_host, N))) return rc;
    // bk[i] = TGSW(s_i) under the TLWE key (cb/lwe_functions.cpp:504-506), coefficient domain, then the same transform as gate_load_keys
    const size_t npoly = (size_t)n * 2 * l * 2;
    DevTmp bk;
    CU(cudaMalloc(&bk.p, npoly * N * sizeof(int32_t)));
    CU(launch_tlwe_gadget_keygen32(bk.as<int32_t>(), i_tlwe.as<int32_t>(), w_tlwe, s_lwe.as<int32_t>(), n, l, p->bk_Bgbit, bk_stdev, seed, 1u, 0));
    CU(launch_poly_to_spectrum32(ctx->g_bkfft, bk.as<int32_t>(), ctx->tw1024, N, (int)npoly, 2.0 / N, 0));
    if (bk_raw_host) CU(cudaMemcpy(bk_raw_host, bk.p, npoly * N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    // ks[i][j][d] = LWE(s'_i d 2^(32-(j+1)basebit)) under the LWE key (cb/lwe_functions.cpp:113-133)
    const size_t raw = (size_t)N * p->ks_t * base * (n + 1);
    DevTmp ks;
    CU(cudaMalloc(&ks.p, raw * sizeof(int32_t)));
    CU(launch_lwe_ks_keygen(ks.as<int32_t>(), s_tlwe.as<int32_t>(), s_lwe.as<int32_t>(), N, n, p->ks_t, p->ks_basebit, ks_stdev, seed, 2u, 0));
    CU(gate_ks_repack(ctx->g_ks, ks.as<int32_t>(), *p));
    if (ks_raw_host) CU(cudaMemcpy(ks_raw_host, ks.p, raw * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CU(cudaDeviceSynchronize());
    ctx->gate_ready = true;
    return TFHE_B200_OK;
}

int tfhe_b200_cb_keygen(tfhe_b200_ctx* ctx, const tfhe_b200_cb_params* p, double bkstdev_lvl2, double ksstdev_lvl10, double ksstdev_lvl21,
                        uint64_t seed, int32_t* key_lvl0_host, int32_t* key_lvl1_host, int32_t* key_lvl2_host, int with_privks) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(key_lvl0_host && key_lvl1_host && key_lvl2_host, "cb_keygen: the secret keys need somewhere to go");
    int rc = tfhe_b200_cb_alloc_keys(ctx, p, with_privks); if (rc) return rc;
    const int n0 = p->n_lvl0, N1 = p->N_lvl1, N2 = p->N_lvl2, l2 = p->ell_lvl2;
    uint64_t st = seed ^ 0x2545F4914F6CDD1Dull;
    binary_key(key_lvl0_host, n0, st);                     // poc:357-369
    binary_key(key_lvl1_host, N1, st);
    binary_key(key_lvl2_host, N2, st);
    key_lvl2_host[N2] = -1;                                // extended key: b is switched like an a-coefficient (:365-367)
    DevTmp s0, i0, s1, i1, s2, i2; int w0 = 0, w1 = 0, w2 = 0;
    if ((rc = upload_key(ctx, s0, i0, &w0, key_lvl0_host, n0))) return rc;
    if ((rc = upload_key(ctx, s1, i1, &w1, key_lvl1_host, N1))) return rc;
    {   // lvl2 key: N2 binary coefficients as the TLWE64 key, N2 + 1 entries (the last is -1) as the key being switched
        std::vector<int32_t> set;
        for (int i = 0; i < N2; i++) if (key_lvl2_host[i]) set.push_back(i);
        w2 = (int)set.size();
        CU(cudaMalloc(&s2.p, sizeof(int32_t) * (N2 + 1)));
        CU(cudaMemcpy(s2.p, key_lvl2_host, sizeof(int32_t) * (N2 + 1), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&i2.p, sizeof(int32_t) * (set.size() + 1)));
        if (!set.empty()) CU(cudaMemcpy(i2.p, set.data(), sizeof(int32_t) * set.size(), cudaMemcpyHostToDevice));
    }
    {   // preKS: LWE(key_lvl1[i] u 2^(32-(j+1)b)) under key_lvl0  (:372-383)
        const int base = 1 << p->ksbasebit_lvl10, t = p->kslength_lvl10;
        DevTmp raw;
        CU(cudaMalloc(&raw.p, (size_t)N1 * t * base * (n0 + 1) * sizeof(int32_t)));
        CU(launch_lwe_ks_keygen(raw.as<int32_t>(), s1.as<int32_t>(), s0.as<int32_t>(), N1, n0, t, p->ksbasebit_lvl10, ksstdev_lvl10, seed, 3u, 0));
        CU(launch_ks_repack(ctx->c_preks, raw.as<int32_t>(), N1, t, base, n0 + 1, (int)pad512(n0 + 1), 0));
        CU(cudaDeviceSynchronize());
    }
    {   // bk: TGSW64(key_lvl0[i]) under key_lvl2  (:388-391, tGsw64Encrypt_lvl2 :215-227)
        const size_t npoly = (size_t)n0 * 2 * l2 * 2;
        DevTmp raw;
        CU(cudaMalloc(&raw.p, npoly * N2 * sizeof(int64_t)));
        CU(launch_tlwe_gadget_keygen64(raw.as<int64_t>(), i2.as<int32_t>(), w2, s0.as<int32_t>(), n0, l2, p->bgbit_lvl2, bkstdev_lvl2, seed, 4u, 0));
        CU(launch_poly_to_spectrum64(ctx->c_bkfft, raw.as<int64_t>(), ctx->tw2048, N2, (int)npoly, 2.0 / N2, 0));
        CU(cudaDeviceSynchronize());
    }
    if (with_privks) {   // privKS[z][i][j][d] = TLWE32(0) + key_lvl2[i] d 2^(32-(j+1)b) on polynomial z  (:406-419), in slices of input rows
        const int base = 1 << p->ksbasebit_lvl21, t = p->kslength_lvl21, cols = 2 * N1;
        const size_t row_raw = (size_t)t * base * cols;                       // int32 per input row i
        int slice = (int)((size_t)(256u << 20) / (row_raw * sizeof(int32_t)));
        if (slice < 1) slice = 1;
        DevTmp raw;
        CU(cudaMalloc(&raw.p, (size_t)slice * row_raw * sizeof(int32_t)));
        for (int z = 0; z < 2; z++)
            for (int r0 = 0; r0 <= N2; r0 += slice) {
                const int nr = N2 + 1 - r0 < slice ? N2 + 1 - r0 : slice;
                const size_t row0 = ((size_t)z * (N2 + 1) + r0) * t * base;
                CU(launch_tlwe_privks_keygen(raw.as<int32_t>(), i1.as<int32_t>(), w1, s2.as<int32_t>(), N2 + 1, t, p->ksbasebit_lvl21, ksstdev_lvl21,
                                             seed, 5u, row0, (size_t)nr * t * base, 0));
                CU(launch_ks_repack_rows(ctx->c_privks + (size_t)z * ctx->c_privks_u_stride, raw.as<int32_t>(), N2 + 1, r0, nr, t, base, cols, cols, 0));
                CU(cudaDeviceSynchronize());
            }
    }
    ctx->cb_ready = true;
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ high-precision FFT */
static int hp_tables(tfhe_b200_ctx* ctx, int N, const uint64_t** om, const uint64_t** ob) {
    NEED(N == 2048 || N == 4096, "hp FFT: N must be 2048 or 4096");
    const int slot = N == 2048 ? 0 : 1;
    if (!ctx->hp_omega[slot]) {
        const int n = 2 * N;
        std::vector<uint64_t> h((size_t)4 * n);
        for (int inv = 0; inv < 2; inv++) {
            make_hp_tables(n, inv, h.data());
            uint64_t** dst = inv ? &ctx->hp_ombar[slot] : &ctx->hp_omega[slot];
            CU(cudaMalloc(dst, h.size() * sizeof(uint64_t)));
            CU(cudaMemcpy(*dst, h.data(), h.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        }
    }
    *om = ctx->hp_omega[slot]; *ob = ctx->hp_ombar[slot];
    return TFHE_B200_OK;
}
int tfhe_b200_hp_iFFT_batch(tfhe_b200_ctx* ctx, tfhe_b200_cplx96* out_dev, const int64_t* in_dev, int N, int count, void* stream) {
    ENTER();
    const uint64_t *om, *ob; int rc = hp_tables(ctx, N, &om, &ob); if (rc) return rc;
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (out_dev && in_dev), "null buffer");
    CU(launch_hp_ifft(out_dev, in_dev, om, N, count, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_hp_FFT_batch(tfhe_b200_ctx* ctx, int64_t* out_dev, const tfhe_b200_cplx96* in_dev, int N, int count, void* stream) {
    ENTER();
    const uint64_t *om, *ob; int rc = hp_tables(ctx, N, &om, &ob); if (rc) return rc;
    NEED(count >= 0, "count < 0"); NEED(count == 0 || (out_dev && in_dev), "null buffer");
    CU(launch_hp_fft(out_dev, in_dev, ob, N, count, (cudaStream_t)stream));
    return TFHE_B200_OK;
}

/* ------------------------------------------------------------------ diagnostics */
int tfhe_b200_profile_enable(tfhe_b200_ctx* ctx, int on) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->profiling = on != 0;
    return TFHE_B200_OK;
}
int tfhe_b200_profile_read(tfhe_b200_ctx* ctx, double ms[3], int launches[3]) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(ms && launches, "profile_read: null output");
    for (int i = 0; i < 3; i++) { ms[i] = 0; launches[i] = 0; }
    for (auto& sp : ctx->spans) {
        CU(cudaEventSynchronize(sp.b));
        float t = 0; CU(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.cat] += t; launches[sp.cat]++;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
    return TFHE_B200_OK;
}
int tfhe_b200_probe_fp64_tflops(tfhe_b200_ctx* ctx, double* tflops) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(tflops, "probe: null output");
    CU(cudaSetDevice(ctx->device));
    CU(probe_fp64(tflops));
    return TFHE_B200_OK;
}
int tfhe_b200_probe_real96_gprods(tfhe_b200_ctx* ctx, double* gprods) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(gprods, "probe: null output");
    CU(cudaSetDevice(ctx->device));
    CU(probe_real96(gprods));
    return TFHE_B200_OK;
}
int tfhe_b200_probe_read_gbs(tfhe_b200_ctx* ctx, size_t bytes, int passes, double* gbs) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    NEED(gbs && bytes >= 4096 && passes >= 1, "probe: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(probe_read(bytes, passes, gbs));
    return TFHE_B200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
