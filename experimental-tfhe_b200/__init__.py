"""experimental-tfhe_b200 -- host-side loader for the B200-native TFHE bootstrapping engine.

The product is ``libtfhe_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/tfhe_b200.h``).  This module only binds that ABI with ctypes so tests and ``bench.py``
can drive it with torch device pointers; it contains no arithmetic and NO fallback: if the shared
library is missing or no B200 is present the calls raise.

Reference API mirrored here (names and argument meaning follow the reference, each call takes a batch):
``tfhe_blindRotate_FFT``, ``tfhe_blindRotateAndExtract_FFT``, ``tfhe_bootstrap_woKS_FFT``,
``lweKeySwitch``, ``tfhe_bootstrap_FFT`` (cb/lwe_functions.cpp:163-171,337-446), ``boots*`` gates
(upstream), ``preKeySwitch``/``preModSwitch``/``circuitBootstrapWoKS``/``circuitPrivKS``/
``tfhe_CircuitBootstrapFFT`` (cb/poc_CircuitBootstrapping.cpp:437-873) and the hp ``iFFT``/``FFT``
(hp/code.cpp:391-512).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFHE_B200_LIB") or os.path.join(_HERE, "libtfhe_b200.so")   # override: development builds only

OK = 0
OP_NOT, OP_COPY, OP_MUX = 16, 17, 18
GATES = {"NAND": 0, "AND": 1, "OR": 2, "NOR": 3, "XOR": 4, "XNOR": 5, "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9}


class GateParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n", "N", "k", "bk_l", "bk_Bgbit", "ks_t", "ks_basebit")]


class CBParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n_lvl0", "N_lvl1", "N_lvl2", "bgbit_lvl1", "ell_lvl1", "bgbit_lvl2", "ell_lvl2",
        "kslength_lvl10", "ksbasebit_lvl10", "kslength_lvl21", "ksbasebit_lvl21")]


class EngineError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libtfhe_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise EngineError("building libtfhe_b200.so failed")
    return LIB_PATH


_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_SIGS = {
    "tfhe_b200_ctx_create": [ctypes.POINTER(_P), _I],
    "tfhe_b200_ctx_destroy": [_P],
    "tfhe_b200_sm_count": [_P],
    "tfhe_b200_synchronize": [_P, _P],
    "tfhe_b200_gate_load_keys": [_P, ctypes.POINTER(GateParams), _P, _P],
    "tfhe_b200_gate_alloc_keys": [_P, ctypes.POINTER(GateParams)],
    "tfhe_b200_gate_key_blob": [_P, _I, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_gate_commit_keys": [_P],
    "tfhe_b200_cb_alloc_keys": [_P, ctypes.POINTER(CBParams), _I],
    "tfhe_b200_cb_key_blob": [_P, _I, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_cb_commit_keys": [_P],
    "tfhe_b200_cb_export_keys": [_P, _P, ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_cb_import_keys": [_P, _P, ctypes.c_size_t],
    "tfhe_b200_ciphertext_pack": [_I, _P, _P, _P, ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_ciphertext_unpack": [_P, ctypes.c_size_t, ctypes.POINTER(_I), _P, _P, ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_blindRotate_FFT_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_blindRotateAndExtract_FFT_batch": [_P, _P, _P, _P, _P, _I, _P],
    "tfhe_b200_bootstrap_woKS_FFT_batch": [_P, _P, ctypes.c_int32, _P, _I, _P],
    "tfhe_b200_lweKeySwitch_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_bootstrap_FFT_batch": [_P, _P, ctypes.c_int32, _P, _I, _P],
    "tfhe_b200_bootsGate_batch": [_P, _I, _P, _P, _P, _I, _P],
    "tfhe_b200_bootsNOT_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_bootsMUX_batch": [_P, _P, _P, _P, _P, _I, _P],
    "tfhe_b200_circuit_eval_batch": [_P, _P, _I, _P, _I, _I, _P],
    "tfhe_b200_gate_export_keys": [_P, _P, ctypes.POINTER(ctypes.c_size_t)],
    "tfhe_b200_gate_import_keys": [_P, _P, ctypes.c_size_t],
    "tfhe_b200_tGswToFFTConvert_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_tGswFFTExternMulToTLwe_batch": [_P, _P, _P, _I, _I, _I, _I, _P],
    "tfhe_b200_CMux_batch": [_P, _P, _P, _I, _P, _P, _I, _I, _I, _P],
    "tfhe_b200_LUT_vertical_packing_batch": [_P, _P, _P, _I, _P, _I, _I, _I, _P],
    "tfhe_b200_bootsGate_batch_host": [_P, _I, _P, _P, _P, _I],
    "tfhe_b200_IntPolynomial_ifft_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_TorusPolynomial64_ifft_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_TorusPolynomial_fft_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_TorusPolynomial64_fft_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_LagrangeHalfCPolynomialAddMul_batch": [_P, _P, _P, _P, _I, _I, _P],
    "tfhe_b200_cb_load_keys": [_P, ctypes.POINTER(CBParams), _P, _P, _P],
    "tfhe_b200_preKeySwitch_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_preModSwitch_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_circuitBootstrapWoKS_batch": [_P, _P, ctypes.c_int64, _P, _I, _P],
    "tfhe_b200_circuitPrivKS_batch": [_P, _P, _I, _P, _I, _P],
    "tfhe_b200_CircuitBootstrapFFT_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_CircuitBootstrapFFT_batch_host": [_P, _P, _P, _I],
    "tfhe_b200_hp_iFFT_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_hp_FFT_batch": [_P, _P, _P, _I, _I, _P],
    "tfhe_b200_profile_enable": [_P, _I],
    "tfhe_b200_profile_read": [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)],
    "tfhe_b200_probe_fp64_tflops": [_P, ctypes.POINTER(ctypes.c_double)],
    "tfhe_b200_blindRotate64_FFT_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_cb_load_exact_key": [_P, _P],
    "tfhe_b200_cb_set_exact": [_P, _I],
    "tfhe_b200_blindRotate64_exact_batch": [_P, _P, _P, _I, _P],
    "tfhe_b200_gate_get_params": [_P, ctypes.POINTER(GateParams)],
    "tfhe_b200_gate_keygen": [_P, ctypes.POINTER(GateParams), ctypes.c_double, ctypes.c_double, ctypes.c_uint64, _P, _P, _P, _P],
    "tfhe_b200_cb_keygen": [_P, ctypes.POINTER(CBParams), ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, _P, _P, _P, _I],
    "tfhe_b200_probe_real96_gprods": [_P, ctypes.POINTER(ctypes.c_double)],
    "tfhe_b200_probe_read_gbs": [_P, ctypes.c_size_t, _I, ctypes.POINTER(ctypes.c_double)],
}
EXPORTS = sorted(list(_SIGS) + ["tfhe_b200_last_error"])


def load():
    """dlopen libtfhe_b200.so; raises if it was not built (there is no Python/CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _I
    lib.tfhe_b200_last_error.argtypes = [_P]
    lib.tfhe_b200_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def _ptr(x):
    """Accept torch tensors (device or host), numpy arrays, ints, None."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    return int(x)


CT_KINDS = {"LWE32": 1, "LWE64": 2, "TLWE32": 3, "TGSW32": 4}


def ciphertext_pack(kind, samples):
    """numpy array of samples -> wire blob (tfhe_b200_ciphertext_pack); kind in CT_KINDS"""
    import numpy as np
    lib = load()
    a = np.ascontiguousarray(samples)
    dims = (ctypes.c_int64 * 4)(*(list(a.shape) + [0] * (4 - a.ndim)))
    n = ctypes.c_size_t()
    rc = lib.tfhe_b200_ciphertext_pack(CT_KINDS[kind], dims, None, None, ctypes.byref(n))
    if rc != OK:
        raise EngineError(f"ciphertext_pack failed ({rc}): {lib.tfhe_b200_last_error(None).decode()}")
    buf = np.empty(n.value, np.uint8)
    rc = lib.tfhe_b200_ciphertext_pack(CT_KINDS[kind], dims, _ptr(a), _ptr(buf), ctypes.byref(n))
    if rc != OK:
        raise EngineError(f"ciphertext_pack failed ({rc}): {lib.tfhe_b200_last_error(None).decode()}")
    return buf


def ciphertext_unpack(blob):
    """wire blob -> (kind name, numpy array)"""
    import numpy as np
    lib = load()
    kind = _I(); dims = (ctypes.c_int64 * 4)(); n = ctypes.c_size_t()
    rc = lib.tfhe_b200_ciphertext_unpack(_ptr(blob), int(blob.nbytes), ctypes.byref(kind), dims, None, ctypes.byref(n))
    if rc != OK:
        raise EngineError(f"ciphertext_unpack failed ({rc}): {lib.tfhe_b200_last_error(None).decode()}")
    name = [k for k, v in CT_KINDS.items() if v == kind.value][0]
    shape = [d for d in dims if d > 0]
    out = np.empty(shape, np.int64 if name == "LWE64" else np.int32)
    rc = lib.tfhe_b200_ciphertext_unpack(_ptr(blob), int(blob.nbytes), None, None, _ptr(out), ctypes.byref(n))
    if rc != OK:
        raise EngineError(f"ciphertext_unpack failed ({rc}): {lib.tfhe_b200_last_error(None).decode()}")
    return name, out


class Engine:
    """One context per process and GPU (tfhe_b200_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _P()
        rc = self.lib.tfhe_b200_ctx_create(ctypes.byref(h), int(device))
        if rc != OK:
            raise EngineError(f"tfhe_b200_ctx_create failed ({rc}): {self.lib.tfhe_b200_last_error(None).decode()}")
        self.h = h
        self.device = device
        self.gate_params = None
        self.cb_params = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.tfhe_b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != OK:
            raise EngineError(f"{what} failed ({rc}): {self.lib.tfhe_b200_last_error(self.h).decode()}")

    @staticmethod
    def _stream(stream):
        if stream is None:
            try:
                import torch
                return torch.cuda.current_stream().cuda_stream
            except Exception:
                return None
        return _ptr(stream)

    def sm_count(self):
        return self.lib.tfhe_b200_sm_count(self.h)

    def synchronize(self, stream=None):
        self._ck(self.lib.tfhe_b200_synchronize(self.h, self._stream(stream)), "synchronize")

    # ------------------------------------------------------------------ diagnostics
    def profile_enable(self, on=True):
        self._ck(self.lib.tfhe_b200_profile_enable(self.h, int(bool(on))), "profile_enable")

    def profile_read(self):
        """-> ({'blind_rotate': ms, 'keyswitch': ms, 'other': ms}, {same keys: launches})"""
        ms = (ctypes.c_double * 3)(); n = (ctypes.c_int * 3)()
        self._ck(self.lib.tfhe_b200_profile_read(self.h, ms, n), "profile_read")
        names = ("blind_rotate", "keyswitch", "other")
        return dict(zip(names, list(ms))), dict(zip(names, list(n)))

    def probe_fp64_tflops(self):
        v = ctypes.c_double()
        self._ck(self.lib.tfhe_b200_probe_fp64_tflops(self.h, ctypes.byref(v)), "probe_fp64_tflops")
        return v.value

    def probe_real96_gprods(self):
        v = ctypes.c_double()
        self._ck(self.lib.tfhe_b200_probe_real96_gprods(self.h, ctypes.byref(v)), "probe_real96_gprods")
        return v.value

    def probe_read_gbs(self, nbytes, passes):
        v = ctypes.c_double()
        self._ck(self.lib.tfhe_b200_probe_read_gbs(self.h, nbytes, passes, ctypes.byref(v)), "probe_read_gbs")
        return v.value

    # ------------------------------------------------------------------ gate path
    def load_gate_keys(self, params, bk_host, ks_host):
        p = GateParams(**params) if isinstance(params, dict) else params
        self._ck(self.lib.tfhe_b200_gate_load_keys(self.h, ctypes.byref(p), _ptr(bk_host), _ptr(ks_host)), "gate_load_keys")
        self.gate_params = p

    def gate_keygen(self, params, bk_stdev, ks_stdev, seed, want_raw=False):
        """Keys generated on the device (tfhe_b200_gate_keygen).  Returns (lwe_key, tlwe_key[, bk_raw, ks_raw]) as numpy arrays."""
        import numpy as np
        p = GateParams(**params) if isinstance(params, dict) else params
        lwe = np.zeros(p.n, np.int32); tlwe = np.zeros(p.N, np.int32)
        bk = np.zeros((p.n, 2 * p.bk_l, 2, p.N), np.int32) if want_raw else None
        ks = np.zeros((p.N, p.ks_t, 1 << p.ks_basebit, p.n + 1), np.int32) if want_raw else None
        self._ck(self.lib.tfhe_b200_gate_keygen(self.h, ctypes.byref(p), bk_stdev, ks_stdev, seed, _ptr(lwe), _ptr(tlwe), _ptr(bk), _ptr(ks)), "gate_keygen")
        self.gate_params = p
        return (lwe, tlwe, bk, ks) if want_raw else (lwe, tlwe)

    def cb_keygen(self, params, bkstdev_lvl2, ksstdev_lvl10, ksstdev_lvl21, seed, with_privks=True):
        import numpy as np
        p = CBParams(**params) if isinstance(params, dict) else params
        k0 = np.zeros(p.n_lvl0, np.int32); k1 = np.zeros(p.N_lvl1, np.int32); k2 = np.zeros(p.N_lvl2 + 1, np.int32)
        self._ck(self.lib.tfhe_b200_cb_keygen(self.h, ctypes.byref(p), bkstdev_lvl2, ksstdev_lvl10, ksstdev_lvl21, seed, _ptr(k0), _ptr(k1), _ptr(k2),
                                              int(with_privks)), "cb_keygen")
        self.cb_params = p
        return k0, k1, k2

    def alloc_gate_keys(self, params):
        p = GateParams(**params) if isinstance(params, dict) else params
        self._ck(self.lib.tfhe_b200_gate_alloc_keys(self.h, ctypes.byref(p)), "gate_alloc_keys")
        self.gate_params = p

    def commit_gate_keys(self):
        self._ck(self.lib.tfhe_b200_gate_commit_keys(self.h), "gate_commit_keys")

    def gate_key_blob(self, which):
        ptr, nbytes = _P(), ctypes.c_size_t()
        self._ck(self.lib.tfhe_b200_gate_key_blob(self.h, which, ctypes.byref(ptr), ctypes.byref(nbytes)), "gate_key_blob")
        return ptr.value, nbytes.value

    def tfhe_blindRotate_FFT(self, accum, bara, count, stream=None):
        self._ck(self.lib.tfhe_b200_blindRotate_FFT_batch(self.h, _ptr(accum), _ptr(bara), count, self._stream(stream)), "tfhe_blindRotate_FFT")

    def tfhe_blindRotateAndExtract_FFT(self, result, v, barb, bara, count, stream=None):
        self._ck(self.lib.tfhe_b200_blindRotateAndExtract_FFT_batch(self.h, _ptr(result), _ptr(v), _ptr(barb), _ptr(bara), count,
                                                                    self._stream(stream)), "tfhe_blindRotateAndExtract_FFT")

    def tfhe_bootstrap_woKS_FFT(self, result, mu, x, count, stream=None):
        self._ck(self.lib.tfhe_b200_bootstrap_woKS_FFT_batch(self.h, _ptr(result), mu, _ptr(x), count, self._stream(stream)), "tfhe_bootstrap_woKS_FFT")

    def lweKeySwitch(self, result, sample, count, stream=None):
        self._ck(self.lib.tfhe_b200_lweKeySwitch_batch(self.h, _ptr(result), _ptr(sample), count, self._stream(stream)), "lweKeySwitch")

    def tfhe_bootstrap_FFT(self, result, mu, x, count, stream=None):
        self._ck(self.lib.tfhe_b200_bootstrap_FFT_batch(self.h, _ptr(result), mu, _ptr(x), count, self._stream(stream)), "tfhe_bootstrap_FFT")

    def bootsGate(self, op, result, ca, cb, count, stream=None):
        op = GATES[op] if isinstance(op, str) else op
        self._ck(self.lib.tfhe_b200_bootsGate_batch(self.h, op, _ptr(result), _ptr(ca), _ptr(cb), count, self._stream(stream)), "bootsGate")

    def bootsNAND(self, result, ca, cb, count, stream=None):
        self.bootsGate("NAND", result, ca, cb, count, stream)

    def bootsAND(self, result, ca, cb, count, stream=None):
        self.bootsGate("AND", result, ca, cb, count, stream)

    def bootsOR(self, result, ca, cb, count, stream=None):
        self.bootsGate("OR", result, ca, cb, count, stream)

    def bootsXOR(self, result, ca, cb, count, stream=None):
        self.bootsGate("XOR", result, ca, cb, count, stream)

    def bootsNOT(self, result, ca, count, stream=None):
        self._ck(self.lib.tfhe_b200_bootsNOT_batch(self.h, _ptr(result), _ptr(ca), count, self._stream(stream)), "bootsNOT")

    def bootsMUX(self, result, a, b, c, count, stream=None):
        self._ck(self.lib.tfhe_b200_bootsMUX_batch(self.h, _ptr(result), _ptr(a), _ptr(b), _ptr(c), count, self._stream(stream)), "bootsMUX")

    def export_gate_keys(self):
        """the loaded gate keys in the engine's wire format (bytes-like numpy array)"""
        import numpy as np
        n = ctypes.c_size_t(0)
        self._ck(self.lib.tfhe_b200_gate_export_keys(self.h, None, ctypes.byref(n)), "gate_export_keys")
        buf = np.empty(n.value, np.uint8)
        self._ck(self.lib.tfhe_b200_gate_export_keys(self.h, buf.ctypes.data, ctypes.byref(n)), "gate_export_keys")
        return buf

    def import_gate_keys(self, blob):
        import numpy as np
        b = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob)
        self._ck(self.lib.tfhe_b200_gate_import_keys(self.h, b.ctypes.data, b.size), "gate_import_keys")

    def circuit_eval(self, gates, wires, n_wires, count, stream=None):
        """gates: int32 array [n_gates][5] = (op, out, in0, in1, in2) (tfhe_b200_gate); wires: device tensor [n_wires][count][n+1]."""
        import numpy as np
        g = np.ascontiguousarray(gates, dtype=np.int32).reshape(-1, 5)
        self._ck(self.lib.tfhe_b200_circuit_eval_batch(self.h, g.ctypes.data, len(g), _ptr(wires), n_wires, count, self._stream(stream)), "circuit_eval")

    def bootsGate_host(self, op, result_host, ca_host, cb_host, count):
        op = GATES[op] if isinstance(op, str) else op
        self._ck(self.lib.tfhe_b200_bootsGate_batch_host(self.h, op, _ptr(result_host), _ptr(ca_host), _ptr(cb_host), count), "bootsGate_host")

    # ------------------------------------------------------------------ TRGSW x TRLWE (N = 1024, Torus32)
    def tGswToFFTConvert(self, gswfft, gsw, l, count, stream=None):
        self._ck(self.lib.tfhe_b200_tGswToFFTConvert_batch(self.h, _ptr(gswfft), _ptr(gsw), l, count, self._stream(stream)), "tGswToFFTConvert")

    def tGswFFTExternMulToTLwe(self, accum, gswfft, per_sample, l, Bgbit, count, stream=None):
        self._ck(self.lib.tfhe_b200_tGswFFTExternMulToTLwe_batch(self.h, _ptr(accum), _ptr(gswfft), int(per_sample), l, Bgbit, count,
                                                                 self._stream(stream)), "tGswFFTExternMulToTLwe")

    def CMux(self, result, gswfft, per_sample, d1, d0, l, Bgbit, count, stream=None):
        self._ck(self.lib.tfhe_b200_CMux_batch(self.h, _ptr(result), _ptr(gswfft), int(per_sample), _ptr(d1), _ptr(d0), l, Bgbit, count,
                                               self._stream(stream)), "CMux")

    def LUT_vertical_packing(self, result, selfft, nsel, table, l, Bgbit, count, stream=None):
        self._ck(self.lib.tfhe_b200_LUT_vertical_packing_batch(self.h, _ptr(result), _ptr(selfft), nsel, _ptr(table), l, Bgbit, count,
                                                               self._stream(stream)), "LUT_vertical_packing")

    # ------------------------------------------------------------------ transforms
    def IntPolynomial_ifft(self, result, poly, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_IntPolynomial_ifft_batch(self.h, _ptr(result), _ptr(poly), N, count, self._stream(stream)), "IntPolynomial_ifft")

    def TorusPolynomial64_ifft(self, result, poly, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_TorusPolynomial64_ifft_batch(self.h, _ptr(result), _ptr(poly), N, count, self._stream(stream)), "TorusPolynomial64_ifft")

    def TorusPolynomial_fft(self, result, lagr, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_TorusPolynomial_fft_batch(self.h, _ptr(result), _ptr(lagr), N, count, self._stream(stream)), "TorusPolynomial_fft")

    def TorusPolynomial64_fft(self, result, lagr, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_TorusPolynomial64_fft_batch(self.h, _ptr(result), _ptr(lagr), N, count, self._stream(stream)), "TorusPolynomial64_fft")

    def LagrangeHalfCPolynomialAddMul(self, res, a, b, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_LagrangeHalfCPolynomialAddMul_batch(self.h, _ptr(res), _ptr(a), _ptr(b), N, count, self._stream(stream)),
                 "LagrangeHalfCPolynomialAddMul")

    # ------------------------------------------------------------------ circuit bootstrapping
    def load_cb_keys(self, params, preKS_host, bk_host, privKS_host=None):
        p = CBParams(**params) if isinstance(params, dict) else params
        self._ck(self.lib.tfhe_b200_cb_load_keys(self.h, ctypes.byref(p), _ptr(preKS_host), _ptr(bk_host), _ptr(privKS_host)), "cb_load_keys")
        self.cb_params = p

    def blindRotate64_FFT(self, accum, bara, count, stream=None):
        self._ck(self.lib.tfhe_b200_blindRotate64_FFT_batch(self.h, _ptr(accum), _ptr(bara), count, self._stream(stream)), "blindRotate64_FFT")

    def load_cb_exact_key(self, bk_host):
        self._ck(self.lib.tfhe_b200_cb_load_exact_key(self.h, _ptr(bk_host)), "cb_load_exact_key")

    def set_cb_exact(self, on):
        self._ck(self.lib.tfhe_b200_cb_set_exact(self.h, int(bool(on))), "cb_set_exact")

    def blindRotate64_exact(self, accum, bara, count, stream=None):
        self._ck(self.lib.tfhe_b200_blindRotate64_exact_batch(self.h, _ptr(accum), _ptr(bara), count, self._stream(stream)), "blindRotate64_exact")

    def alloc_cb_keys(self, params, with_privks=True):
        p = CBParams(**params) if isinstance(params, dict) else params
        self._ck(self.lib.tfhe_b200_cb_alloc_keys(self.h, ctypes.byref(p), int(with_privks)), "cb_alloc_keys")
        self.cb_params = p

    def cb_key_blob(self, which):
        ptr, nbytes = _P(), ctypes.c_size_t()
        self._ck(self.lib.tfhe_b200_cb_key_blob(self.h, which, ctypes.byref(ptr), ctypes.byref(nbytes)), "cb_key_blob")
        return ptr.value, nbytes.value

    def commit_cb_keys(self):
        self._ck(self.lib.tfhe_b200_cb_commit_keys(self.h), "cb_commit_keys")

    def export_cb_keys(self):
        """numpy uint8 blob: header + the three device key blobs (tfhe_b200_cb_export_keys)"""
        import numpy as np
        n = ctypes.c_size_t()
        self._ck(self.lib.tfhe_b200_cb_export_keys(self.h, None, ctypes.byref(n)), "cb_export_keys(size)")
        buf = np.empty(n.value, np.uint8)
        self._ck(self.lib.tfhe_b200_cb_export_keys(self.h, _ptr(buf), ctypes.byref(n)), "cb_export_keys")
        return buf

    def import_cb_keys(self, blob):
        self._ck(self.lib.tfhe_b200_cb_import_keys(self.h, _ptr(blob), int(blob.nbytes)), "cb_import_keys")

    def preKeySwitch(self, result, x, count, stream=None):
        self._ck(self.lib.tfhe_b200_preKeySwitch_batch(self.h, _ptr(result), _ptr(x), count, self._stream(stream)), "preKeySwitch")

    def preModSwitch(self, result, x, count, stream=None):
        self._ck(self.lib.tfhe_b200_preModSwitch_batch(self.h, _ptr(result), _ptr(x), count, self._stream(stream)), "preModSwitch")

    def circuitBootstrapWoKS(self, result, mu, abar, count, stream=None):
        self._ck(self.lib.tfhe_b200_circuitBootstrapWoKS_batch(self.h, _ptr(result), mu, _ptr(abar), count, self._stream(stream)), "circuitBootstrapWoKS")

    def circuitPrivKS(self, result, u, x, count, stream=None):
        self._ck(self.lib.tfhe_b200_circuitPrivKS_batch(self.h, _ptr(result), u, _ptr(x), count, self._stream(stream)), "circuitPrivKS")

    def tfhe_CircuitBootstrapFFT(self, result, sample, count, stream=None):
        self._ck(self.lib.tfhe_b200_CircuitBootstrapFFT_batch(self.h, _ptr(result), _ptr(sample), count, self._stream(stream)), "tfhe_CircuitBootstrapFFT")

    def tfhe_CircuitBootstrapFFT_host(self, result_host, sample_host, count):
        self._ck(self.lib.tfhe_b200_CircuitBootstrapFFT_batch_host(self.h, _ptr(result_host), _ptr(sample_host), count), "tfhe_CircuitBootstrapFFT_host")

    # ------------------------------------------------------------------ high-precision FFT
    def hp_iFFT(self, out, inp, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_hp_iFFT_batch(self.h, _ptr(out), _ptr(inp), N, count, self._stream(stream)), "hp_iFFT")

    def hp_FFT(self, out, inp, N, count, stream=None):
        self._ck(self.lib.tfhe_b200_hp_FFT_batch(self.h, _ptr(out), _ptr(inp), N, count, self._stream(stream)), "hp_FFT")
