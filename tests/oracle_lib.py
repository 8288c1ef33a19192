"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

GATES = ["NAND", "AND", "OR", "NOR", "XOR", "XNOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]


def build():
    r = subprocess.run(["make", "-C", ORACLE_DIR], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building liboracle.so failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return LIB


class GateParams(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("N", ctypes.c_int), ("k", ctypes.c_int), ("bk_l", ctypes.c_int), ("bk_Bgbit", ctypes.c_int),
                ("ks_t", ctypes.c_int), ("ks_basebit", ctypes.c_int), ("bk_stdev", ctypes.c_double), ("ks_stdev", ctypes.c_double)]


class GateKeys(ctypes.Structure):
    _fields_ = [("p", GateParams), ("lwe_key", ctypes.POINTER(ctypes.c_int32)), ("tlwe_key", ctypes.POINTER(ctypes.c_int32)),
                ("bk", ctypes.POINTER(ctypes.c_int32)), ("bkFFT", ctypes.POINTER(ctypes.c_double)), ("ks", ctypes.POINTER(ctypes.c_int32))]


class CBParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_lvl0", "N_lvl1", "N_lvl2", "bgbit_lvl1", "ell_lvl1", "bgbit_lvl2", "ell_lvl2",
                                            "kslength_lvl10", "ksbasebit_lvl10", "kslength_lvl21", "ksbasebit_lvl21")] + \
               [(n, ctypes.c_double) for n in ("bkstdev_lvl2", "ksstdev_lvl10", "ksstdev_lvl21")]


class CBKeys(ctypes.Structure):
    _fields_ = [("p", CBParams), ("key_lvl0", ctypes.POINTER(ctypes.c_int32)), ("key_lvl1", ctypes.POINTER(ctypes.c_int32)),
                ("key_lvl2", ctypes.POINTER(ctypes.c_int32)), ("preKS", ctypes.POINTER(ctypes.c_int32)),
                ("bk", ctypes.POINTER(ctypes.c_int64)), ("bkFFT", ctypes.POINTER(ctypes.c_double)),
                ("privKS", ctypes.POINTER(ctypes.c_int32))]


class Rng(ctypes.Structure):
    _fields_ = [("s", ctypes.c_uint64), ("has_spare", ctypes.c_int), ("spare", ctypes.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
        L = _lib
        L.orc_gate_keygen.restype = ctypes.POINTER(GateKeys)
        L.orc_gate_keygen.argtypes = [ctypes.POINTER(GateParams), ctypes.c_uint64]
        L.orc_cb_keygen.restype = ctypes.POINTER(CBKeys)
        L.orc_cb_keygen.argtypes = [ctypes.POINTER(CBParams), ctypes.c_uint64, ctypes.c_int]
        L.orc_rng_u64.restype = ctypes.c_uint64
        L.orc_rng_normal.restype = ctypes.c_double
        L.orc_lwePhase.restype = ctypes.c_int32
        L.orc_modSwitchToTorus32.restype = ctypes.c_int32
        L.orc_lwe64Phase_lvl2.restype = ctypes.c_int64
        L.orc_tgsw32_offset.restype = ctypes.c_uint32
        L.orc_tgsw64_offset.restype = ctypes.c_uint64
        L.orc_double_to_torus32.restype = ctypes.c_int32
        L.orc_double_to_torus32.argtypes = [ctypes.c_double]
        L.orc_double_to_torus64.restype = ctypes.c_int64
        L.orc_double_to_torus64.argtypes = [ctypes.c_double]
        L.orc_lweSymEncrypt.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_lwe32Encrypt_lvl1.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_circuitBootstrapWoKS.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_tfhe_bootstrap_woKS_FFT.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
        L.orc_tfhe_bootstrap_FFT.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
    return _lib


def p(a):
    """numpy array -> void*"""
    return a.ctypes.data_as(ctypes.c_void_p)


def view(ptr, shape, dtype):
    n = int(np.prod(shape))
    ct = {np.int32: ctypes.c_int32, np.int64: ctypes.c_int64, np.float64: ctypes.c_double}[dtype]
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ct)), shape=(n,))
    return arr.reshape(shape)


def rng(seed):
    r = Rng()
    lib().orc_rng_seed(ctypes.byref(r), ctypes.c_uint64(seed))
    return r


class GateOracle:
    """Gate-bootstrapping keys + the reference-named operations on numpy arrays."""

    def __init__(self, seed=42, **overrides):
        L = lib()
        self.params = GateParams()
        L.orc_gate_params_default(ctypes.byref(self.params))
        for k_, v_ in overrides.items():
            setattr(self.params, k_, v_)
        self.K = L.orc_gate_keygen(ctypes.byref(self.params), ctypes.c_uint64(seed))
        pp = self.params
        self.n, self.N, self.l, self.Bgbit, self.t, self.basebit = pp.n, pp.N, pp.bk_l, pp.bk_Bgbit, pp.ks_t, pp.ks_basebit
        kk = self.K.contents
        self.lwe_key = view(kk.lwe_key, (pp.n,), np.int32)
        self.tlwe_key = view(kk.tlwe_key, (pp.N,), np.int32)
        self.bk = view(kk.bk, (pp.n, 2 * pp.bk_l, 2, pp.N), np.int32)
        self.bkFFT = view(kk.bkFFT, (pp.n, 2 * pp.bk_l, 2, pp.N), np.float64)
        self.ks = view(kk.ks, (pp.N, pp.ks_t, 1 << pp.ks_basebit, pp.n + 1), np.int32)
        self.MU = 1 << 29

    def __del__(self):
        try:
            lib().orc_gate_keys_free(self.K)
        except Exception:
            pass

    def engine_params(self):
        pp = self.params
        return dict(n=pp.n, N=pp.N, k=pp.k, bk_l=pp.bk_l, bk_Bgbit=pp.bk_Bgbit, ks_t=pp.ks_t, ks_basebit=pp.ks_basebit)

    def encrypt_bits(self, bits, seed):
        r = rng(seed)
        out = np.empty((len(bits), self.n + 1), np.int32)
        for i, b in enumerate(bits):
            lib().orc_bootsSymEncrypt(p(out[i]), int(b), self.K, ctypes.byref(r))
        return out

    def phase(self, samples):
        return np.array([lib().orc_lwePhase(p(s), self.K.contents.lwe_key, self.n) for s in np.ascontiguousarray(samples)], np.int32)

    def phase_N(self, samples):
        """phase of extracted LWE(N) samples under the TLWE key"""
        return np.array([lib().orc_lwePhase(p(s), self.K.contents.tlwe_key, self.N) for s in np.ascontiguousarray(samples)], np.int32)

    def decrypt_bits(self, samples):
        return (self.phase(samples) > 0).astype(np.int32)

    def bootsGate(self, op, ca, cb, threads=0):
        op = GATES.index(op) if isinstance(op, str) else op
        ca = np.ascontiguousarray(ca, np.int32); cb = np.ascontiguousarray(cb, np.int32)
        out = np.empty_like(ca)
        lib().orc_bootsGate_batch(p(out), op, p(ca), p(cb), self.K, len(ca), threads)
        return out

    def bootsMUX(self, a, b, c):
        out = np.empty_like(a)
        for i in range(len(a)):
            lib().orc_bootsMUX(p(out[i]), p(a[i]), p(b[i]), p(c[i]), self.K)
        return out

    def bootstrap_woKS(self, mu, x):
        x = np.ascontiguousarray(x, np.int32)
        out = np.empty((len(x), self.N + 1), np.int32)
        lib().orc_tfhe_bootstrap_woKS_FFT_batch(p(out), self.K, ctypes.c_int32(mu), p(x), len(x), 0)
        return out

    def bootstrap(self, mu, x):
        x = np.ascontiguousarray(x, np.int32)
        out = np.empty((len(x), self.n + 1), np.int32)
        for i in range(len(x)):
            lib().orc_tfhe_bootstrap_FFT(p(out[i]), self.K, mu, p(x[i]))
        return out

    def keyswitch(self, samples):
        samples = np.ascontiguousarray(samples, np.int32)
        out = np.empty((len(samples), self.n + 1), np.int32)
        for i in range(len(samples)):
            lib().orc_lweKeySwitch(p(out[i]), self.K.contents.ks, p(samples[i]), self.N, self.n, self.t, self.basebit)
        return out

    def blindRotate(self, accum, bara):
        accum = np.array(accum, np.int32, copy=True); bara = np.ascontiguousarray(bara, np.int32)
        for i in range(len(accum)):
            lib().orc_tfhe_blindRotate_FFT(p(accum[i]), self.K.contents.bkFFT, p(bara[i]), self.n, self.N, self.l, self.Bgbit)
        return accum

    def blindRotateAndExtract(self, v, barb, bara):
        v = np.ascontiguousarray(v, np.int32); bara = np.ascontiguousarray(bara, np.int32)
        out = np.empty((len(bara), self.N + 1), np.int32)
        for i in range(len(bara)):
            lib().orc_tfhe_blindRotateAndExtract_FFT(p(out[i]), p(v), self.K.contents.bkFFT, int(barb[i]), p(bara[i]),
                                                     self.n, self.N, self.l, self.Bgbit)
        return out


class CBOracle:
    """Circuit-bootstrapping keys (cb/poc_CircuitBootstrapping.cpp parameter set) + operations."""

    def __init__(self, seed=42, with_privks=True, **overrides):
        L = lib()
        self.params = CBParams()
        L.orc_cb_params_default(ctypes.byref(self.params))
        for k_, v_ in overrides.items():
            setattr(self.params, k_, v_)
        self.K = L.orc_cb_keygen(ctypes.byref(self.params), ctypes.c_uint64(seed), int(with_privks))
        pp = self.params
        kk = self.K.contents
        self.n0, self.N1, self.N2 = pp.n_lvl0, pp.N_lvl1, pp.N_lvl2
        self.key_lvl0 = view(kk.key_lvl0, (pp.n_lvl0,), np.int32)
        self.key_lvl1 = view(kk.key_lvl1, (pp.N_lvl1,), np.int32)
        self.key_lvl2 = view(kk.key_lvl2, (pp.N_lvl2 + 1,), np.int32)
        self.preKS = view(kk.preKS, (pp.N_lvl1, pp.kslength_lvl10, 1 << pp.ksbasebit_lvl10, pp.n_lvl0 + 1), np.int32)
        self.bk = view(kk.bk, (pp.n_lvl0, 2 * pp.ell_lvl2, 2, pp.N_lvl2), np.int64)
        self.privKS = None
        if with_privks:
            self.privKS = view(kk.privKS, (2, pp.N_lvl2 + 1, pp.kslength_lvl21, 1 << pp.ksbasebit_lvl21, 2, pp.N_lvl1), np.int32)

    def __del__(self):
        try:
            lib().orc_cb_keys_free(self.K)
        except Exception:
            pass

    def engine_params(self):
        pp = self.params
        return {f: getattr(pp, f) for f in ("n_lvl0", "N_lvl1", "N_lvl2", "bgbit_lvl1", "ell_lvl1", "bgbit_lvl2", "ell_lvl2",
                                            "kslength_lvl10", "ksbasebit_lvl10", "kslength_lvl21", "ksbasebit_lvl21")}

    def encrypt_lvl1(self, messages, stdev, seed):
        r = rng(seed)
        out = np.empty((len(messages), self.N1 + 1), np.int32)
        for i, m in enumerate(messages):
            lib().orc_lwe32Encrypt_lvl1(p(out[i]), int(np.int32(m)), stdev, self.K, ctypes.byref(r))
        return out

    def preKeySwitch(self, x):
        x = np.ascontiguousarray(x, np.int32)
        out = np.empty((len(x), self.n0 + 1), np.int32)
        for i in range(len(x)):
            lib().orc_preKeySwitch(p(out[i]), p(x[i]), self.K)
        return out

    def preModSwitch(self, x):
        x = np.ascontiguousarray(x, np.int32)
        out = np.empty((len(x), self.n0 + 1), np.int32)
        for i in range(len(x)):
            lib().orc_preModSwitch(p(out[i]), p(x[i]), self.n0, self.N2)
        return out

    def circuitBootstrapWoKS(self, mu, abar, threads=1):
        abar = np.ascontiguousarray(abar, np.int32)
        out = np.empty((len(abar), self.N2 + 1), np.int64)
        def one(i):
            lib().orc_circuitBootstrapWoKS(p(out[i]), ctypes.c_int64(mu), p(abar[i]), self.K)
        if threads > 1:      # ctypes releases the GIL; the oracle functions keep their scratch on the stack / heap per call
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, range(len(abar))))
        else:
            for i in range(len(abar)):
                one(i)
        return out

    def circuitPrivKS(self, u, x):
        x = np.ascontiguousarray(x, np.int64)
        out = np.empty((len(x), 2, self.N1), np.int32)
        for i in range(len(x)):
            lib().orc_circuitPrivKS(p(out[i]), u, p(x[i]), self.K)
        return out

    def CircuitBootstrapFFT(self, samples):
        samples = np.ascontiguousarray(samples, np.int32)
        out = np.empty((len(samples), 2, self.params.ell_lvl1, 2, self.N1), np.int32)
        for i in range(len(samples)):
            lib().orc_tfhe_CircuitBootstrapFFT(p(out[i]), p(samples[i]), self.K)
        return out

    def phase_lvl2(self, samples):
        return np.array([lib().orc_lwe64Phase_lvl2(p(s), self.K) for s in np.ascontiguousarray(samples, np.int64)], np.int64)

    def tlwe_phase_lvl1(self, tlwe):
        tlwe = np.ascontiguousarray(tlwe, np.int32)
        out = np.empty((self.N1,), np.int32)
        lib().orc_tLwe32Phase_lvl1(p(out), p(tlwe), self.K)
        return out


def hp_tables(N):
    """(powomega, powombar) as uint64 arrays [2N][4] = {re_lo, re_hi, im_lo, im_hi}"""
    n = 2 * N
    om = np.zeros((n, 4), np.uint64); ob = np.zeros((n, 4), np.uint64)
    lib().orc_hp_precomp_iFFT(p(om), n)
    lib().orc_hp_precomp_FFT(p(ob), n)
    return om, ob


def hp_iFFT(inp, N, om):
    inp = np.ascontiguousarray(inp, np.int64)
    out = np.zeros((N // 2, 4), np.uint64)
    lib().orc_hp_iFFT(p(out), p(inp), 2 * N, p(om))
    return out


def hp_FFT(inp, N, ob):
    tmp = np.array(inp, np.uint64, copy=True)
    out = np.zeros((N,), np.int64)
    lib().orc_hp_FFT(p(out), p(tmp), 2 * N, p(ob))
    return out
