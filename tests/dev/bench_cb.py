"""Device-side timings for BASELINE.json configs 4 and 5 (development tool; the contract bench is bench.py).

config 4: 4,096 tfhe_CircuitBootstrapFFT at the reference's active parameter set (cb/poc_CircuitBootstrapping.cpp:70-85)
config 5: 128-bit fixed-point anticyclic FFT, N = 2048 / 4096, batch 16,384 (hp/code.cpp)
"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O

mod = importlib.import_module("experimental-tfhe_b200")
eng = mod.Engine(0)
out = {}

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

MODE = sys.argv[2] if len(sys.argv) > 2 else ""
# ---------------- config 5: hp FFT
for N in ((2048, 4096) if MODE != "nohp" else ()):
    B = 16384
    x = torch.randint(-2**63, 2**63 - 1, (B, N), dtype=torch.int64, device="cuda")
    spec = torch.empty((B, N // 2, 4), dtype=torch.int64, device="cuda")
    back = torch.empty((B, N), dtype=torch.int64, device="cuda")
    t_i = timeit(lambda: eng.hp_iFFT(spec, x, N, B))
    t_f = timeit(lambda: eng.hp_FFT(back, spec, N, B))
    err = (back - x).abs().max().item()
    out[f"hp_fft_N{N}"] = {"batch": B, "iFFT_ms": t_i, "FFT_ms": t_f, "iFFT_per_s": B / t_i * 1e3, "FFT_per_s": B / t_f * 1e3,
                           "us_per_iFFT": t_i * 1e3 / B, "roundtrip_max_err_lsb": err}
    print(json.dumps({f"hp_fft_N{N}": out[f"hp_fft_N{N}"]}), flush=True)

if MODE == "hponly":
    sys.exit(0)
# ---------------- config 4: circuit bootstrap
t0 = time.time()
c = O.CBOracle(seed=42, with_privks=True)
print(f"oracle keygen {time.time() - t0:.1f}s", flush=True)
eng.load_cb_keys(c.engine_params(), c.preKS, c.bk, c.privKS)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.randint(-2**31, 2**31 - 1, (B, c.N1 + 1), dtype=torch.int64, device="cuda").to(torch.int32)
ell1 = c.params.ell_lvl1
res = torch.empty((B, 2, ell1, 2, c.N1), dtype=torch.int32, device="cuda")
eng.profile_enable(True)
t_cb = timeit(lambda: eng.tfhe_CircuitBootstrapFFT(res, x, B), reps=2)
ms, n = eng.profile_read()
eng.profile_enable(False)
tot = sum(ms.values())
out["circuit_bootstrap"] = {"batch": B, "ms": t_cb, "cb_per_s": B / t_cb * 1e3, "kernel_ms_share": {k: v / tot for k, v in ms.items()},
                            "blind_rotate_ms": ms["blind_rotate"] / 3, "keyswitch_ms": ms["keyswitch"] / 3,
                            "fp64_tflops_blind_rotate": 352.3e6 * 2 * B / (ms["blind_rotate"] / 3 * 1e-3) / 1e12}
print(json.dumps({"circuit_bootstrap": out["circuit_bootstrap"]}), flush=True)
# ---------------- config 4 continued: the TRGSW outputs feed a vertical-packing LUT (8 selector bits per look-up, 256-entry table)
nsel = 8
L = B // nsel
selfft = torch.empty((L * nsel, 2 * ell1, 2, c.N1), dtype=torch.float64, device="cuda")
table = torch.randint(-2**31, 2**31 - 1, (1 << nsel, c.N1), dtype=torch.int64, device="cuda").to(torch.int32)
lut = torch.empty((L, 2, c.N1), dtype=torch.int32, device="cuda")
t_conv = timeit(lambda: eng.tGswToFFTConvert(selfft, res, ell1, L * nsel))
t_lut = timeit(lambda: eng.LUT_vertical_packing(lut, selfft, nsel, table, ell1, c.params.bgbit_lvl1, L))
out["lut_vertical_packing"] = {"lookups": L, "selector_bits": nsel, "tGswToFFTConvert_ms": t_conv, "lut_ms": t_lut,
                               "cmux_per_s": L * ((1 << nsel) - 1) / t_lut * 1e3, "lookups_per_s": L / t_lut * 1e3}
print(json.dumps({"lut_vertical_packing": out["lut_vertical_packing"]}), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_cb_hp.json"), "w"), indent=1)
