"""Quick device-side timings of the gate path kernels (development tool, not the contract bench)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mod = importlib.import_module("experimental-tfhe_b200")
eng = mod.Engine(0)
g = O.GateOracle(42)
eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
gen = torch.Generator(device="cuda").manual_seed(44)
ca = torch.randint(-2**31, 2**31 - 1, (B, g.n + 1), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
cb = torch.randint(-2**31, 2**31 - 1, (B, g.n + 1), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
u = torch.empty((B, g.N + 1), dtype=torch.int32, device="cuda")
out = torch.empty((B, g.n + 1), dtype=torch.int32, device="cuda")

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_br = timeit(lambda: eng.tfhe_bootstrap_woKS_FFT(u, g.MU, ca, B))
t_ks = timeit(lambda: eng.lweKeySwitch(out, u, B))
t_gate = timeit(lambda: eng.bootsGate("NAND", out, ca, cb, B))
print(f"B={B}  blind-rotate {t_br:.2f} ms ({B/t_br*1e3:.0f}/s)  keyswitch {t_ks:.2f} ms ({B/t_ks*1e3:.0f}/s)  NAND {t_gate:.2f} ms ({B/t_gate*1e3:.0f} gates/s)")
flop = 94.72e6 * B
print(f"  FP64 algorithmic: {flop/t_br/1e9:.2f} TFLOP/s in blind rotation")
