"""Key-switch only timing (development tool): gate lweKeySwitch on random inputs."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mod = importlib.import_module("experimental-tfhe_b200")
eng = mod.Engine(0)
g = O.GateOracle(42)
eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
u = torch.randint(-2**31, 2**31 - 1, (B, g.N + 1), dtype=torch.int64, device="cuda").to(torch.int32)
out = torch.empty((B, g.n + 1), dtype=torch.int32, device="cuda")
eng.lweKeySwitch(out, u, B); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): eng.lweKeySwitch(out, u, B)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps
adds = B * g.N * g.params.ks_t * 512 * 0.75
print(f"B={B} keyswitch {t:.2f} ms  {B/t*1e3:.0f}/s  {adds/t/1e9:.2f} Tadd/s")
