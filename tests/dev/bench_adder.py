"""BASELINE configs[2]: batch of 32-bit ripple-carry adders through tfhe_b200_circuit_eval_batch (development timing tool).
Usage: python tests/dev/bench_adder.py [adders]      prints adders/s and gates/s, checks every decrypted sum."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
from test_gpu_circuit import adder_netlist

mod = importlib.import_module("experimental-tfhe_b200")
eng = mod.Engine(0)
g = O.GateOracle(42)
eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bits = 32
rng = np.random.default_rng(1)
A = rng.integers(0, 2**32, size=B, dtype=np.uint64); Bv = rng.integers(0, 2**32, size=B, dtype=np.uint64)
gates, w = adder_netlist(bits)
wires = torch.zeros((w["n_wires"], B, g.n + 1), dtype=torch.int32, device="cuda")
# one encryption per bit value, replicated (encrypting 2 * 32 * B samples on the host would dominate the run time)
enc = {v: torch.from_numpy(g.encrypt_bits(np.full(64, v), 10 + v)).cuda() for v in (0, 1)}
for i in range(bits):
    for bus, val in ((w["a0"], A), (w["b0"], Bv)):
        bit = torch.from_numpy(((val >> np.uint64(i)) & np.uint64(1)).astype(np.int64)).cuda()
        idx = torch.arange(B, device="cuda") % 64
        wires[bus + i] = torch.where(bit[:, None] == 1, enc[1][idx], enc[0][idx])
wires[w["cin"]] = enc[0][torch.arange(B, device="cuda") % 64]
saved = wires.clone()
eng.circuit_eval(gates, wires, w["n_wires"], B); torch.cuda.synchronize()      # warm-up (scratch allocation)
wires.copy_(saved)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.circuit_eval(gates, wires, w["n_wires"], B); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
res = wires[w["s0"]: w["s0"] + bits].cpu().numpy()
total = np.zeros(B, np.uint64)
for i in range(bits):
    total |= g.decrypt_bits(res[i]).astype(np.uint64) << np.uint64(i)
ok = bool(np.array_equal(total, (A + Bv) & np.uint64(2**32 - 1)))
print(f"{B} x 32-bit adders: {ms:.1f} ms  {B / ms * 1e3:.0f} adders/s  {160 * B / ms * 1e3:.0f} bootstrapped gates/s  sums correct: {ok}")
