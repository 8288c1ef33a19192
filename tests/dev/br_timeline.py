"""Per-phase cycle counts of one CMUX, measured by lane 0 of every warp of CTA 0 (development tool).

Needs the instrumented private build:  nvcc ... -DBR_TIMELINE -c csrc/br_kernels.cu ; link as experimental-tfhe_b200/build/lib_timeline.so
(see profiles/r1_notes.md).  The product library carries no instrumentation.
"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O

mod = importlib.import_module("experimental-tfhe_b200")
mod.LIB_PATH = os.path.join(ROOT, "experimental-tfhe_b200", "build", "lib_timeline.so")
eng = mod.Engine(0)
g = O.GateOracle(42)
eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 24 * 2
ca = torch.randint(-2**31, 2**31 - 1, (B, g.n + 1), dtype=torch.int64, device="cuda").to(torch.int32)
u = torch.empty((B, g.N + 1), dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.tfhe_bootstrap_woKS_FFT(u, g.MU, ca, B)
torch.cuda.synchronize()
lib = ctypes.CDLL(mod.LIB_PATH)
buf = (ctypes.c_longlong * 1024)()
assert lib.tfhe_b200_dev_timeline(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(32, 32)
names = {0: "loop/other", 1: "decompose / stash load", 2: "fwd pass A (depths 0-3)", 3: "fwd transpose", 4: "fwd pass B (depths 4-7)", 5: "key LDG issue",
         6: "fwd exchange + depth 8", 7: "MAC q=0", 8: "MAC q=1", 9: "load spectral acc (TMEM)", 10: "inv depth 8 + exchange", 11: "inv pass B",
         12: "inv transpose", 13: "inv pass A", 14: "torus convert + ACC update", 15: "final sync", 16: "KeyPipe acquire (wait full)", 17: "KeyPipe release (+ copy issue by last warp)", 18: "wait::st"}
ncmux = 500.0 * 1     # nonzero bara almost always; one bootstrap per warp in the last launch
nw = int(os.environ.get("TL_WARPS", "12"))
w = a[:nw].mean(axis=0) / ncmux
tot = w.sum()
print(f"cycles per CMUX per warp (mean over the warps of CTA 0): {tot:.0f}")
for k in range(19):
    print(f"  {k:2d} {names[k]:32s} {w[k]:8.0f}  {w[k] / tot * 100:5.1f}%")
