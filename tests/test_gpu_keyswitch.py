"""Key switching (lweKeySwitch cb/lwe_functions.cpp:136-171, preKeySwitch cb/poc_CircuitBootstrapping.cpp:437-465, circuitPrivKS
:667-698) is pure integer work: every packing / kernel must reproduce the oracle BIT FOR BIT.

The default path is the tensor-core kernel (csrc/ks_tc_kernels.cu: one-hot digits x byte planes of the key, 128-sample tiles);
the CUDA-core kernels (csrc/ks_kernels.cu: rows, and the paired base-16 form of a base-4 key) are selected per process with
TFHE_B200_KS / TFHE_B200_KS_PAIR and are exercised here in subprocesses.  Cases: tile boundaries (127 / 128 / 129 / 300 samples),
extreme coefficients (all digits zero / maximal, INT_MIN / INT_MAX), the full BASELINE batch (65,536, checked on a random sample of
rows plus a duplicate-rows property), bases 2 / 4 / 8 (other parameter sets), the private key switch across a tile boundary."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _ks(eng, g, x):
    out = torch.empty((len(x), g.n + 1), dtype=torch.int32, device=DEV)
    eng.lweKeySwitch(out, dev(x), len(x))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_tile_boundaries_and_extreme_coefficients(gate_engine, gate_oracle):
    g = gate_oracle
    rng = np.random.default_rng(21)
    for B in (127, 128, 129, 300):
        x = rng.integers(-2**31, 2**31 - 1, size=(B, g.N + 1), dtype=np.int64).astype(np.int32)
        # rows 0..5: all digits zero, all digits maximal, the two ends of the torus, the rounding edge of prec_offset on either side
        half = 1 << (32 - (1 + g.params.ks_basebit * g.params.ks_t))
        for r, v in enumerate((0, -1, -2**31, 2**31 - 1, -half, -half - 1)):
            x[r, : g.N] = v
        assert np.array_equal(_ks(gate_engine, g, x), g.keyswitch(x)), f"B={B}"


def test_full_batch_sampled_rows_and_duplicates(gate_engine, gate_oracle):
    """BASELINE batch: 65,536 samples.  A random sample of rows against the oracle; duplicated input rows (placed in different
    tiles) must give identical outputs."""
    g = gate_oracle
    B = 65536
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randint(-2**31, 2**31 - 1, (B, g.N + 1), dtype=torch.int64, device=DEV, generator=gen).to(torch.int32)
    x[40000] = x[3]; x[65535] = x[3]; x[128] = x[127]
    out = torch.empty((B, g.n + 1), dtype=torch.int32, device=DEV)
    gate_engine.lweKeySwitch(out, x, B)
    torch.cuda.synchronize()
    assert torch.equal(out[40000], out[3]) and torch.equal(out[65535], out[3]) and torch.equal(out[128], out[127])
    rows = np.concatenate([[0, 127, 128, 65535], np.random.default_rng(6).integers(0, B, size=60)])
    ref = g.keyswitch(x[torch.from_numpy(rows).to(DEV)].cpu().numpy())
    assert np.array_equal(out[torch.from_numpy(rows).to(DEV)].cpu().numpy(), ref)


def test_privks_across_a_tile_boundary(engine, cb_oracle):
    c = cb_oracle
    engine.load_cb_keys(c.engine_params(), c.preKS, c.bk, c.privKS)
    rng = np.random.default_rng(8)
    B = 131
    x = rng.integers(-2**63, 2**63 - 1, size=(B, c.N2 + 1), dtype=np.int64)
    x[0] = 0; x[1] = -1; x[2] = -2**63; x[3] = 2**63 - 1
    check = [0, 1, 2, 3, 64, 127, 128, 130]
    for u in (0, 1):
        out = torch.empty((B, 2, c.N1), dtype=torch.int32, device=DEV)
        engine.circuitPrivKS(out, u, dev(x), B)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy()[check], c.circuitPrivKS(u, x[check])), f"u={u}"


_CHILD = r"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle_lib as O
mod = importlib.import_module("experimental-tfhe_b200")
rng = np.random.default_rng(17)
for kw in (dict(), dict(n=320, bk_l=3, bk_Bgbit=8, ks_t=5, ks_basebit=3), dict(n=64, bk_l=1, bk_Bgbit=10, ks_t=16, ks_basebit=1),
           dict(n=100, ks_t=7, ks_basebit=2)):
    g = O.GateOracle(seed=7, **kw)
    eng = mod.Engine(0)
    eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
    for B in (1, 70, 129):
        x = rng.integers(-2**31, 2**31 - 1, size=(B, g.N + 1), dtype=np.int64).astype(np.int32)
        out = torch.empty((B, g.n + 1), dtype=torch.int32, device="cuda:0")
        eng.lweKeySwitch(out, torch.from_numpy(x).cuda(), B)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g.keyswitch(x)), (kw, B)
    blob = eng.export_gate_keys()
    eng2 = mod.Engine(0)
    eng2.import_gate_keys(blob)
    out2 = torch.empty((B, g.n + 1), dtype=torch.int32, device="cuda:0")
    eng2.lweKeySwitch(out2, torch.from_numpy(x).cuda(), B)
    torch.cuda.synchronize()
    assert np.array_equal(out2.cpu().numpy(), g.keyswitch(x)), ("after export/import", kw)
    del eng, eng2
print("ok")
"""


@pytest.mark.parametrize("env", [dict(), dict(TFHE_B200_KS="cuda"), dict(TFHE_B200_KS="cuda", TFHE_B200_KS_PAIR="0")],
                         ids=["tensor-core", "cuda-cores-paired", "cuda-cores-rows"])
def test_every_packing_every_base(env):
    """The packing is chosen once per process, so each one runs in its own interpreter: bases 4 (t = 8 paired / t = 7 unpaired),
    8 and 2, one and several tiles, and the key wire format in that packing."""
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", _CHILD % {"root": ROOT}], env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
