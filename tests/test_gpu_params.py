"""Parity at parameter sets other than P_gate ("make all of these runtime parameters").

Set A: n=320, l=3, Bgbit=8, key switch t=5 / basebit=3   -- three gadget levels (the tensor-memory stash serves two of them),
                                                            base-8 key switch on Torus32 inputs (rows read straight from the ring)
Set B: n=64,  l=1, Bgbit=10, key switch t=16 / basebit=1  -- single level (no stash), base-2 key switch
Integer stages bit-exact, one CMUX within 1 LSB of the exact integer external product, gates decrypt (set A).
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SETS = {
    "A": dict(n=320, bk_l=3, bk_Bgbit=8, ks_t=5, ks_basebit=3),
    "B": dict(n=64, bk_l=1, bk_Bgbit=10, ks_t=16, ks_basebit=1),
}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module", params=sorted(SETS))
def setup(request):
    g = O.GateOracle(seed=7, **SETS[request.param])
    mod = importlib.import_module("experimental-tfhe_b200")
    eng = mod.Engine(0)                     # its own context: the session engine keeps the P_gate keys
    eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
    return request.param, g, eng


def test_keyswitch_bit_exact_other_bases(setup):
    _, g, eng = setup
    rng = np.random.default_rng(3)
    for B in (1, 45):
        x = rng.integers(-2**31, 2**31 - 1, size=(B, g.N + 1), dtype=np.int64).astype(np.int32)
        out = torch.empty((B, g.n + 1), dtype=torch.int32, device=DEV)
        eng.lweKeySwitch(out, dev(x), B)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g.keyswitch(x))


def test_single_cmux_vs_exact_other_levels(setup):
    _, g, eng = setup
    rng = np.random.default_rng(11)
    B = 5
    acc = rng.integers(-2**31, 2**31 - 1, size=(B, 2, g.N), dtype=np.int64).astype(np.int32)
    bara = np.zeros((B, g.n), np.int32)
    steps = [0, 1, g.n // 2, g.n - 2, g.n - 1]
    amounts = [1, 1023, 1024, 2047, 600]
    for b in range(B):
        bara[b, steps[b]] = amounts[b]
    dacc = dev(acc)
    eng.tfhe_blindRotate_FFT(dacc, dev(bara), B)
    torch.cuda.synchronize()
    got = dacc.cpu().numpy()
    for b in range(B):
        tmp = np.empty((2, g.N), np.int32)
        for q in range(2):
            O.lib().orc_torusPolynomialMulByXaiMinusOne(O.p(tmp[q]), amounts[b], O.p(acc[b, q]), g.N)
        O.lib().orc_tGswExternMulToTLwe(O.p(tmp), O.p(np.ascontiguousarray(g.bk[steps[b]])), g.N, g.l, g.Bgbit)
        exact = tmp.astype(np.int64) + acc[b].astype(np.int64)
        diff = (got[b].astype(np.int64) - exact + 2**31) % 2**32 - 2**31
        assert np.abs(diff).max() <= 1, f"sample {b}: {np.abs(diff).max()} LSB"


def test_gates_decrypt_set_A(setup):
    name, g, eng = setup
    if name != "A":
        pytest.skip("set B (one gadget level) is not a decryptable parameter set; its stages are checked above")
    rng = np.random.default_rng(9)
    B = 96
    a = rng.integers(0, 2, size=B); b = rng.integers(0, 2, size=B)
    ca, cb = g.encrypt_bits(a, 21), g.encrypt_bits(b, 22)
    out = torch.empty((B, g.n + 1), dtype=torch.int32, device=DEV)
    for op, fn in (("NAND", lambda x, y: 1 - (x & y)), ("XOR", lambda x, y: x ^ y), ("OR", lambda x, y: x | y)):
        eng.bootsGate(op, out, dev(ca), dev(cb), B)
        torch.cuda.synchronize()
        res = out.cpu().numpy()
        assert np.array_equal(g.decrypt_bits(res), fn(a, b)), op
        assert np.array_equal(g.decrypt_bits(g.bootsGate(op, ca[:8], cb[:8])), fn(a, b)[:8]), f"oracle {op}"
