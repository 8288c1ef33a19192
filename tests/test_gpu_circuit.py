"""BASELINE configs[2]: 32-bit ripple-carry adder circuits on the batched gate engine (tfhe_b200_circuit_eval_batch).
The reference has no circuit layer; the check is the decrypted result against plain integer arithmetic."""
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
mod = importlib.import_module("experimental-tfhe_b200")
G = mod.GATES


def adder_netlist(bits):
    """same wire map as tfhe_b200_compat::ripple_carry_adder_netlist (host/tfhe_b200_compat.hpp)"""
    a0, b0, cin, s0 = 0, bits, 2 * bits, 2 * bits + 1
    t0, g0, p0 = s0 + bits, s0 + 2 * bits, s0 + 3 * bits
    c0 = p0 + bits
    gates = [(G["XOR"], t0 + i, a0 + i, b0 + i, 0) for i in range(bits)]
    gates += [(G["AND"], g0 + i, a0 + i, b0 + i, 0) for i in range(bits)]
    gates.append((mod.OP_COPY, c0, cin, 0, 0))
    for i in range(bits):
        gates += [(G["XOR"], s0 + i, t0 + i, c0 + i, 0), (G["AND"], p0 + i, t0 + i, c0 + i, 0), (G["OR"], c0 + i + 1, g0 + i, p0 + i, 0)]
    return np.array(gates, np.int32), dict(a0=a0, b0=b0, cin=cin, s0=s0, c0=c0, n_wires=c0 + bits + 1)


def test_ripple_carry_adder32(gate_engine, gate_oracle):
    g = gate_oracle
    bits, B = 32, 48
    rng = np.random.default_rng(77)
    A = rng.integers(0, 2**32, size=B, dtype=np.uint64); Bv = rng.integers(0, 2**32, size=B, dtype=np.uint64)
    A[0], Bv[0] = 2**32 - 1, 1                      # full carry propagation
    A[1], Bv[1] = 0, 0
    cin = rng.integers(0, 2, size=B)
    gates, w = adder_netlist(bits)
    wires = torch.zeros((w["n_wires"], B, g.n + 1), dtype=torch.int32, device=DEV)
    for i in range(bits):
        wires[w["a0"] + i] = torch.from_numpy(g.encrypt_bits((A >> np.uint64(i)) & np.uint64(1), 1000 + i)).to(DEV)
        wires[w["b0"] + i] = torch.from_numpy(g.encrypt_bits((Bv >> np.uint64(i)) & np.uint64(1), 2000 + i)).to(DEV)
    wires[w["cin"]] = torch.from_numpy(g.encrypt_bits(cin, 3000)).to(DEV)
    gate_engine.circuit_eval(gates, wires, w["n_wires"], B)
    torch.cuda.synchronize()
    res = wires.cpu().numpy()
    total = np.zeros(B, np.uint64)
    for i in range(bits):
        total |= g.decrypt_bits(res[w["s0"] + i]).astype(np.uint64) << np.uint64(i)
    cout = g.decrypt_bits(res[w["c0"] + bits]).astype(np.uint64)
    expect = A + Bv + cin.astype(np.uint64)
    assert np.array_equal(total, expect & np.uint64(2**32 - 1))
    assert np.array_equal(cout, expect >> np.uint64(32))


def test_netlist_semantics_not_mux_inplace(gate_engine, gate_oracle):
    """NOT, MUX, COPY, an output wire that is also an input, and a run that must NOT be merged (gate 2 reads gate 1's output)."""
    g = gate_oracle
    B = 40
    rng = np.random.default_rng(5)
    x = [rng.integers(0, 2, size=B) for _ in range(4)]
    gates = np.array([
        (mod.OP_NOT, 4, 0, 0, 0),            # w4 = !x0
        (G["AND"], 5, 0, 1, 0),              # w5 = x0 & x1
        (G["AND"], 6, 1, 5, 0),              # w6 = x1 & w5   (same op, wires +1, but reads the previous output: no merge)
        (mod.OP_MUX, 7, 2, 3, 4, ),          # w7 = x2 ? x3 : w4
        (G["XOR"], 7, 7, 6, 0),              # w7 ^= w6       (in place)
        (mod.OP_COPY, 8, 7, 0, 0),
    ], np.int32)
    wires = torch.zeros((9, B, g.n + 1), dtype=torch.int32, device=DEV)
    for i in range(4):
        wires[i] = torch.from_numpy(g.encrypt_bits(x[i], 50 + i)).to(DEV)
    gate_engine.circuit_eval(gates, wires, 9, B)
    torch.cuda.synchronize()
    res = wires.cpu().numpy()
    w5 = x[0] & x[1]; w6 = x[1] & w5; w7 = np.where(x[2] == 1, x[3], 1 - x[0]) ^ w6
    assert np.array_equal(g.decrypt_bits(res[4]), 1 - x[0])
    assert np.array_equal(g.decrypt_bits(res[6]), w6)
    assert np.array_equal(g.decrypt_bits(res[8]), w7)


def test_circuit_eval_errors(gate_engine, gate_oracle):
    g = gate_oracle
    wires = torch.zeros((2, 4, g.n + 1), dtype=torch.int32, device=DEV)
    with pytest.raises(mod.EngineError):
        gate_engine.circuit_eval(np.array([(G["AND"], 2, 0, 1, 0)], np.int32), wires, 2, 4)      # wire out of range
    with pytest.raises(mod.EngineError):
        gate_engine.circuit_eval(np.array([(12, 1, 0, 1, 0)], np.int32), wires, 2, 4)            # unknown op
    gate_engine.circuit_eval(np.zeros((0, 5), np.int32), wires, 2, 4)                             # empty netlist is fine


def test_circuit_in_a_cuda_graph(gate_engine, gate_oracle):
    """After a warm-up call nothing is allocated or synchronised inside tfhe_b200_circuit_eval_batch, so a whole circuit can be captured
    into a CUDA graph and replayed ("each circuit level is one batched gate launch (CUDA graph)")."""
    g = gate_oracle
    bits, B = 4, 24
    rng = np.random.default_rng(3)
    A = rng.integers(0, 2**bits, size=B, dtype=np.uint64); Bv = rng.integers(0, 2**bits, size=B, dtype=np.uint64)
    gates, w = adder_netlist(bits)
    wires = torch.zeros((w["n_wires"], B, g.n + 1), dtype=torch.int32, device=DEV)

    def load_inputs(Av, Bw, seed):
        for i in range(bits):
            wires[w["a0"] + i] = torch.from_numpy(g.encrypt_bits((Av >> np.uint64(i)) & np.uint64(1), seed + i)).to(DEV)
            wires[w["b0"] + i] = torch.from_numpy(g.encrypt_bits((Bw >> np.uint64(i)) & np.uint64(1), seed + 100 + i)).to(DEV)
        wires[w["cin"]] = torch.from_numpy(g.encrypt_bits(np.zeros(B, np.int64), seed + 200)).to(DEV)

    def read_sum():
        res = wires.cpu().numpy()
        total = np.zeros(B, np.uint64)
        for i in range(bits):
            total |= g.decrypt_bits(res[w["s0"] + i]).astype(np.uint64) << np.uint64(i)
        return total | (g.decrypt_bits(res[w["c0"] + bits]).astype(np.uint64) << np.uint64(bits))

    load_inputs(A, Bv, 500)
    gate_engine.circuit_eval(gates, wires, w["n_wires"], B)           # warm-up: sizes the scratch
    torch.cuda.synchronize()
    assert np.array_equal(read_sum(), A + Bv)
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(graph, stream=s):
        gate_engine.circuit_eval(gates, wires, w["n_wires"], B, stream=torch.cuda.current_stream().cuda_stream)
    A2 = rng.integers(0, 2**bits, size=B, dtype=np.uint64); B2 = rng.integers(0, 2**bits, size=B, dtype=np.uint64)
    load_inputs(A2, B2, 900)                                          # new inputs in the captured buffers
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(read_sum(), A2 + B2)
