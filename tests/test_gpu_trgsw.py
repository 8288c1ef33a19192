"""TRGSW x TRLWE products at N = 1024 / Torus32 (BASELINE configs[3] "feeding a vertical-packing LUT").

tGswFFTExternMulToTLwe (cb/tgsw_functions.cpp:424-449) within 1 LSB of the oracle's exact integer external product; CMux and the
vertical-packing LUT have no reference code (the CMux is a commented stub, cb/poc_CircuitBootstrapping.cpp:877-879): they are
checked by what they must select, first with noiseless TGSW samples, then end to end on circuit-bootstrapped selector bits."""
import numpy as np
import pytest
import torch

import oracle_lib as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N = 1024


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def centered(x):
    return (x.astype(np.int64) + 2**31) % 2**32 - 2**31


def noiseless_tgsw(bit, l, Bgbit):
    """TGSW of `bit` with zero masks and zero noise: row (bloc, i) = m * 2^(32-(i+1)Bgbit) at coefficient 0 of polynomial bloc
    (cb/poc_CircuitBootstrapping.cpp:215-227 without the encryptions of zero) -- valid under every key."""
    g = np.zeros((2 * l, 2, N), np.int32)
    for bloc in range(2):
        for i in range(l):
            g[bloc * l + i, bloc, 0] = np.int32(np.uint32((bit << (32 - (i + 1) * Bgbit)) & 0xFFFFFFFF))
    return g


@pytest.mark.parametrize("l,Bgbit", [(2, 8), (2, 10), (3, 8)])
def test_extern_mul_vs_exact(engine, l, Bgbit):
    rng = np.random.default_rng(100 + l * 16 + Bgbit)
    B = 6
    gsw = rng.integers(-2**31, 2**31 - 1, size=(B, 2 * l, 2, N), dtype=np.int64).astype(np.int32)
    acc = rng.integers(-2**31, 2**31 - 1, size=(B, 2, N), dtype=np.int64).astype(np.int32)
    gfft = torch.empty((B, 2 * l, 2, N), dtype=torch.float64, device=DEV)
    engine.tGswToFFTConvert(gfft, dev(gsw), l, B)
    for per_sample in (1, 0):
        dacc = dev(acc)
        engine.tGswFFTExternMulToTLwe(dacc, gfft, per_sample, l, Bgbit, B)
        torch.cuda.synchronize()
        got = dacc.cpu().numpy()
        for b in range(B):
            exact = acc[b].copy()
            O.lib().orc_tGswExternMulToTLwe(O.p(exact), O.p(np.ascontiguousarray(gsw[b if per_sample else 0])), N, l, Bgbit)
            assert np.abs(centered(got[b] - exact)).max() <= 1, f"per_sample={per_sample} sample {b}"


def test_cmux_selects(engine):
    l, Bgbit = 2, 8
    rng = np.random.default_rng(7)
    B = 8
    bits = np.array([0, 1, 1, 0, 1, 0, 0, 1])
    gsw = np.stack([noiseless_tgsw(int(b), l, Bgbit) for b in bits])
    gfft = torch.empty((B, 2 * l, 2, N), dtype=torch.float64, device=DEV)
    engine.tGswToFFTConvert(gfft, dev(gsw), l, B)
    d1 = rng.integers(-2**31, 2**31 - 1, size=(B, 2, N), dtype=np.int64).astype(np.int32)
    d0 = rng.integers(-2**31, 2**31 - 1, size=(B, 2, N), dtype=np.int64).astype(np.int32)
    out = torch.empty((B, 2, N), dtype=torch.int32, device=DEV)
    engine.CMux(out, gfft, 1, dev(d1), dev(d0), l, Bgbit, B)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = np.where(bits[:, None, None] == 1, d1, d0)
    # a noiseless selector reproduces d1 - d0 up to the gadget's precision.  The Torus32 decomposition offset carries no rounding
    # bit (cb/tgsw_functions.cpp:30-36), so the l*Bgbit kept bits are a truncation: error in (-2^(32-l*Bgbit), 0] (+FFT LSB)
    assert np.abs(centered(got - want)).max() <= 2**(32 - l * Bgbit) + 1
    # in place on d1
    d1d = dev(d1)
    engine.CMux(d1d, gfft, 1, d1d, dev(d0), l, Bgbit, B)
    torch.cuda.synchronize()
    assert np.array_equal(d1d.cpu().numpy(), got)


def test_lut_noiseless_selectors(engine):
    l, Bgbit, nsel = 2, 8, 4
    rng = np.random.default_rng(8)
    B = 5
    idx = np.array([0, 15, 6, 9, 3])
    gsw = np.stack([[noiseless_tgsw((int(i) >> j) & 1, l, Bgbit) for j in range(nsel)] for i in idx])      # [B][nsel][2l][2][N]
    gfft = torch.empty((B * nsel, 2 * l, 2, N), dtype=torch.float64, device=DEV)
    engine.tGswToFFTConvert(gfft, dev(gsw.reshape(B * nsel, 2 * l, 2, N)), l, B * nsel)
    table = rng.integers(-2**31, 2**31 - 1, size=(1 << nsel, N), dtype=np.int64).astype(np.int32)
    out = torch.empty((B, 2, N), dtype=torch.int32, device=DEV)
    engine.LUT_vertical_packing(out, gfft, nsel, dev(table), l, Bgbit, B)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.abs(centered(got[:, 0, :])).max() <= nsel          # trivial samples stay (almost) trivial: a = 0 up to FFT rounding
    assert np.abs(centered(got[:, 1, :] - table[idx])).max() <= nsel * (2**(32 - l * Bgbit) + 1)


def test_lut_on_circuit_bootstrapped_bits(engine, cb_oracle):
    """config 3 end to end: LWE bits -> tfhe_CircuitBootstrapFFT -> TRGSW selectors -> vertical-packing LUT -> TRLWE of table[index]."""
    c = cb_oracle
    engine.load_cb_keys(c.engine_params(), c.preKS, c.bk, c.privKS)
    l, Bgbit, nsel = c.params.ell_lvl1, c.params.bgbit_lvl1, 3
    rng = np.random.default_rng(9)
    B = 6
    idx = rng.integers(0, 1 << nsel, size=B)
    bits = np.array([[(int(i) >> j) & 1 for j in range(nsel)] for i in idx]).reshape(-1)                     # sample-major, selector j
    x = c.encrypt_lvl1((bits.astype(np.int64) << 31).astype(np.int32), 2.0**-20, seed=47)
    sel = torch.empty((B * nsel, 2, l, 2, c.N1), dtype=torch.int32, device=DEV)
    engine.tfhe_CircuitBootstrapFFT(sel, dev(x), B * nsel)
    selfft = torch.empty((B * nsel, 2 * l, 2, c.N1), dtype=torch.float64, device=DEV)
    engine.tGswToFFTConvert(selfft, sel, l, B * nsel)
    table = (rng.integers(0, 16, size=(1 << nsel, c.N1)).astype(np.int64) << 28).astype(np.int32)           # messages in sixteenths
    out = torch.empty((B, 2, c.N1), dtype=torch.int32, device=DEV)
    engine.LUT_vertical_packing(out, selfft, nsel, dev(table), l, Bgbit, B)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    worst = 0
    for b in range(B):
        err = centered(c.tlwe_phase_lvl1(got[b]) - table[idx[b]])
        worst = max(worst, int(np.abs(err).max()))
    assert worst < 2**26, f"phase error 2^{np.log2(max(worst, 1)):.1f}: a sixteenth (2^28) would not decode safely"
