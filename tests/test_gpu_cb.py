"""GPU parity tests for circuit bootstrapping (cb/poc_CircuitBootstrapping.cpp) through the C ABI.

Integer stages (preKeySwitch, preModSwitch, circuitPrivKS) are bit-exact against the oracle.  The 64-bit-torus blind
rotation runs through FP64 FFTs that keep 53 of ~85 product bits, so one external product must stay within
2^29 LSB of the exact integer product (the reference itself reaches 2^27.6) and end-to-end phases must decode.
"""
import numpy as np
import pytest
import torch

import oracle_lib as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def cb_engine(engine, cb_oracle):
    c = cb_oracle
    engine.load_cb_keys(c.engine_params(), c.preKS, c.bk, c.privKS)
    return engine


def test_preKeySwitch_preModSwitch_bit_exact(cb_engine, cb_oracle):
    c = cb_oracle
    rng = np.random.default_rng(1)
    B = 45
    x = rng.integers(-2**31, 2**31 - 1, size=(B, c.N1 + 1), dtype=np.int64).astype(np.int32)
    pre = torch.empty((B, c.n0 + 1), dtype=torch.int32, device=DEV)
    ms = torch.empty((B, c.n0 + 1), dtype=torch.int32, device=DEV)
    cb_engine.preKeySwitch(pre, dev(x), B)
    cb_engine.preModSwitch(ms, pre, B)
    torch.cuda.synchronize()
    ref_pre = c.preKeySwitch(x)
    assert np.array_equal(pre.cpu().numpy(), ref_pre)
    assert np.array_equal(ms.cpu().numpy(), c.preModSwitch(ref_pre))
    # modswitch edge values (wrap at the top of the torus, cb/poc_CircuitBootstrapping.cpp:481-482)
    edge = np.zeros((1, c.n0 + 1), np.int32)
    edge[0, :6] = [0, -1, 2**31 - 1, -2**31, 2**19 - 1, 2**19]
    cb_engine.preModSwitch(ms[:1], dev(edge), 1)
    torch.cuda.synchronize()
    assert np.array_equal(ms[:1].cpu().numpy(), c.preModSwitch(edge))


def test_circuitPrivKS_bit_exact(cb_engine, cb_oracle):
    c = cb_oracle
    rng = np.random.default_rng(2)
    B = 5
    x = rng.integers(-2**63, 2**63 - 1, size=(B, c.N2 + 1), dtype=np.int64)
    for u in (0, 1):
        out = torch.empty((B, 2, c.N1), dtype=torch.int32, device=DEV)
        cb_engine.circuitPrivKS(out, u, dev(x), B)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), c.circuitPrivKS(u, x))


def test_fft64_product_tolerance(cb_engine):
    """N=2048 / Torus64 / 8 products: deviation from the exact product <= 2^29 LSB (reference: max 2^27.6)."""
    rng = np.random.default_rng(3)
    N = 2048
    d = rng.integers(-256, 256, size=(8, N), dtype=np.int32)
    t = rng.integers(-2**63, 2**63 - 1, size=(8, N), dtype=np.int64)
    sd = torch.empty((8, N), dtype=torch.float64, device=DEV); st = torch.empty((8, N), dtype=torch.float64, device=DEV)
    cb_engine.IntPolynomial_ifft(sd, dev(d), N, 8)
    cb_engine.TorusPolynomial64_ifft(st, dev(t), N, 8)
    acc = torch.zeros((N,), dtype=torch.float64, device=DEV)
    for i in range(8):
        cb_engine.LagrangeHalfCPolynomialAddMul(acc, sd[i], st[i], N, 1)
    res = torch.empty((N,), dtype=torch.int64, device=DEV)
    cb_engine.TorusPolynomial64_fft(res, acc, N, 1)
    torch.cuda.synchronize()
    exact = np.zeros(N, np.int64)
    for i in range(8):
        O.lib().orc_torus64PolynomialMultAddNaive(O.p(exact), O.p(d[i]), O.p(t[i]), N)
    diff = (res.cpu().numpy() - exact).astype(np.int64)      # wraps mod 2^64
    assert np.abs(diff).max() <= 2**29, f"deviation 2^{np.log2(np.abs(diff).max()):.1f} LSB"


def test_circuitBootstrapWoKS_phase(cb_engine, cb_oracle):
    """Phase of the extracted LWE64 is mu for input phase in [1/4,3/4), else 0, noise like the oracle's."""
    c = cb_oracle
    B = 6
    msg = np.array([0, 1, 1, 0, 1, 0], dtype=np.int64) * (1 << 31)
    x = c.encrypt_lvl1(msg.astype(np.int32), 2.0**-20, seed=45)
    abar = c.preModSwitch(c.preKeySwitch(x))
    mu = 1 << 56
    out = torch.empty((B, c.N2 + 1), dtype=torch.int64, device=DEV)
    cb_engine.circuitBootstrapWoKS(out, mu, dev(abar), B)
    torch.cuda.synchronize()
    ph = c.phase_lvl2(out.cpu().numpy())
    expect = np.where(msg != 0, mu, 0)
    err = (ph - expect).astype(np.int64)
    ref = c.circuitBootstrapWoKS(mu, abar[:2])
    err_ref = (c.phase_lvl2(ref) - expect[:2]).astype(np.int64)
    assert np.abs(err).max() < 2**44, f"phase error 2^{np.log2(np.abs(err).max() + 1):.1f}"
    assert np.abs(err).max() < 64 * (np.abs(err_ref).max() + 2**36)


def test_full_circuit_bootstrap(cb_engine, cb_oracle):
    """tfhe_CircuitBootstrapFFT: the four TRGSW rows decrypt to -K*mu_w (u=0) and mu_w (u=1), cb/poc...:852-855."""
    c = cb_oracle
    B = 5
    msg = np.array([1, 0, 1, 1, 0], dtype=np.int64) * (1 << 31)
    x = c.encrypt_lvl1(msg.astype(np.int32), 2.0**-20, seed=46)
    ell1 = c.params.ell_lvl1
    out = torch.empty((B, 2, ell1, 2, c.N1), dtype=torch.int32, device=DEV)
    cb_engine.tfhe_CircuitBootstrapFFT(out, dev(x), B)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    host = np.zeros_like(got)
    cb_engine.tfhe_CircuitBootstrapFFT_host(host, x, B)
    for i in range(B):
        bit = int(msg[i] != 0)
        for w in range(ell1):
            mu_w = 1 << (32 - (w + 1) * c.params.bgbit_lvl1)
            for arr in (got, host):
                ph1 = c.tlwe_phase_lvl1(arr[i, 1, w]).astype(np.int64)
                exp1 = np.zeros(c.N1, np.int64); exp1[0] = bit * mu_w
                assert np.abs(ph1 - exp1).max() < 2**13
                ph0 = c.tlwe_phase_lvl1(arr[i, 0, w]).astype(np.int64)
                exp0 = -bit * mu_w * c.key_lvl1.astype(np.int64)
                assert np.abs(ph0 - exp0).max() < 2**13
