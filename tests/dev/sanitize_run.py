"""Small invocation of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tests/dev/sanitize_run.py
Sizes are tiny (n = 8 CMUX steps, a handful of CTAs) so that the instrumented run ends in minutes; results are still checked
(gates decrypt, key switches bit-exact, 128-bit FFT bit-exact), so a pass means "no hazard reported AND right answers".
Development tool: transcripts go to profiles/ (VERDICT r1 "What's weak" #8)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O

which = sys.argv[1] if len(sys.argv) > 1 else "gate,cb,hp"
mod = importlib.import_module("experimental-tfhe_b200")
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

if "gate" in which:
    for variant in os.environ.get("SAN_VARIANTS", "default").split(","):
        if variant != "default": os.environ["TFHE_B200_BR_VARIANT"] = variant
        g = O.GateOracle(seed=5, n=8)
        eng = mod.Engine(0)
        eng.load_gate_keys(g.engine_params(), g.bk, g.ks)
        a = np.array([0, 1, 0, 1] * 5); b = np.array([0, 0, 1, 1] * 5)          # 20 gates: 3 CTAs, the last one ragged
        ca, cb = g.encrypt_bits(a, 1), g.encrypt_bits(b, 2)
        out = torch.empty((len(a), g.n + 1), dtype=torch.int32, device="cuda")
        eng.bootsNAND(out, dev(ca), dev(cb), len(a)); torch.cuda.synchronize()
        assert np.array_equal(g.decrypt_bits(out.cpu().numpy()), 1 - (a & b)), "NAND decrypts wrongly"
        u = np.random.default_rng(0).integers(-2**31, 2**31 - 1, size=(37, g.N + 1), dtype=np.int64).astype(np.int32)
        ks = torch.empty((37, g.n + 1), dtype=torch.int32, device="cuda")
        eng.lweKeySwitch(ks, dev(u), 37); torch.cuda.synchronize()
        assert np.array_equal(ks.cpu().numpy(), g.keyswitch(u)), "key switch not bit-exact"
        print(f"gate path ok ({variant})", flush=True)
        del eng

if "cb" in which:
    c = O.CBOracle(seed=9, with_privks=True, n_lvl0=8, kslength_lvl21=4, kslength_lvl10=2)
    eng = mod.Engine(0)
    eng.load_cb_keys(c.engine_params(), c.preKS, c.bk, c.privKS)
    B = 3
    msg = np.array([0, 1, 1], dtype=np.int64) * (1 << 31)
    x = c.encrypt_lvl1(msg.astype(np.int32), 2.0 ** -25, seed=4)
    ell1 = c.params.ell_lvl1
    res = torch.empty((B, 2, ell1, 2, c.N1), dtype=torch.int32, device="cuda")
    eng.tfhe_CircuitBootstrapFFT(res, dev(x), B); torch.cuda.synchronize()
    got = res.cpu().numpy()
    # integer stages bit-exact.  GPU and oracle accumulators are different encryptions of the same phase, and with
    # kslength_lvl21 = 4 the private key switch keeps 12 bits of each of the 2049 coefficients: the rows agree by phase to
    # about 2^23 (rounding 2^(32-13) x sqrt(1024)), compared here at 2^26
    pk = torch.empty((B, c.n0 + 1), dtype=torch.int32, device="cuda")
    eng.preKeySwitch(pk, dev(x), B); torch.cuda.synchronize()
    assert np.array_equal(pk.cpu().numpy(), c.preKeySwitch(x)), "preKeySwitch not bit-exact"
    ref = c.CircuitBootstrapFFT(x)
    for i in range(B):
        for u in range(2):
            for w in range(ell1):
                d = (c.tlwe_phase_lvl1(got[i, u, w]).astype(np.int64) - c.tlwe_phase_lvl1(ref[i, u, w]).astype(np.int64) + 2**31) % 2**32 - 2**31
                assert np.abs(d).max() < 2**26, "circuit bootstrap rows differ from the oracle's"
    print("circuit-bootstrap path ok", flush=True)
    del eng

if "hp" in which:
    eng = mod.Engine(0)
    for N in (2048, 4096):
        om, ob = O.hp_tables(N)
        x = np.random.default_rng(N).integers(-2**63, 2**63 - 1, size=(3, N), dtype=np.int64)
        spec = torch.empty((3, N // 2, 4), dtype=torch.int64, device="cuda"); back = torch.empty((3, N), dtype=torch.int64, device="cuda")
        eng.hp_iFFT(spec, dev(x), N, 3); eng.hp_FFT(back, spec, N, 3); torch.cuda.synchronize()
        ref_s = np.stack([O.hp_iFFT(x[i], N, om) for i in range(3)])
        assert np.array_equal(spec.cpu().numpy().view(np.uint64), ref_s), "hp iFFT not bit-exact"
        ref_b = np.stack([O.hp_FFT(ref_s[i], N, ob) for i in range(3)])
        assert np.array_equal(back.cpu().numpy(), ref_b), "hp FFT not bit-exact"
    print("hp path ok", flush=True)
print("sanitize_run: all selected paths ok")
