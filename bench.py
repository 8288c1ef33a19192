#!/usr/bin/env python
"""bench.py -- gate bootstraps/sec on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (bootsNAND = linear combination + modulus switch + blind rotation + sample
extraction + key switching) over one batch of 65,536 independent synthetic NAND gates per GPU (BASELINE.json
configs[1]); ranks own disjoint batches and replicated keys, no collective runs inside the timed region (weak scaling).

The JSON line carries
  value      device-resident throughput (inputs already in HBM), CUDA events on the launching stream, max over ranks
  e2e        the same metric through tfhe_b200_bootsGate_batch_host (pinned host buffers, H2D + D2H inside the timing)
  roofline   the blind-rotation kernel against the FP64 pipe (this path is FP64-bound), peak measured
             in-run by a DFMA probe; roofline_hbm gives the HBM view against MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU path (oracle gate restatement on the reference's spqlios kernels) on all host cores

--impl reference times that CPU path alone (oracle/_ref when built from /root/reference, else the portable port).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BATCH = 65536
FLOP_PER_GATE = 94.72e6          # 500 CMUX x 189,440 flop (radix-2 count)
HBM_BYTES_PER_GATE = 6012        # 2 LWE in + 1 LWE out, (n+1) x 4 B each
METRIC = "gate bootstraps/sec (batch 64k)"
WORKLOAD = "65536 bootsNAND per GPU, n=500 N=1024 k=1 l=2 Bgbit=10, KS t=8 basebit=2"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_gate_rate(count, threads):
    """Reference CPU path: prefer oracle/_ref/ref_harness (reference spqlios kernels compiled in place), else the port."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if os.path.exists(harness):
        try:
            r = subprocess.run([harness, "bench-gate", str(count), str(threads), "1"], capture_output=True, text=True, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
            return d["gates_per_s"], "reference", f"{count} bootsNAND on {threads} threads: oracle gate path over the reference's spqlios-fma FFT (oracle/_ref)"
        except Exception:
            pass
    import numpy as np
    import oracle_lib as O
    g = O.GateOracle(42)
    rng = np.random.default_rng(44)
    ca = rng.integers(-2**31, 2**31 - 1, size=(count, g.n + 1), dtype=np.int64).astype(np.int32)
    cb = rng.integers(-2**31, 2**31 - 1, size=(count, g.n + 1), dtype=np.int64).astype(np.int32)
    g.bootsGate("NAND", ca[:threads], cb[:threads], threads)
    t0 = time.perf_counter()
    g.bootsGate("NAND", ca, cb, threads)
    dt = time.perf_counter() - t0
    return count / dt, "port", f"{count} bootsNAND on {threads} threads: oracle port with its portable FFT"


def _harness(args, timeout=900):
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    try:
        r = subprocess.run([harness] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    except Exception:
        return None


def cpu_cb_rate(count, threads):
    """Reference CPU circuit bootstrap: the oracle's tfhe_CircuitBootstrapFFT (bit-identical to the reference's, tests/golden/pin_log.txt)
    over the reference's spqlios kernels, OpenMP over samples (oracle/_ref); else the portable port on one thread."""
    d = _harness(["bench-cb", count, threads])
    if d:
        return d["cb_per_s"], "reference", f"{count} tfhe_CircuitBootstrapFFT on {threads} threads: oracle path over the reference's spqlios-fma FFT (oracle/_ref)", threads
    import numpy as np
    import oracle_lib as O
    c = O.CBOracle(42)
    x = np.random.default_rng(45).integers(-2**31, 2**31 - 1, size=(2, c.N1 + 1), dtype=np.int64).astype(np.int32)
    t0 = time.perf_counter(); c.CircuitBootstrapFFT(x); dt = time.perf_counter() - t0
    return 2 / dt, "port", "2 tfhe_CircuitBootstrapFFT on 1 thread: oracle port with its portable FFT", 1


def cpu_hp_rate(N, count, threads):
    d = _harness(["bench-hp", N, count, threads])
    if d:
        return d["ifft_per_s"], d["fft_per_s"], "port", f"{count} transforms each way on {threads} threads: hp/code.cpp restated (oracle/hpfft_oracle.c, -Ofast)", threads
    import numpy as np
    import oracle_lib as O
    om, ob = O.hp_tables(N)
    x = np.random.default_rng(46).integers(-2**63, 2**63 - 1, size=(8, N), dtype=np.int64)
    t0 = time.perf_counter(); sp = [O.hp_iFFT(x[i], N, om) for i in range(8)]; t1 = time.perf_counter()
    [O.hp_FFT(sp[i], N, ob) for i in range(8)]; t2 = time.perf_counter()
    return 8 / (t1 - t0), 8 / (t2 - t1), "port", "8 transforms each way on 1 thread: hp/code.cpp restated (oracle/hpfft_oracle.c)", 1


CB_PARAMS = dict(n_lvl0=500, N_lvl1=1024, N_lvl2=2048, bgbit_lvl1=8, ell_lvl1=2, bgbit_lvl2=9, ell_lvl2=4,
                 kslength_lvl10=6, ksbasebit_lvl10=2, kslength_lvl21=10, ksbasebit_lvl21=3)      # cb/poc_CircuitBootstrapping.cpp:70-85
FLOP_PER_CB = 704.512e6          # 2 blind rotations x 500 CMUX x 704,512 flop
CB_BATCH = 4096


def _tensor_peak_tops():
    """int8 peak of the tensor pipe as measured by tools/imma_probe.cu (profiles/r2_imma_probe3.txt): a 128 x 256 x 32 u8 MMA retires in
    128 cycles when three warps keep the pipe fed = 8192 MAC per clock per SM; x 148 SMs x the SM clock of MEASURED_PEAKS.json (the key
    switch does not pull the clock down the way a long bf16 GEMM does, so 2 x bf16_tflops would under-state it)."""
    mhz = 1965.0
    try:
        mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", mhz))
    except Exception:
        pass
    return 2.0 * 8192 * 148 * mhz * 1e6 / 1e12, f"probe-measured 8192 u8 MAC/clk/SM (tools/imma_probe.cu) x 148 SMs x {mhz:.0f} MHz"


def _ks_tensor_roofline(kernel, samples, rows, t, basebit, cols_pad, nz, ms):
    """One-hot GEMM of a key switch on the tensor cores (csrc/ks_tc_kernels.cu): per sample, step and output column 32 rows x 4 byte
    planes of u8 x u8 multiply-accumulates are issued; only 1 in 2^basebit rows of the one-hot operand is non-zero, so the USEFUL
    integer additions (the reference's count) are given next to the issued tensor operations."""
    q = 32 // (1 << basebit)
    steps = (rows * t + q - 1) // q
    macs = float(samples) * nz * steps * 32 * 4 * cols_pad
    peak, src = _tensor_peak_tops()
    ach = 2.0 * macs / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak, "peak_source": src,
            "traffic": None, "key_image_bytes": nz * (cols_pad // 128) * steps * 16384,
            "useful_int32_adds": float(samples) * nz * rows * t * (1.0 - 2.0 ** -basebit) * cols_pad,
            "algorithmic": "u8 x u8 -> s32 MACs issued: samples x steps x 32 rows x 4 byte planes x padded columns (one-hot operand)"}


def _traffic(kernel):
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]
            return tr
        except Exception:
            continue
    return None


def section_circuit_bootstrap(torch, np, eng, par, dist, rank, world, dev, fp64_peak, steps, cpu_baseline):
    """BASELINE configs[3]: 4,096 independent tfhe_CircuitBootstrapFFT per GPU at the reference's active parameter set
    (cb/poc_CircuitBootstrapping.cpp:70-85, timing loop :1008-1016), keys of those shapes (uniform random), weak scaling."""
    p = CB_PARAMS
    B = CB_BATCH
    pre = bk = priv = None
    t0 = time.perf_counter()
    if rank == 0:
        krng = np.random.default_rng(43)
        bk = krng.integers(-2**63, 2**63 - 1, size=(p["n_lvl0"], 2 * p["ell_lvl2"], 2, p["N_lvl2"]), dtype=np.int64)
        pre = krng.integers(-2**31, 2**31 - 1, size=(p["N_lvl1"], p["kslength_lvl10"], 1 << p["ksbasebit_lvl10"], p["n_lvl0"] + 1), dtype=np.int32)
        priv = krng.integers(-2**31, 2**31 - 1, size=(2, p["N_lvl2"] + 1, p["kslength_lvl21"], 1 << p["ksbasebit_lvl21"], 2, p["N_lvl1"]), dtype=np.int32)
    t_gen = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    par.replicate_cb_keys(eng, p, pre, bk, priv, device=dev)
    torch.cuda.synchronize()
    t_rep = par.max_over_ranks(time.perf_counter() - t0, device=dev)
    del pre, bk, priv
    gen = torch.Generator(device=dev).manual_seed(45 + rank)
    x = torch.randint(-2**31, 2**31 - 1, (B, p["N_lvl1"] + 1), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    res = torch.empty((B, 2, p["ell_lvl1"], 2, p["N_lvl1"]), dtype=torch.int32, device=dev)
    h_x = torch.empty(x.shape, dtype=torch.int32).pin_memory(); h_x.copy_(x)
    h_res = torch.empty(res.shape, dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.tfhe_CircuitBootstrapFFT(res, x, B)
    barrier()
    eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.tfhe_CircuitBootstrapFFT(res, x, B)
    e1.record()
    barrier()
    ms = par.max_over_ranks(e0.elapsed_time(e1), device=dev) / steps
    kern_ms, kern_n = eng.profile_read()
    eng.profile_enable(False)
    eng.tfhe_CircuitBootstrapFFT_host(h_res, h_x, B)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.tfhe_CircuitBootstrapFFT_host(h_res, h_x, B)
    torch.cuda.synchronize()
    t_e2e = par.max_over_ranks(time.perf_counter() - t0, device=dev) / steps
    checksum = int(h_res[:, 1, 0, 1, 0].to(torch.int64).sum().item())
    barrier()
    if rank != 0:
        return None
    br_ms = kern_ms["blind_rotate"] / max(kern_n["blind_rotate"], 1)
    ks_ms = kern_ms["keyswitch"] / steps                       # preKS + the one launch of all four private key switches
    tf = FLOP_PER_CB * B / (br_ms * 1e-3) / 1e12
    tr_br, tr_ks = _traffic("blind_rotate_kernel_n2048"), _traffic("keyswitch_kernel_privks")
    out = {"metric": "circuit bootstraps/sec (batch 4096)", "value": world * B / (ms * 1e-3), "unit": "circuit bootstraps/s", "ms_per_step": ms,
           "steps": steps, "scaling": "weak",
           "config": {"workload": "4096 tfhe_CircuitBootstrapFFT per GPU, n0=500 N1=1024 N2=2048 l1=2 Bg1=2^8 l2=4 Bg2=2^9, KS10 6x2 bit, KS21 10x3 bit",
                      "batch_per_gpu": B},
           "e2e": {"value": world * B / t_e2e, "unit": "circuit bootstraps/s", "h2d_bytes_per_step": int(h_x.numel()) * 4,
                   "d2h_bytes_per_step": int(h_res.numel()) * 4, "api": "tfhe_b200_CircuitBootstrapFFT_batch_host", "result_checksum": checksum},
           "kernel_ms_per_step": {"blind_rotate": br_ms, "keyswitch": ks_ms, "other": kern_ms["other"] / steps},
           "roofline": {"bound": "fp64", "kernel": "blind_rotate_kernel<10,int64_t>", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": tf / fp64_peak if fp64_peak else None,
                        "traffic": tr_br["dram_bytes_per_launch"] if tr_br and tr_br.get("batch") == B else None,
                        "traffic_source": tr_br["source"] if tr_br and tr_br.get("batch") == B else None,
                        "algorithmic": "704.5 MFLOP per circuit bootstrap (both mu_w in one launch: 8192 rotations of 352.3 MFLOP)"},
           "privks": {"kernel": "keyswitch_tc_kernel<int64_t,3> (4 private key switches, one launch) + keyswitch_tc_kernel<int32_t,2> (preKS)",
                      "ms": ks_ms, "int32_adds_per_cb": 146.9e6,
                      "roofline": _ks_tensor_roofline("keyswitch_tc_kernel<int64_t,3>", B * p["ell_lvl1"], p["N_lvl2"] + 1, p["kslength_lvl21"],
                                                      p["ksbasebit_lvl21"], 2 * p["N_lvl1"], 2, ks_ms),
                      "traffic": tr_ks["dram_bytes_per_launch"] if tr_ks and tr_ks.get("batch") == B else None,
                      "traffic_source": tr_ks["source"] if tr_ks and tr_ks.get("batch") == B else None},
           "key_replication_ms": t_rep * 1e3, "key_bytes_replicated": int(sum(eng.cb_key_blob(w)[1] for w in (0, 1, 2))), "host_keygen_s": t_gen,
           "gpu_launches": int(sum(kern_n.values()))}
    if cpu_baseline and world == 1:
        threads = host_threads()
        rate, kind, sample, used = cpu_cb_rate(max(2 * threads, 8), threads)
        out["cpu_baseline"] = {"value": rate, "unit": "circuit bootstraps/s", "cores": used, "kind": kind, "sample": sample}
    return out


def section_hp_fft(torch, np, eng, par, dist, rank, world, dev, cpu_baseline):
    """BASELINE configs[4] (standalone): 128-bit fixed-point anticyclic FFT, N = 2048 / 4096, 16,384 polynomials per GPU
    (hp/code.cpp:391-512; the reference times it at :574-586)."""
    B = 16384
    peak = eng.probe_real96_gprods() if rank == 0 else 0.0
    res = {}
    for N, cprod in ((2048, 6144), (4096, 13312)):              # complex fixed-point products per transform
        gen = torch.Generator(device=dev).manual_seed(46 + rank)
        x = torch.randint(-2**63, 2**63 - 1, (B, N), dtype=torch.int64, device=dev, generator=gen)
        spec = torch.empty((B, N // 2, 4), dtype=torch.int64, device=dev)
        back = torch.empty((B, N), dtype=torch.int64, device=dev)
        eng.hp_iFFT(spec, x, N, B); eng.hp_FFT(back, spec, N, B)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            eng.hp_iFFT(spec, x, N, B)
        ev[1].record()
        for _ in range(reps):
            eng.hp_FFT(back, spec, N, B)
        ev[2].record()
        torch.cuda.synchronize()
        t_i = par.max_over_ranks(ev[0].elapsed_time(ev[1]), device=dev) / reps
        t_f = par.max_over_ranks(ev[1].elapsed_time(ev[2]), device=dev) / reps
        err = int((back - x).abs().max().item())
        if rank == 0:
            ach_i = 4 * cprod * B / (t_i * 1e-3) / 1e9
            ach_f = 4 * cprod * B / (t_f * 1e-3) / 1e9
            res[f"N{N}"] = {"batch_per_gpu": B, "iFFT_per_s": world * B / (t_i * 1e-3), "FFT_per_s": world * B / (t_f * 1e-3), "iFFT_ms": t_i, "FFT_ms": t_f,
                            "roundtrip_max_err_lsb": err,
                            "roofline": {"bound": "int64-multiply", "kernel": "hp_ifft_kernel / hp_fft_kernel", "unit": "G real96 products/s",
                                         "achieved": ach_i, "achieved_fft": ach_f, "peak": peak, "frac": ach_i / peak if peak else None,
                                         "frac_fft": ach_f / peak if peak else None,
                                         "algorithmic": f"{cprod} complex products x 4 real96 products per transform (SURVEY 8d)",
                                         "peak_source": "real96_mul probe (same arithmetic, operands in registers) measured in this run",
                                         "hbm_bytes_per_transform": N * 8 + N // 2 * 32}}
            if cpu_baseline and world == 1:
                threads = host_threads()
                ri, rf, kind, sample, used = cpu_hp_rate(N, 64 * threads, threads)
                res[f"N{N}"]["cpu_baseline"] = {"value": ri, "value_fft": rf, "unit": "transforms/s", "cores": used, "kind": kind, "sample": sample}
        del x, spec, back
    if rank != 0:
        return None
    return {"metric": "128-bit fixed-point anticyclic FFT transforms/sec (batch 16384)", "unit": "transforms/s", "scaling": "weak",
            "value": res["N2048"]["iFFT_per_s"], **res}


def adder_netlist(mod, bits):
    """32-bit ripple-carry adder, 5 bootstrapped gates per bit: wires a[bits] b[bits] cin | x[bits] | s[bits] | carries"""
    G = mod.GATES
    a0, b0, cin = 0, bits, 2 * bits
    x0 = cin + 1; s0 = x0 + bits; c0 = s0 + bits; t0 = c0 + bits; u0 = t0 + bits
    gates = [(G["XOR"], x0 + i, a0 + i, b0 + i, 0) for i in range(bits)]            # one merged launch of `bits` gates
    gates += [(G["AND"], t0 + i, a0 + i, b0 + i, 0) for i in range(bits)]           # and another
    for i in range(bits):
        c_in = cin if i == 0 else c0 + i - 1
        gates.append((G["XOR"], s0 + i, x0 + i, c_in, 0))
        gates.append((G["AND"], u0 + i, x0 + i, c_in, 0))
        gates.append((G["OR"], c0 + i, t0 + i, u0 + i, 0))
    return gates, dict(a0=a0, b0=b0, cin=cin, s0=s0, n_wires=u0 + bits)


def section_adder32(torch, np, mod, eng, par, dist, rank, world, dev, total_adders):
    """BASELINE configs[2]: 32-bit ripple-carry adders, the batch sharded over the GPUs (strong scaling: `total_adders` in total),
    gate keys replicated.  Inputs are synthetic ciphertexts: timing is data independent; sums are checked in tests/test_gpu_circuit.py."""
    lo, hi = par.shard_range(total_adders, rank, world)
    B = hi - lo
    bits = 32
    gates, w = adder_netlist(mod, bits)
    n = eng.gate_params.n
    gen = torch.Generator(device=dev).manual_seed(47 + rank)
    wires = torch.randint(-2**31, 2**31 - 1, (w["n_wires"], max(B, 1), n + 1), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    if B > 0:
        eng.circuit_eval(gates, wires, w["n_wires"], B)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if B > 0:
        eng.circuit_eval(gates, wires, w["n_wires"], B)
    e1.record()
    torch.cuda.synchronize()
    ms = par.max_over_ranks(e0.elapsed_time(e1), device=dev)
    if rank != 0:
        return None
    return {"metric": "32-bit ripple-carry adders/sec", "value": total_adders / (ms * 1e-3), "unit": "adders/s", "ms": ms, "scaling": "strong",
            "bootstrapped_gates_per_s": 160 * total_adders / (ms * 1e-3),
            "config": {"workload": f"{total_adders} adders in total, sharded {world} ways (contiguous ranges), 160 bootstrapped gates each, "
                                   "carry chain = 64 dependent levels of `adders per GPU` gates", "adders_per_gpu": B}}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# stdout carries exactly ONE JSON line.  Libraries write banners to file descriptor 1 behind Python's back (NCCL prints
# "NCCL version ..." there when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and
# the result line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def _guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    per_step = max(threads * 16, 16)                  # ~0.25 s/gate/thread-ish: 16 gates per thread per step ~ 0.3 s
    rates = []
    kind = sample = None
    for i in range(args.warmup + args.steps):
        rate, kind, sample = cpu_gate_rate(per_step, threads)
        if i >= args.warmup:
            rates.append(rate)
    value = sum(rates) / len(rates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_step / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (uniform random ciphertexts; keys from the oracle's key generator, seed 42)",
            "config": {"workload": WORKLOAD, "cpu_sample_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "gates/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="gates per GPU per step (default: the BASELINE configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gate-only", action="store_true", help="skip the circuit-bootstrap / hp-FFT / adder sections")
    ap.add_argument("--adders", type=int, default=8192, help="32-bit adders in total (sharded over the GPUs)")
    args = ap.parse_args()
    _guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    mod = importlib.import_module("experimental-tfhe_b200")
    par = importlib.import_module("experimental-tfhe_b200.parallel")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = mod.Engine(local)

    # keys: synthetic key material of the P_gate shapes (uniform random torus values, seed 42 -- key generation is client side and
    # out of scope, and the oracle is not used on this arm; timing does not depend on the values: a real key's polynomials are
    # uniform too).  Generated on rank 0's host, transformed on its GPU, replicated by NCCL broadcast.
    B = args.batch
    params = dict(n=500, N=1024, k=1, bk_l=2, bk_Bgbit=10, ks_t=8, ks_basebit=2)
    bk_host = ks_host = None
    if rank == 0:
        krng = np.random.default_rng(42)
        bk_host = krng.integers(-2**31, 2**31 - 1, size=(params["n"], 2 * params["bk_l"], 2, params["N"]), dtype=np.int64).astype(np.int32)
        ks_host = krng.integers(-2**31, 2**31 - 1, size=(params["N"], params["ks_t"], 1 << params["ks_basebit"], params["n"] + 1),
                                dtype=np.int64).astype(np.int32)
    torch.cuda.synchronize()
    t_rep0 = time.perf_counter()
    par.replicate_gate_keys(eng, params, bk_host, ks_host, device=dev)
    torch.cuda.synchronize()
    gate_key_ms = par.max_over_ranks(time.perf_counter() - t_rep0, device=dev) * 1e3      # ingest on rank 0 + broadcast, once per job
    n = params["n"]

    # synthetic ciphertexts: i.i.d. uniform int32 (timing is data independent), distinct per rank
    gen = torch.Generator(device=dev).manual_seed(44 + rank)
    ca = torch.randint(-2**31, 2**31 - 1, (B, n + 1), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    cb = torch.randint(-2**31, 2**31 - 1, (B, n + 1), dtype=torch.int64, device=dev, generator=gen).to(torch.int32)
    out = torch.empty((B, n + 1), dtype=torch.int32, device=dev)
    h_ca = torch.empty((B, n + 1), dtype=torch.int32).pin_memory(); h_ca.copy_(ca)
    h_cb = torch.empty((B, n + 1), dtype=torch.int32).pin_memory(); h_cb.copy_(cb)
    h_out = torch.empty((B, n + 1), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.bootsGate("NAND", out, ca, cb, B)

    for _ in range(args.warmup):
        step()
    barrier()
    fp64_peak = eng.probe_fp64_tflops() if rank == 0 else 0.0
    l2_gbs = eng.probe_read_gbs(32 << 20, 64) if rank == 0 else 0.0
    barrier()

    # ---- timed region 1: device resident
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = par.max_over_ranks(e0.elapsed_time(e1), device=dev)
    kern_ms, kern_n = eng.profile_read()
    eng.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(par.sum_over_ranks(sum(kern_n.values()), device=dev))

    # ---- timed region 2: end to end through the host-buffer C-ABI call (H2D + compute + D2H)
    eng.bootsGate_host("NAND", h_out, h_ca, h_cb, B)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.bootsGate_host("NAND", h_out, h_ca, h_cb, B)
    torch.cuda.synchronize()
    t_e2e = par.max_over_ranks(time.perf_counter() - t0, device=dev)
    checksum = int(h_out[:, n].to(torch.int64).sum().item())      # the device->host result is really read
    barrier()

    # ---- strong scaling of the same metric: 65,536 gates IN TOTAL, sharded over the ranks (tail waves and launch overhead show here)
    lo, hi = par.shard_range(B, rank, world)
    sB = hi - lo
    eng.bootsGate("NAND", out, ca, cb, sB)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        eng.bootsGate("NAND", out, ca, cb, sB)
    s1.record()
    barrier()
    strong_ms = par.max_over_ranks(s0.elapsed_time(s1), device=dev) / args.steps

    extra = {}
    if not args.gate_only:
        del h_ca, h_cb, h_out
        extra["adder32"] = section_adder32(torch, np, mod, eng, par, dist, rank, world, dev, args.adders)
        extra["hp_fft"] = section_hp_fft(torch, np, eng, par, dist, rank, world, dev, not args.no_cpu_baseline)
        extra["circuit_bootstrap"] = section_circuit_bootstrap(torch, np, eng, par, dist, rank, world, dev, fp64_peak,
                                                               min(args.steps, 3), not args.no_cpu_baseline)

    if rank == 0:
        gates = world * B * args.steps
        value = gates / (ms_total * 1e-3)
        br_ms = kern_ms["blind_rotate"] / max(kern_n["blind_rotate"], 1)
        achieved_tf = FLOP_PER_GATE * B / (br_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        # DRAM bytes per launch of the blind-rotation kernel, from the committed ncu --set full capture of this same command
        # (profiles/r1_traffic.json says which report); only quoted when the capture was taken at this batch size
        traffic, traffic_src = None, None
        tr = _traffic("blind_rotate_kernel")
        if tr and tr.get("batch") == B:
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
        ks_ms = kern_ms["keyswitch"] / max(kern_n["keyswitch"], 1)
        line = {
            "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (uniform random ciphertexts and key material of the P_gate shapes)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "parallelism": f"batch-sharded x{world}, keys replicated",
                       "l2": "inputs+outputs per step (393 MB + 268 MB scratch) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": gates / t_e2e, "unit": "gates/s", "h2d_bytes_per_step": 2 * B * (n + 1) * 4, "d2h_bytes_per_step": B * (n + 1) * 4,
                    "api": "tfhe_b200_bootsGate_batch_host", "result_checksum": checksum},
            "gpu_launches": launches,
            "kernel_ms_per_step": {k: v / args.steps for k, v in kern_ms.items()},
            "roofline": {"bound": "fp64", "kernel": "blind_rotate_kernel<9,int32_t>", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "DFMA probe measured in this run (nominal 37.2 at 1965 MHz)",
                         "algorithmic": "94.72 MFLOP per gate bootstrap x gates per launch (SURVEY 8d)",
                         "l2_read_gbs_measured": l2_gbs},
            "roofline_keyswitch": _ks_tensor_roofline("keyswitch_tc_kernel<int32_t,2>", B, 1024, 8, 2, 512, 1, ks_ms),
            "roofline_l2": {"bound": "l2", "achieved": 32.768e6 * B / (br_ms * 1e-3) / 1e9, "peak": l2_gbs, "unit": "GB/s",
                            "note": "bootstrapping-key stream, 32.77 MB per bootstrap per accumulator (SURVEY 8d), against the L2 read "
                                    "bandwidth measured in this run: the co-bound of the FP64 roofline, not the limiter"},
            "roofline_hbm": {"bound": "hbm", "achieved": HBM_BYTES_PER_GATE * B * args.steps / (ms_total * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "peak_source": hbm_src,
                             "note": "ciphertext I/O only (6012 B per gate); the 32.8 MB key stream is L2 resident"},
            "clocks": clocks,
            "strong": {"scaling": "strong", "value": B / (strong_ms * 1e-3), "unit": "gates/s", "ms_per_step": strong_ms,
                       "config": {"workload": f"{B} bootsNAND in total, sharded {world} ways", "gates_per_gpu": sB}},
            "key_replication_ms": gate_key_ms, "key_bytes_replicated": int(sum(eng.gate_key_blob(w)[1] for w in (0, 1))),
        }
        line.update({k: v for k, v in extra.items() if v is not None})
        if not args.no_cpu_baseline and world == 1:
            threads = host_threads()
            rate, kind, sample = cpu_gate_rate(threads * 256, threads)
            rate1, _, _ = cpu_gate_rate(64, 1)            # one thread, for comparison with the reference README's per-core figures
            line["cpu_baseline"] = {"value": rate, "unit": "gates/s", "cores": threads, "kind": kind, "sample": sample,
                                    "single_thread_value": rate1}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
