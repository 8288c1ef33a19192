"""CPU tests: the C-ABI library loads and exports every symbol include/tfhe_b200.h declares, fails loudly without a
GPU (no fallback), and the host-side sharding/replication helpers work across 2 ranks on gloo."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pkg():
    return importlib.import_module("experimental-tfhe_b200")


def test_library_exports_every_declared_symbol():
    mod = pkg()
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    header = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tfhe_b200_[A-Za-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    lib = ctypes.CDLL(mod.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/tfhe_b200.h but not exported"
    assert sorted(mod.EXPORTS) == declared, "python binding table and header disagree"
    # nothing from the oracle is linked into the product
    nm = subprocess.run(["nm", "-D", mod.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in nm


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mod = pkg()
    with pytest.raises(mod.EngineError) as e:
        mod.Engine(0)
    assert "no CUDA device" in str(e.value) or "-4" in str(e.value)


def test_product_sources_do_not_reference_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "experimental-tfhe_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, f"{f} mentions the oracle"


def test_gate_tables_match_between_oracle_and_engine():
    """kGate in capi.cu and g_gate in the oracle encode the same (c, ka, kb) per gate."""
    capi = open(os.path.join(ROOT, "experimental-tfhe_b200", "csrc", "capi.cu")).read()
    orc = open(os.path.join(ROOT, "oracle", "tfhe_oracle.c")).read()
    t1 = re.search(r"kGate\[TFHE_B200_NUM_GATES\] = \{(.*?)\};", capi, re.S).group(1)
    t2 = re.search(r"g_gate\[ORC_NUM_GATES\] = \{(.*?)\};", orc, re.S).group(1)
    n1 = [int(x) for x in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", t1))]
    n2 = [int(x) for x in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", t2))]
    assert n1 == n2 and len(n1) == 30


def test_shard_range_covers_everything():
    par = importlib.import_module("experimental-tfhe_b200.parallel")
    for count in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import importlib, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
par = importlib.import_module("experimental-tfhe_b200.parallel")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lo, hi = par.shard_range(1001, rank, 2)
total = par.sum_over_ranks(hi - lo)
assert total == 1001, total
blob = torch.arange(4096, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
par.broadcast_bytes(blob, src=0)
assert torch.equal(blob, torch.arange(4096, dtype=torch.int64).to(torch.uint8))
t = par.max_over_ranks(1.0 + rank)
assert t == 2.0, t
dist.barrier()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharding_and_key_broadcast(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/tfhe_b200.h must compile as C99 (no C++ or CUDA types in the signatures)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "tfhe_b200.h"\nint main(void) { tfhe_b200_gate g = {0, 0, 0, 0, 0}; (void)g; return 0; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bench_reference_arm_contract():
    """bench.py --impl reference: one JSON line on stdout with the contract's keys (CPU only; bounded sample)."""
    import json, subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "gates/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_engine_arm_refuses_without_gpu():
    """No CPU fallback: on a machine without a CUDA device the engine arm stops with a clear message instead of printing a number."""
    import subprocess, sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_exact_ntt_arithmetic_on_host(tmp_path):
    """The Goldilocks field arithmetic and the limb-split NTT external product of the exact Torus64 path are __host__ __device__ code
    (experimental-tfhe_b200/csrc/exact_ntt.cuh): compiled with g++ and checked against 128-bit integer arithmetic and a schoolbook negacyclic product
    mod 2^64 -- no GPU, no oracle."""
    exe = str(tmp_path / "ntt_host_check")
    src = os.path.join(ROOT, "tests", "cpp", "ntt_host_check.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "experimental-tfhe_b200", "csrc"), "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "NTT HOST CHECK: all ok" in r.stdout, r.stdout + r.stderr


def test_ciphertext_wire_format_roundtrip():
    """tfhe_b200_ciphertext_pack / _unpack are host-side (no device): round trip, damaged payload, truncated blob."""
    import numpy as np
    mod = importlib.import_module("experimental-tfhe_b200")
    rng = np.random.default_rng(2)
    for kind, arr in (("LWE32", rng.integers(-2**31, 2**31 - 1, size=(5, 501), dtype=np.int64).astype(np.int32)),
                      ("LWE64", rng.integers(-2**63, 2**63 - 1, size=(3, 2049), dtype=np.int64)),
                      ("TLWE32", rng.integers(-2**31, 2**31 - 1, size=(2, 2, 1024), dtype=np.int64).astype(np.int32)),
                      ("TGSW32", rng.integers(-2**31, 2**31 - 1, size=(1, 4, 2, 1024), dtype=np.int64).astype(np.int32))):
        blob = mod.ciphertext_pack(kind, arr)
        assert blob.nbytes == 64 + arr.nbytes
        k2, a2 = mod.ciphertext_unpack(blob)
        assert k2 == kind and a2.dtype == arr.dtype and np.array_equal(a2, arr)
        bad = blob.copy(); bad[100] ^= 1
        with pytest.raises(mod.EngineError, match="checksum"):
            mod.ciphertext_unpack(bad)
        with pytest.raises(mod.EngineError, match="size"):
            mod.ciphertext_unpack(blob[:-8].copy())
