"""GPU test of the C++ host shim (experimental-tfhe_b200/host/tfhe_b200_compat.hpp): compiles tests/cpp/compat_test.cpp
against libtfhe_b200.so (product) and liboracle.so (checker) and runs it.  The program uses the reference's own function
names -- bootsNAND, tfhe_bootstrap_FFT, lweKeySwitch, tfhe_MuxRotate_FFT, preKeySwitch, circuitBootstrapWoKS ... --
on reference-shaped structs."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_against_oracle(tmp_path):
    exe = str(tmp_path / "compat_test")
    pkg = os.path.join(ROOT, "experimental-tfhe_b200")
    orc = os.path.join(ROOT, "oracle")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=gnu++17", os.path.join(ROOT, "tests", "cpp", "compat_test.cpp"), "-o", exe,
           "-I/usr/local/cuda/include", "-L" + pkg, "-ltfhe_b200", "-L" + orc, "-loracle", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + orc, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "COMPAT: all checks passed" in r.stdout
