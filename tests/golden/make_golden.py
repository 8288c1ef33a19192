"""Regenerates tests/golden/*.{i32,i64,f64} from the REFERENCE compiled in place (oracle/Makefile `ref` target).

Needs /root/reference (this container only).  The harness (oracle/ref_harness.cpp) feeds oracle-generated keys
(seed 42) and inputs (seed 45) to the reference's own functions -- spqlios transforms, Karatsuba, preKeySwitch,
preModSwitch, circuitBootstrapWoKS (with the D1-D3 corrections of SURVEY.md Appendix B), circuitPrivKS,
tfhe_CircuitBootstrapFFT -- checks that the oracle restatement reproduces them bit for bit when it runs on the
reference's FFT kernels, and dumps the reference outputs here.  pin_log.txt keeps the harness transcript.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

if not os.path.isdir("/root/reference"):
    sys.exit("make_golden.py needs /root/reference")
subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "golden", HERE], capture_output=True, text=True)
open(os.path.join(HERE, "pin_log.txt"), "w").write(r.stdout)
print(r.stdout)
sys.exit(r.returncode)
