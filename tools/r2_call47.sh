set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_trgsw.py tests/test_gpu_circuit.py -x -q -m gpu 2>&1 | tail -3
TFHE_B200_BR_VARIANT=keytm timeout 100 python tests/dev/quick_bench.py 8192 2>&1 | tail -3
timeout 100 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1
