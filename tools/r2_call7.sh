set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keygen.py tests/test_gpu_compat_cpp.py -x -q -m gpu -s 2>&1 | tail -25 | tee gpurun_out/tests_call7.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee -a gpurun_out/tests_call7.txt
