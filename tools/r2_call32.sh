set -u
mkdir -p gpurun_out
timeout 300 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*\|"blind_rotate_ms": [0-9.]*' | tr '\n' ' ' | tee gpurun_out/sharetw_call32.txt
echo
timeout 900 python -m pytest tests/test_gpu_cb.py tests/test_gpu_cb64.py tests/test_gpu_trgsw.py -x -q -m gpu 2>&1 | tail -4
