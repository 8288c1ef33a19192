set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call22.txt
: > $OUT
timeout 100 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
echo >> $OUT
cat $OUT
timeout 600 python -m pytest tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -4
