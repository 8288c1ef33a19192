set -u
mkdir -p gpurun_out
OUT=gpurun_out/pre0_call41.txt
: > $OUT
for lib in "" nopre0 "" nopre0; do
  echo "lib=${lib:-default (PRE0)}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 100 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
for lib in "" nopre0; do
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"blind_rotate_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
  echo >> $OUT
done
cat $OUT
timeout 900 python -m pytest tests/test_gpu_gate.py tests/test_gpu_trgsw.py tests/test_gpu_params.py tests/test_gpu_cb.py tests/test_gpu_cb64.py -x -q -m gpu 2>&1 | tail -4
