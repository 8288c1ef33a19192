set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call37.txt
: > $OUT
for lib in "" tcs1 tcs4 tcs2d15; do
  echo "lib=${lib:-default (2 steps per turn)}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 100 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"keyswitch_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
  echo >> $OUT
done
cat $OUT
timeout 900 python -m pytest tests/test_gpu_keyswitch.py tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -4
