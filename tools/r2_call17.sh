set -u
mkdir -p gpurun_out
OUT=gpurun_out/privks_call17.txt
: > $OUT
for lib in "" ksw16n12 ksw16n8; do
  echo "lib=${lib:-default}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 300 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*\|"blind_rotate_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
  echo >> $OUT
done
cat $OUT
TFHE_B200_LIB=$PWD/tools/alt/libtfhe_b200_ksw16n12.so timeout 900 python -m pytest tests/test_gpu_cb.py tests/test_gpu_cb64.py -x -q -m gpu 2>&1 | tail -3
