set -u
mkdir -p gpurun_out
OUT=gpurun_out/brv_call49.txt
: > $OUT
for lib in "" v1 v2 v3 v4 v5 ""; do
  echo "lib=${lib:-default}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 100 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"blind_rotate_ms": [0-9.]*' >> $OUT
done
cat $OUT
