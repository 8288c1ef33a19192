set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call40.txt
: > $OUT
for lib in tcm4s2 tcm4s4 tcm4g1s8 tcm4g1s4; do
  echo "lib=${lib}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 100 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"keyswitch_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
  echo >> $OUT
done
cat $OUT
