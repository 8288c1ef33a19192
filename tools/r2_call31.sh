set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call31.txt
: > $OUT
for lib in tcd1 tcd2 tcd3 tcd4 tcd8 tcd15; do
  echo "lib=${lib:-default}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 60 python tests/dev/ks_bench.py 128 20 2>&1 | tail -1 >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 60 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 >> $OUT
done
cat $OUT
