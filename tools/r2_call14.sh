set -u
mkdir -p gpurun_out
OUT=gpurun_out/lock_call14.txt
: > $OUT
for lib in "" pl1 pl3 pl99 pl127; do
  echo "lib=${lib:-default}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
cat $OUT
timeout 900 python -m pytest tests/test_gpu_gate.py tests/test_gpu_trgsw.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -3
