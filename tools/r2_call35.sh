set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call35.txt
: > $OUT
for lib in "" tcm1 tcm3 tcm2cl2 tcm3cl2; do
  echo "lib=${lib:-default (2 MMA warps)}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 100 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
  echo >> $OUT
done
cat $OUT
timeout 900 python -m pytest tests/test_gpu_keyswitch.py tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -4
