set -u
mkdir -p gpurun_out
for lib in "" tools/alt/libtfhe_b200_ksv.so tools/alt/libtfhe_b200_ksn16.so tools/alt/libtfhe_b200_ksvn16.so; do
  echo "lib=${lib:-default}"
  TFHE_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1
  TFHE_B200_LIB=${lib:+$PWD/$lib} timeout 600 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"keyswitch_ms": [0-9.]*'
done | tee gpurun_out/ks_variants_call11.txt
TFHE_B200_LIB=$PWD/tools/alt/libtfhe_b200_ksv.so timeout 600 python -m pytest tests/test_gpu_gate.py tests/test_gpu_params.py tests/test_gpu_cb.py -x -q -m gpu 2>&1 | tail -3
