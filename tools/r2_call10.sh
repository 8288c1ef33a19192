set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gate.py tests/test_gpu_params.py tests/test_gpu_cb.py -x -q -m gpu 2>&1 | tail -4
for lib in "" tools/alt/libtfhe_b200_ksu0.so tools/alt/libtfhe_b200_ksu2.so; do
  echo "lib=${lib:-default(KS_UNIFORM=1)}"
  TFHE_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1
  TFHE_B200_LIB=${lib:+$PWD/$lib} timeout 600 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | cut -c1-330
done | tee gpurun_out/ks_variants_call10.txt
