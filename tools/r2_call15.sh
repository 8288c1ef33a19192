set -u
mkdir -p gpurun_out
OUT=gpurun_out/ab_call15.txt
: > $OUT
for lib in "" head nofd "" head; do
  echo "lib=${lib:-default}" >> $OUT
  TFHE_B200_LIB=${lib:+$PWD/tools/alt/libtfhe_b200_$lib.so} timeout 120 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
timeout 300 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | cut -c1-400 >> $OUT
cat $OUT
timeout 900 python -m pytest tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:keyswitch -c 1 -o gpurun_out/ks_r2b -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --gate-only > gpurun_out/ncu_ks_r2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"keyswitch_kernel" -c 2 -o gpurun_out/privks_r2b -f python tests/dev/bench_cb.py 4096 nohp > gpurun_out/ncu_privks_r2b.log 2>&1
ls -la gpurun_out/*r2b*
