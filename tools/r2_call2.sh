set -u
mkdir -p gpurun_out
(cd tools && ./microbench | tail -9) > gpurun_out/microbench_r2b.txt 2>&1; cat gpurun_out/microbench_r2b.txt
timeout 900 python -m pytest tests/test_gpu_gate.py tests/test_gpu_params.py tests/test_gpu_cb.py tests/test_gpu_trgsw.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/tests_call2.txt
for v in s 2 3 4; do
  echo "variant $v"; TFHE_B200_BR_VARIANT=$v timeout 300 python tests/dev/quick_bench.py 65536 2>&1 | tail -2
done | tee gpurun_out/variants_call2.txt
for v in s 2 3; do
  echo "variant $v"; TFHE_B200_BR_VARIANT=$v timeout 600 python tests/dev/bench_cb.py 4096 2>&1 | grep circuit_bootstrap | cut -c1-400
done | tee gpurun_out/variants_cb_call2.txt
for tool in memcheck racecheck synccheck; do
  SAN_VARIANTS=default,s timeout 500 compute-sanitizer --tool $tool python tests/dev/sanitize_run.py > gpurun_out/san_${tool}_r2.txt 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/san_${tool}_r2.txt
done
