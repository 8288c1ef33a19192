set -u
mkdir -p gpurun_out
OUT=gpurun_out/kspair_call16.txt
: > $OUT
for pair in 1 0; do
  echo "KS_PAIR=$pair" >> $OUT
  TFHE_B200_KS_PAIR=$pair timeout 120 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
cat $OUT
timeout 900 python -m pytest tests/test_gpu_gate.py tests/test_gpu_params.py tests/test_gpu_circuit.py tests/test_gpu_keygen.py tests/test_capi_and_host.py -x -q -m gpu 2>&1 | tail -3
