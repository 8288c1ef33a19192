set -u
mkdir -p gpurun_out
OUT=gpurun_out/deferred_call20.txt
: > $OUT
for v in s d s d; do
  echo "variant $v" >> $OUT
  TFHE_B200_BR_VARIANT=$v timeout 100 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
cat $OUT
TFHE_B200_BR_VARIANT=d timeout 600 python -m pytest tests/test_gpu_gate.py -x -q -m gpu 2>&1 | tail -4
