set -u
mkdir -p gpurun_out
OUT=gpurun_out/lock_call13.txt
: > $OUT
for m in 0 1 2 3 7 15 51 67 99 115 127 255 112 48; do
  echo "lockmask $m" >> $OUT
  TFHE_B200_BR_LOCK=$m timeout 200 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 >> $OUT
done
cat $OUT
TFHE_B200_BR_LOCK=127 timeout 600 python -m pytest tests/test_gpu_gate.py -x -q -m gpu 2>&1 | tail -3
