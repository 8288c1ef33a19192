set -u
mkdir -p gpurun_out
for v in s m M; do
  echo "variant $v"; TFHE_B200_BR_VARIANT=$v timeout 300 python tests/dev/quick_bench.py 65536 2>&1 | tail -2
done | tee gpurun_out/variants_call9.txt
