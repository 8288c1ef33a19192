set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/tests_call6.txt
cat gpurun_out/exact_vs_fp64.json
