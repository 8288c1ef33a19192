set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/tests_call19.txt
timeout 900 python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err
tail -c 400 gpurun_out/bench_r2b_n1.err
cut -c1-600 gpurun_out/bench_r2b_n1.json
