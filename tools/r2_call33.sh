set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_keyswitch.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/tests_call33.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/dev/sanitize_run.py gate,cb > gpurun_out/san_${tool}_r2b.txt 2>&1
  tail -4 gpurun_out/san_${tool}_r2b.txt
done
