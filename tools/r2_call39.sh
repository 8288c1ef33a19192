set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call39.txt
: > $OUT
timeout 100 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 >> $OUT
timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*' | tr '\n' ' ' >> $OUT
echo >> $OUT
cat $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests_call39.txt
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tests/dev/sanitize_run.py gate,cb > gpurun_out/san_${tool}_r2c.txt 2>&1
  tail -2 gpurun_out/san_${tool}_r2c.txt
done
