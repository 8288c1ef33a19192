set -u
mkdir -p gpurun_out
(cd tools && ./microbench) > gpurun_out/microbench_r2.txt 2>&1
tail -12 gpurun_out/microbench_r2.txt
for tool in memcheck racecheck synccheck; do
  SAN_VARIANTS=default,keytm timeout 500 compute-sanitizer --tool $tool python tests/dev/sanitize_run.py > gpurun_out/san_${tool}_r2.txt 2>&1
  echo "$tool rc=$?"; tail -5 gpurun_out/san_${tool}_r2.txt
done
timeout 700 ncu --set full --clock-control none --import-source on -k regex:keyswitch_kernel -c 3 -o gpurun_out/privks_r2 -f python tests/dev/bench_cb.py 1024 > gpurun_out/ncu_privks_r2.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_privks_r2.log
