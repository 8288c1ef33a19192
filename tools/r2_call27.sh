set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call27.txt
: > $OUT
for B in 128 2048 18944; do timeout 100 python tests/dev/ks_bench.py $B 20 2>&1 | tail -1 >> $OUT; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:keyswitch_tc -c 6 --csv --log-file gpurun_out/kstc_launches.csv python tests/dev/bench_cb.py 4096 nohp > /dev/null 2>&1
grep -o '"keyswitch_tc_kernel[^"]*".*' gpurun_out/kstc_launches.csv | cut -c1-200 >> $OUT
cat $OUT
