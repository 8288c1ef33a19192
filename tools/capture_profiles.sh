#!/bin/bash
# Round-end measurement pass (run under gpurun, one GPU): contract bench, ncu launch list, ncu --set full of the two kernels.
# Usage: bash tools/capture_profiles.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
tail -c 3000 $OUT/bench_${TAG}_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>> $OUT/bench_${TAG}_n1.err
cat $OUT/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -c blind_rotate $OUT/launches_${TAG}.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate -c 1 -o $OUT/br_${TAG} -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_br_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:keyswitch -c 1 -o $OUT/ks_${TAG} -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_ks_${TAG}.log 2>&1
ls -la $OUT/*.ncu-rep
