#!/bin/bash
# Round-2 measurement pass (run under gpurun, one GPU): contract bench, reference arm, ncu launch list of the bench command,
# ncu --set full of the dominant kernels (gate blind rotation, gate key switch, N=2048 blind rotation, private key switch, hp FFT).
# Usage: bash tools/capture_profiles_r2.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
tail -c 600 $OUT/bench_${TAG}_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>> $OUT/bench_${TAG}_n1.err
cat $OUT/bench_${TAG}_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -c "blind_rotate" $OUT/launches_${TAG}.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -c 1 -o $OUT/br_${TAG} -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --gate-only > $OUT/ncu_br_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:keyswitch -c 1 -o $OUT/ks_${TAG} -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --gate-only > $OUT/ncu_ks_${TAG}.log 2>&1
# circuit bootstrap at the BASELINE batch: the N=2048 blind rotation and the private key switch (second keyswitch launch of a step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blind_rotate_kernel|keyswitch" -c 3 -o $OUT/cb_${TAG} -f \
    python tests/dev/bench_cb.py 4096 nohp > $OUT/ncu_cb_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hp_ -c 2 -o $OUT/hp_${TAG} -f \
    python tests/dev/bench_cb.py 0 hponly > $OUT/ncu_hp_${TAG}.log 2>&1
ls -la $OUT/*_${TAG}.ncu-rep
