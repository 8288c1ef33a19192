set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:keyswitch_tc -c 1 -o gpurun_out/kstc_r2c -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --gate-only > gpurun_out/ncu_kstc_r2c.log 2>&1
tail -3 gpurun_out/ncu_kstc_r2c.log
ls -la gpurun_out/kstc_r2c.ncu-rep
