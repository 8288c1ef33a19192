set -u
mkdir -p gpurun_out
TFHE_B200_BR_VARIANT=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -o gpurun_out/br_r2_f2k3 -f python tests/dev/quick_bench.py 65536 > gpurun_out/ncu_br_r2_f2k3.log 2>&1
tail -3 gpurun_out/ncu_br_r2_f2k3.log
