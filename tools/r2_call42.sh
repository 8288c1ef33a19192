set -u
mkdir -p gpurun_out
timeout 100 python tests/dev/ks_bench.py 65536 5 2>&1 | tail -1 | tee gpurun_out/kstc_call42.txt
timeout 200 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | grep -o '"ms": [0-9.]*\|"keyswitch_ms": [0-9.]*' | tr '\n' ' ' | tee -a gpurun_out/kstc_call42.txt
echo
timeout 900 python -m pytest tests/test_gpu_keyswitch.py tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:keyswitch_tc -c 1 -o gpurun_out/ks_r2h -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --gate-only > gpurun_out/ncu_ks_r2h.log 2>&1
tail -2 gpurun_out/ncu_ks_r2h.log
