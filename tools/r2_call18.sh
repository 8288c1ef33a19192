set -u
mkdir -p gpurun_out
OUT=gpurun_out/kstc_call18.txt
: > $OUT
timeout 300 python -m pytest tests/test_gpu_gate.py -x -q -m gpu -k "keyswitch or KeySwitch or ks" 2>&1 | tail -5 | tee -a $OUT
timeout 120 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 | tee -a $OUT
TFHE_B200_KS=cuda timeout 120 python tests/dev/quick_bench.py 65536 2>&1 | tail -2 | head -1 | tee -a $OUT
timeout 600 python -m pytest tests/test_gpu_gate.py tests/test_gpu_cb.py tests/test_gpu_params.py -x -q -m gpu 2>&1 | tail -5 | tee -a $OUT
timeout 300 python tests/dev/bench_cb.py 4096 nohp 2>&1 | grep circuit_bootstrap | cut -c1-420 | tee -a $OUT
