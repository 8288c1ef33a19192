set -u
mkdir -p gpurun_out
TFHE_B200_KS=cuda timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests_cuda_ks_r2.txt
