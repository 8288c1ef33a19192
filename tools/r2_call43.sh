set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2h_n2.json 2> gpurun_out/bench_r2h_n2.err
tail -c 300 gpurun_out/bench_r2h_n2.err
cut -c1-300 gpurun_out/bench_r2h_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err
cut -c1-300 gpurun_out/bench_r2h_n1.json
