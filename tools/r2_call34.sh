set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err
tail -c 600 gpurun_out/bench_r2b_n2.err
cut -c1-400 gpurun_out/bench_r2b_n2.json
