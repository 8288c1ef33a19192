set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
echo rc=$?; tail -c 1200 gpurun_out/bench_r2_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
for k in ('circuit_bootstrap','hp_fft','adder32','strong'):
    print(k, json.dumps(d.get(k))[:700])
P
timeout 300 python -m pytest tests/test_gpu_hp.py -q -m gpu 2>&1 | tail -3
