set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/tests_call4.txt
timeout 900 python bench.py > gpurun_out/bench_call4.json 2> gpurun_out/bench_call4.err
tail -c 1500 gpurun_out/bench_call4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_call4.json'))
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
for k in ('circuit_bootstrap','hp_fft','adder32','strong'):
    print(k, json.dumps(d.get(k))[:1500])
P
