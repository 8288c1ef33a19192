set -u
mkdir -p gpurun_out
for c in 0 8 0 8 16 6; do echo "chunks=$c"; TFHE_B200_HOST_CHUNKS=$c timeout 200 python bench.py --gate-only --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; done | tee gpurun_out/e2e_call46.txt
