timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_full.json 2> gpurun_out/n2_full.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_full.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'scaling', d['scaling'], 'weak', (d.get('weak_scaling') or {}).get('value'), 'e2e', d['e2e']['value'])
for k,v in d['secondary'].items(): print(' ', k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
