set -x
mkdir -p gpurun_out
timeout 900 python tools/time_grad_nd.py 2>&1 | tee gpurun_out/time_grad_nd.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['e2e'].get('pageable'))
print(json.dumps(l.get('config5'),indent=1))
print(json.dumps(l.get('multi_device_context')))
le=l.get('learn_eval',{})
for k,v in le.items():
    if isinstance(v,dict) and 'ms' in v: print(k, {kk:v[kk] for kk in ('ms','ms_pdps','ms_gradient')}, v.get('learn_run',{}).get('seconds'))
print(l['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_pdps.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -5
