# Round-2 run 10 (1 GPU): async halo exchange in kernel B and in the sum-of-regularisers resident solve: parity + times,
# fresh ncu capture of kernel B, full bench
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -8 gpurun_out/gputests.txt
timeout 600 python tools/time_resident.py 2>&1 | tee gpurun_out/time_resident.txt
timeout 600 python tools/time_sumregs_pdps128.py 2>&1 | tee gpurun_out/time_sumregs_pdps128.txt
timeout 300 python tools/profile_case.py resident 300 > gpurun_out/plain_resident.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_resident -c 1 -f -o gpurun_out/prof_resident python tools/profile_case.py resident 300 > gpurun_out/ncu_resident.log 2>&1
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 500 gpurun_out/bench_r2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['e2e'].get('pageable',{}).get('value'), l['roofline']['kernel'], l['roofline']['frac'], l['roofline']['traffic'], l['clocks'])
le=l.get('learn_eval',{})
for k,v in le.items():
    if isinstance(v,dict) and 'ms' in v: print(k, {kk:v[kk] for kk in ('ms','ms_pdps','ms_gradient')}, v.get('learn_run',{}).get('seconds'))
print(json.dumps(l.get('config5'))[:900])
PY
