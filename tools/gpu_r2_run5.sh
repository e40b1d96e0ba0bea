# Round-2 run 5 (1 GPU): staged copies for pageable caller buffers (e2e.pageable), full parity suite
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -8 gpurun_out/gputests.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_stage.json 2> gpurun_out/bench_stage.err; tail -c 500 gpurun_out/bench_stage.err
BPLTV_HOST_STAGING=0 timeout 900 python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/bench_nostage.json 2> gpurun_out/bench_nostage.err
python - <<'PY'
import json
for n in ('stage','nostage'):
    l=json.loads(open(f'gpurun_out/bench_{n}.json').read().strip().splitlines()[-1])
    print(n, l['value'], l['e2e']['value'], l['e2e']['device_ms'], l['e2e']['pageable'])
PY
