python -m pytest tests/test_gpu_pdps.py -m gpu -x -q 2>&1 | tail -3
for st in 0 2 3 4; do for mb in 3 4; do echo "STAGES=$st MINB=$mb"; BPLTV_MARCH_STAGES=$st BPLTV_MARCH_MINB=$mb python bench.py --steps 8 --warmup 3 --no-extras 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value %.1f ms %.1f frac %.3f e2e-dev-pdps %.1f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['device_ms']['pdps']), d['clocks']['sm_mhz'])"; done; done
for st in 0 3; do BPLTV_MARCH_STAGES=$st python bench.py --steps 8 --warmup 3 --no-extras --arith fast 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FAST value %.1f ms %.1f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"; done
