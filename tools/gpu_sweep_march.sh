python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for mb in 2 3 4; do for pf in 0 2 4; do echo "MINB=$mb PREFETCH=$pf"; BPLTV_MARCH_MINB=$mb BPLTV_MARCH_PREFETCH=$pf python bench.py --steps 12 --warmup 3 --no-extras 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value %.1f ms %.1f frac %.3f clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['clocks']))"; done; done
BPLTV_MARCH_VEC=4 python bench.py --steps 12 --warmup 3 --no-extras 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VEC4 value %.1f ms %.1f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
python bench.py --steps 12 --warmup 3 --arith fast 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FAST value %.1f ms %.1f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'])); print(json.dumps(d['learn_eval']))"
