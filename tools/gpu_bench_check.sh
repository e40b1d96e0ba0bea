python -m pytest tests/test_gpu_pdps.py -m gpu -x -q 2>&1 | tail -3
for full in 1 0; do echo "FULL=$full"; BPLTV_MARCH_FULL=$full python bench.py --steps 8 --warmup 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value %.1f ms %.1f frac %.3f e2e %.1f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']), d['clocks']['sm_mhz'], d['per_gpu_value_other_arith']); print('  ', {k:(round(v['ms'],1), round(v['learn_run']['seconds'],2), v['learn_run']['evaluations']) for k,v in d['learn_eval'].items()})"; done
