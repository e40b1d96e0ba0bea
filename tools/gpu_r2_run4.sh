# Round-2 run 4 (1 GPU): default depth 4 of the temporally blocked kernel (fp64) — pins, bench with extras, launch list,
# full ncu capture of the dispatched kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pdps.py tests/test_gpu_sweep.py tests/test_parallel.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2.json 2>> gpurun_out/bench_r2.err; tail -c 800 gpurun_out/bench_r2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 300 python tools/profile_case.py tblock 16 > gpurun_out/plain_tblock.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 1 -c 2 -f -o gpurun_out/prof_tblock_t4 python tools/profile_case.py tblock 16 > gpurun_out/ncu_tblock.log 2>&1
tail -n 2 gpurun_out/plain_tblock.log gpurun_out/ncu_tblock.log
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['e2e'].get('pageable',{}).get('value'), l['roofline']['kernel'], l['roofline']['frac'], l['clocks'])
PY
du -sh gpurun_out
