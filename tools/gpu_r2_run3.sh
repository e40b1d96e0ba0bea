# Round-2 run 3 (1 GPU): parity tests after the guard fix, the resident kernels with the joint projection, sustained
# bench at T = 2 and T = 4 (strict), a fresh ncu capture of the resident kernel.
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -25 gpurun_out/gputests.txt
timeout 600 python tools/time_resident.py 2>&1 | tee gpurun_out/time_resident.txt
timeout 600 python tools/time_sumregs_pdps128.py 2>&1 | tee gpurun_out/time_sumregs_pdps128.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err
BPLTV_TBLOCK_T=4 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_t4.json 2> gpurun_out/bench_t4.err
python - <<'PY'
import json
for n in ('t2','t4'):
    l=json.loads(open(f'gpurun_out/bench_{n}.json').read().strip().splitlines()[-1])
    print(n, l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['kernel'], l['clocks'])
PY
timeout 300 python tools/profile_case.py resident 300 > gpurun_out/plain_resident.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pdps_resident -c 1 -f -o gpurun_out/prof_resident python tools/profile_case.py resident 300 > gpurun_out/ncu_resident.log 2>&1
tail -n 2 gpurun_out/plain_resident.log gpurun_out/ncu_resident.log
du -sh gpurun_out
