set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_k20.json 2> gpurun_out/bench_k20.err; cut -c1-400 gpurun_out/bench_k20.json; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_k20.json')); print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['roofline']['frac'], d['e2e'])
PY
python bench.py --steps 20 --warmup 3 --no-extras --arith fast > gpurun_out/bench_k20_fast.json 2> gpurun_out/bench_k20_fast.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_k20_fast.json')); print('FAST',{k:d[k] for k in ('value','ms_per_step','clocks')}, d['roofline']['frac'], d['e2e'])
PY
# ncu: launch list of the bench command, then full capture of the top kernel on a short case
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/profile_case.py pdps 12 > gpurun_out/plain_pdps.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pdps_march -s 6 -c 3 -o gpurun_out/prof_march python tools/profile_case.py pdps 12 > gpurun_out/ncu_pdps.log 2>&1
python tools/profile_case.py grad 50 > gpurun_out/plain_grad.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grad_ -c 12 -o gpurun_out/prof_grad python tools/profile_case.py grad 50 > gpurun_out/ncu_grad.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_pdps.log gpurun_out/ncu_grad.log gpurun_out/ncu_bench.log
