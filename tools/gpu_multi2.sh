set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_sumregs.py -q -k "multi_device" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','scaling','loss')}, 'e2e',d['e2e']['value'], d['clocks'])
PY
