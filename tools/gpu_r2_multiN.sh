# Round-2, N GPUs (default 8): torchrun bench through the library's NCCL all-reduce; config 5 strong scaling
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','scaling','loss')}, 'e2e',d['e2e']['value'], d['clocks'], d['config']['collective'])
print(json.dumps(d.get('config5'))[:1200])
PY
