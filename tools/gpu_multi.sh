N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_gradient.py -m gpu -x -q -k multi_device 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 400 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','scaling','loss')}, 'frac',d['roofline']['frac'], 'e2e',d['e2e']['value'], d['clocks'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-200
