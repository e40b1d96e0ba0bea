# BASELINE config 5 on N GPUs (weak: 128 images 256x256 per GPU) + the bench line at N GPUs
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
O=$((128 * N))
if [ "$N" = "1" ]; then
  python tools/config5.py $O 5000 > gpurun_out/config5_n1.json 2> gpurun_out/config5_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/config5.py $O 5000 > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  tail -c 300 gpurun_out/bench_n$N.err
fi
tail -c 300 gpurun_out/config5_n$N.err; tail -n 1 gpurun_out/config5_n$N.json
python - $N <<'PY'
import json,sys,os
p=f'gpurun_out/bench_n{sys.argv[1]}.json'
if os.path.exists(p):
    d=json.loads(open(p).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('n_gpus','value','ms_per_step','scaling','loss')}, 'frac',d['roofline']['frac'], 'e2e',d['e2e']['value'], d['clocks'])
PY
