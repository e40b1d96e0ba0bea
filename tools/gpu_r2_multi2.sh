# Round-2, 2 GPUs: the library's own NCCL all-reduce (bpltv_comm_init) — threads test, torchrun bench at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 900 python -m pytest tests -m gpu -q -k "multi_device or two_ranks or communicator" 2>&1 | tail -6
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','scaling','loss')}, 'e2e',d['e2e']['value'], d['clocks'])
print(json.dumps(d.get('config5'))[:1500])
PY
