"""BASELINE config 5: synthetic O×256×256 bilevel learning step (one tv_op_learning_function
evaluation per branch, 5000 inner iterations), images sharded over the ranks, ONE NCCL all-reduce
of [cost, grad] per evaluation.  `torchrun --nproc-per-node N tools/config5.py [O] [iters]`."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bpldenoising_b200 as bp  # noqa: E402
from bpldenoising_b200.parallel import allreduce_costgrad, shard_range  # noqa: E402

O = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
b, c = shard_range(O, world, rank)
t0 = time.perf_counter()
truth, noisy = bp.synthetic_dataset(256, 256, c, seed=20240602 + 1000 * rank)
tgen = time.perf_counter() - t0
ctx = bp.Context([local], 64)
ctx.set_dataset((truth, noisy))
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
cg = torch.zeros(2, dtype=torch.float64, device=dev)
out = {"O": O, "ranks": world, "images_per_rank": c, "iters": iters}
for name, Delta in (("nonreg", 0.1), ("reg", 1e-7)):
    eo = bp.eval_opts(bp.pdps_opts(maxiter=iters))
    for rep in range(2):   # first = warm-up (allocations)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.learn_eval_device(0.1, Delta, cg.data_ptr(), eo, stream=stream.cuda_stream)
        allreduce_costgrad(cg)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[name] = {"ms": float(ms.item()), "cost": float(cg[0].item()), "grad": float(cg[1].item()),
                 "gpixel_iter_per_s_incl_gradient": O * 65536 * iters / float(ms.item()) / 1e6}
if rank == 0:
    print(json.dumps(out))
ctx.close()
if world > 1:
    dist.destroy_process_group()
