"""Image sharding across ranks (one process per GPU) and the single collective of
the path: one all-reduce of [cost, grad...] per learning-function evaluation
(SURVEY §8e).  Every image is an independent lower-level and adjoint solve — the
reference just loops `for i = 1:O` and sums
(/root/reference/src/TVLearningFunctionVec.jl:72-83, :163-175) — so there is no
data-path exchange.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(O: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, begin+count) of the O images owned by `rank`;
    blocks of ceil(O/world) images, identical to libbpltv's per-device sharding."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (O + world - 1) // world
    begin = min(O, rank * per)
    return begin, min(O, begin + per) - begin


def allreduce_costgrad(costgrad, group=None):
    """Sum the [cost, grad...] vector over ranks (torch.distributed; NCCL on GPUs,
    gloo on CPU).  `costgrad` is a torch tensor modified in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(costgrad, op=dist.ReduceOp.SUM, group=group)
    return costgrad


def join_job(ctx, group=None):
    """One process per GPU under torchrun: attach `ctx` (a single-device Context holding this rank's shard) to the job's
    NCCL communicator INSIDE libbpltv (bpltv_comm_init).  torch.distributed is only the courier of the 128-byte id
    (a broadcast from rank 0); afterwards every learn_eval of `ctx` returns the whole job's loss and gradient, summed by
    the library's own ncclAllReduce, and `allreduce_costgrad` is not needed."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [None]
    if rank == 0:
        try:
            box[0] = ctx.comm_unique_id()
        except Exception as e:      # noqa: BLE001 - the other ranks are waiting in the broadcast: tell them
            box[0] = ("error", f"{type(e).__name__}: {e}")
    dist.broadcast_object_list(box, src=0, group=group)
    if not isinstance(box[0], (bytes, bytearray)):
        raise RuntimeError(f"rank 0 could not create the NCCL id: {box[0]}")
    ctx.comm_init(world, rank, bytes(box[0]))
    return world, rank
