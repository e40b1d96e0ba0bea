"""Host logic of the N>1 path on CPU: image sharding + the single all-reduce of
[cost, grad...], world_size 2 over gloo.  The per-shard numbers come
from the oracle here (this is a test of the sharding/collective plumbing)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_partitions(bp):
    for O in (0, 1, 7, 10, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [bp.shard_range(O, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == O
            pos = 0
            for b, c in spans:
                assert b == min(O, pos) or c == 0
                pos = b + c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= (O + world - 1) // world
    with pytest.raises(ValueError):
        bp.shard_range(4, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from bpldenoising_b200.parallel import allreduce_costgrad, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
    t = np.asfortranarray(z["faces_train_128_10/true"][:24, :24, :5] / 255.0)
    f = np.asfortranarray(z["faces_train_128_10/data"][:24, :24, :5] / 255.0)
    b, c = shard_range(5, world, rank)
    _, cost, g = orc.tv_op_learning_function(0.05, (t[:, :, b:b + c], f[:, :, b:b + c]), 0.1, maxiter=200)
    v = torch.tensor([cost, g], dtype=torch.float64)
    allreduce_costgrad(v)
    if rank == 0:
        np.save(os.path.join(tmp, "sum.npy"), v.numpy())
    dist.destroy_process_group()


def test_gloo_allreduce_of_cost_and_gradient(tmp_path, oracle, datasets):
    import torch.multiprocessing as mp

    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sum.npy")
    t, f = (a[:24, :24, :5].copy(order="F") for a in datasets["faces_train_128_10"])
    _, cost, g = oracle.tv_op_learning_function(0.05, (t, f), 0.1, maxiter=200)
    assert abs(got[0] - cost) <= 1e-13 * cost
    assert abs(got[1] - g) <= 1e-12 * abs(g)
