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


def test_nccl_binding_loads_and_makes_an_id(bp):
    """The library binds NCCL at run time (bpltv_comm_*, include/bpltv.h): the id rank 0 creates needs no GPU."""
    a, b = bp.Context.comm_unique_id(), bp.Context.comm_unique_id()
    assert len(a) == 128 and len(b) == 128 and a != b and any(a)


@pytest.mark.gpu
def test_communicator_of_one_rank_changes_nothing(bp, ctx, oracle, datasets):
    """A one-rank job: learn_eval through the library's ncclAllReduce returns what it returns without a communicator."""
    t, f = (a[:64, :64, :3].copy(order="F") for a in datasets["faces_train_128_10"])
    ctx.set_dataset((t, f))
    eo = bp.eval_opts(bp.pdps_opts(maxiter=300))
    u0, c0, g0 = ctx.learn_eval(0.07, 0.1, eo)
    ctx.comm_init(1, 0, bp.Context.comm_unique_id())
    try:
        u1, c1, g1 = ctx.learn_eval(0.07, 0.1, eo)
    finally:
        ctx.comm_destroy()
    assert np.array_equal(u0, u1) and c0 == c1 and g0 == g1


@pytest.mark.gpu
def test_two_ranks_sum_loss_and_gradient_inside_the_library(bp, oracle, datasets):
    """Two single-device contexts (two GPUs, one host thread each — the shape of a one-process-per-GPU job) join one
    communicator; each holds its shard (shard_range) and each learn_eval returns the WHOLE job's loss and gradient, equal
    to a one-context evaluation of all images to the rounding of the sum.  Scalar and patch λ, TV and sum-of-regularisers."""
    import ctypes
    import threading
    cuda = ctypes.CDLL("libcuda.so.1")
    n = ctypes.c_int(0)
    cuda.cuInit(0); cuda.cuDeviceGetCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    t, f = (a[:64, :64, :5].copy(order="F") for a in datasets["faces_train_128_10"])
    eo = bp.eval_opts(bp.pdps_opts(maxiter=400))
    x = np.array([[0.03, 0.08], [0.05, 0.06]])
    with bp.Context([0], 64) as c:
        c.set_dataset((t, f))
        ref = [c.learn_eval(0.07, 0.1, eo)[1:], c.learn_eval(x, 1e-7, eo)[1:],
               c.sumregs_learn_eval(np.array([0.03, 0.02, 0.01]), 0.1)[1:]]
    uid = bp.Context.comm_unique_id()
    out, err = [None, None], []

    def rank(r):
        try:
            b, cnt = bp.shard_range(5, 2, r)
            with bp.Context([r], 64) as c:
                c.comm_init(2, r, uid)
                c.set_dataset((t[:, :, b:b + cnt].copy(order="F"), f[:, :, b:b + cnt].copy(order="F")))
                out[r] = [c.learn_eval(0.07, 0.1, eo)[1:], c.learn_eval(x, 1e-7, eo)[1:],
                          c.sumregs_learn_eval(np.array([0.03, 0.02, 0.01]), 0.1)[1:]]
        except Exception as e:          # noqa: BLE001 - reported below
            err.append((r, repr(e)))

    th = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    for x_ in th:
        x_.start()
    for x_ in th:
        x_.join(300)
    assert not err, err
    for k in range(3):
        for r in range(2):
            cost, g = out[r][k]
            assert abs(cost - ref[k][0]) <= 1e-13 * ref[k][0]
            assert np.linalg.norm(np.atleast_1d(g) - np.atleast_1d(ref[k][1])) <= 1e-11 * np.linalg.norm(np.atleast_1d(ref[k][1]))
        assert out[0][k][0] == out[1][k][0] and np.array_equal(out[0][k][1], out[1][k][1])     # every rank: the same bits
