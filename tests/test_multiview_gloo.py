"""CPU tests (gloo, world_size 2) of the view-sharding host logic in splatco_b200/multiview.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from splatco_b200.multiview import owner_of_view, shard_views


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        for mv in (1, 4, 8, 11):
            seen = []
            for r in range(world):
                v = shard_views(mv, r, world)
                assert all(owner_of_view(i, world) == r for i in v)
                seen += v
            assert sorted(seen) == list(range(mv))
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splatco_b200.multiview import GradBucket, allreduce_count, broadcast_last_view_stats
        torch.manual_seed(0)
        # replicated "model": per-anchor rows, a plane, an MLP weight, one frozen tensor, one never touched
        params = [torch.nn.Parameter(torch.zeros(50, 3)), torch.nn.Parameter(torch.zeros(1, 5, 8, 8)),
                  torch.nn.Parameter(torch.zeros(32, 99)), torch.nn.Parameter(torch.zeros(4), requires_grad=False),
                  torch.nn.Parameter(torch.zeros(7))]
        mv = 4
        targets = [torch.full((), float(i + 1)) for i in range(mv)]
        # each view contributes grad = (i+1) to params 0..2; param 4 gets a grad only from view 3 (rank 1)
        loss = 0
        for i in shard_views(mv, rank, world):
            loss = loss + targets[i] * (params[0].sum() + params[1].sum() + params[2].sum())
            if i == 3:
                loss = loss + 2.0 * params[4].sum()
        loss.backward()
        bucket = GradBucket(params)
        assert bucket.total == 50 * 3 + 5 * 64 + 32 * 99 + 7
        bucket.allreduce()
        expect = float(sum(range(1, mv + 1)))
        ok = all(torch.allclose(p.grad, torch.full_like(p, expect)) for p in params[:3])
        ok &= torch.allclose(params[4].grad, torch.full_like(params[4], 2.0))
        ok &= params[3].grad is None
        # last-view statistics are broadcast from the owner of view mv-1 (rank 1 when world = 2)
        stats = torch.full((5,), float(rank + 10))
        broadcast_last_view_stats([stats], mv)
        ok &= torch.allclose(stats, torch.full((5,), float(owner_of_view(mv - 1, world) + 10)))
        ok &= allreduce_count(3 + rank, "cpu") == sum(3 + r for r in range(world))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _pair_loss(gens, reals):
    """Stand-in for splatco_b200.loss.multiview_consistency_loss (the CUDA kernel) with the same pair structure."""
    total = 0
    for i in range(len(gens)):
        for j in range(i + 1, len(gens)):
            total = total + (0.5 + 0.1 * i + 0.01 * j) * ((reals[i] - reals[j]) - (gens[i] - gens[j])).abs().mean()
    return total


def _render(theta, i):
    """A differentiable stand-in "renderer": view i's image as a function of the replicated parameter."""
    g = torch.Generator().manual_seed(100 + i)
    basis = torch.rand(3, 6, 7, generator=g)
    return torch.sin(theta[0] * basis * (i + 1)) + theta[1] * basis


def _worker_consistency(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splatco_b200.multiview import GradBucket, gather_view_images, sharded_consistency_loss
        mv = 5                                                     # uneven: rank 0 owns 3 views, rank 1 owns 2
        g = torch.Generator().manual_seed(7)
        reals = [torch.rand(3, 6, 7, generator=g) for _ in range(mv)]
        theta = torch.nn.Parameter(torch.tensor([0.7, -0.3]))
        # single-process answer
        full = _pair_loss([_render(theta, i) for i in range(mv)], reals)
        (want_grad,) = torch.autograd.grad(full, theta)
        # sharded: own views only, all pairs through the gathered images
        local = [_render(theta, i) for i in shard_views(mv, rank, world)]
        gathered = gather_view_images(local, mv, rank, world)
        ok = len(gathered) == mv
        for i in range(mv):
            ok &= torch.allclose(gathered[i], _render(theta, i).detach())
            ok &= gathered[i].requires_grad == (owner_of_view(i, world) == rank)
        loss = sharded_consistency_loss(local, reals, mv, rank, world, loss_fn=_pair_loss)
        ok &= torch.allclose(loss.detach(), full.detach())          # the full sum on every rank
        loss.backward()
        GradBucket([theta]).allreduce()
        ok &= torch.allclose(theta.grad, want_grad, rtol=1e-5, atol=1e-7)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_consistency_loss_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_consistency, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
