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
