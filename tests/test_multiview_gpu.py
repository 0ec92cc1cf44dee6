"""View-sharded gradients on the GPU path: two processes (gloo, both on cuda:0 -- the test box has one
GPU; NCCL needs one device per rank, the host logic and the in-place all-reduce of the shared gradient buffers are
the same) each render two of four views, GradBucket.allreduce() sums; the result must equal one process rendering
all four views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(views, allreduce):
    from types import SimpleNamespace
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.model import AnchorModel
    from splatco_b200.multiview import GradBucket
    from splatco_b200.synthetic import ring_cameras
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    pc = AnchorModel(2500, plane_size=128, num_channels=15, device="cuda", seed=4, scale_factor=1.0)
    pc.feat_planes._feat.activate_level = 2
    pc.feat_planes.Q0 = 0.0
    pc.train()
    W, H = 144, 96
    cams = [c.to("cuda") for c in ring_cameras(4, W, H, radius=3.0)]
    bg = torch.ones(3, device="cuda")
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).cuda() for i in range(4)]
    params = [p for p in pc.parameters() if p.requires_grad]
    total = None
    for i in views:
        vm = prefilter_voxel(cams[i], pc, pipe, bg)
        pkg = render(cams[i], pc, pipe, bg, visible_mask=vm, retain_grad=True)
        l = (pkg["render"] - gts[i]).abs().mean() + 0.01 * pkg["scaling"].prod(dim=1).mean()
        total = l if total is None else total + l
    total.backward()
    if allreduce:
        GradBucket(params).allreduce()
    return [None if p.grad is None else p.grad.detach().cpu().clone() for p in params]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splatco_b200.multiview import shard_views
        g = _grads(shard_views(4, rank, world), allreduce=True)
        if rank == 0:
            torch.save(g, out)
    finally:
        dist.destroy_process_group()


def test_two_ranks_sum_to_the_single_process_gradients(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    want = _grads(range(4), allreduce=False)
    assert len(got) == len(want)
    n = 0
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if a is None:
            continue
        scale = max(b.abs().max().item(), 1e-12)
        assert (a - b).abs().max().item() <= 2e-4 * scale + 1e-9, (tuple(b.shape), (a - b).abs().max().item(), scale)
        n += 1
    assert n >= 40


# ---- cross-view consistency term, sharded (SURVEY §8e collective (2)) ---------------------------------------------
def _images(theta, idx, H=40, W=56):
    """Differentiable stand-in renders of views idx (functions of one replicated parameter) and all ground truths."""
    g = torch.Generator().manual_seed(31)
    base = torch.rand(3, H, W, generator=g)
    reals, basis = [], []
    for i in range(4):
        reals.append(((base + 0.03 * torch.randn(3, H, W, generator=g)).clamp(0, 1) if i < 3 else torch.rand(3, H, W, generator=g)).cuda())
        basis.append(torch.rand(3, H, W, generator=g).cuda())
    gens = [(reals[i] + theta[0] * basis[i] + theta[1] * torch.sin(7.0 * basis[i])).clamp(0, 1) for i in idx]
    return gens, reals


def _worker_consistency(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from splatco_b200.multiview import GradBucket, shard_views, sharded_consistency_loss
        theta = torch.nn.Parameter(torch.tensor([0.11, -0.04], device="cuda"))
        gens, reals = _images(theta, shard_views(4, rank, world))
        loss = sharded_consistency_loss(gens, reals, 4, rank, world)          # the CUDA kernel, all pairs, own views differentiated
        loss.backward()
        GradBucket([theta]).allreduce()
        if rank == 0:
            torch.save((loss.detach().cpu(), theta.grad.detach().cpu()), out)
    finally:
        dist.destroy_process_group()


def test_sharded_consistency_loss_matches_single_process(tmp_path):
    from splatco_b200.loss import multiview_consistency_loss
    out = str(tmp_path / "c.pt")
    mp.spawn(_worker_consistency, args=(2, _free_port(), out), nprocs=2, join=True)
    got_loss, got_grad = torch.load(out)
    theta = torch.nn.Parameter(torch.tensor([0.11, -0.04], device="cuda"))
    gens, reals = _images(theta, range(4))
    want = multiview_consistency_loss(gens, reals)
    want.backward()
    assert want.item() > 0
    assert abs(got_loss.item() - want.item()) <= 1e-6 * abs(want.item())
    assert torch.allclose(got_grad, theta.grad.cpu(), rtol=1e-4, atol=1e-9)
