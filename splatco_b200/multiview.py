"""View-sharded multi-GPU execution of the `--mv` multi-view batch (new work — the
reference is single-process and runs the mv views sequentially, train.py:171-240).

One process per GPU (torchrun), parameters replicated.  Rank r renders the views
{i : i mod world == r} of the iteration's camera list; after the local backward the per-parameter
gradients are summed with ONE NCCL all-reduce over a flat fp32 bucket (per-anchor rows + planes + MLP
weights), and the statistics the reference takes from the last view only (train.py:266,
scene/gaussian_model.py:761-782) are broadcast from the rank that owns view mv-1.  BatchNorm batch
statistics stay per view, as in the reference (no SyncBN).  The backend is whatever the process
group was created with: "nccl" on GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Indices of the iteration's views rendered by `rank` (round-robin)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return [i for i in range(num_views) if i % world == rank]


def owner_of_view(view: int, world: int) -> int:
    return view % world


class GradBucket:
    """Flat fp32 bucket over a fixed parameter list; `allreduce()` sums .grad across ranks in one
    collective and scatters the result back (parameters without a grad contribute zeros, so every
    rank issues the same collective even if it rendered no view that touched them)."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)

    def nbytes(self) -> int:
        return self.total * 4

    def pack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            seg = self.flat[off:off + n]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))
            off += n

    def unpack(self):
        off = 0
        for p, n in zip(self.params, self.sizes):
            seg = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = seg.clone()
            else:
                p.grad.copy_(seg)
            off += n

    def allreduce(self, group=None):
        """Sum .grad over ranks.  Gradients that already live in the flat buffers of the last backward pass
        (everything the decode / TriPlaneAttention kernels accumulate: anchors, planes, MLP weights -- see
        _gradacc.py) are all-reduced IN PLACE, one collective per buffer and no copies; whatever else has a
        gradient goes through the packed bucket."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        from . import _gradacc
        dev = self.params[0].device if self.params else torch.device("cpu")
        flats = _gradacc.last_pass_buffers(dev) if dev.type == "cuda" else []
        if dev.type == "cuda":
            # every rank must issue the same collectives: agree on the buffer layout (a rank that rendered no view of
            # this iteration has none), fall back to the packed bucket everywhere otherwise
            sig = torch.tensor([len(flats), sum(f.numel() for f in flats)], dtype=torch.int64, device=dev)
            lo, hi = sig.clone(), sig.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if not torch.equal(lo, hi):
                flats = []
        owned = {f.untyped_storage().data_ptr() for f in flats}
        rest = [p for p in self.params if p.grad is None or p.grad.untyped_storage().data_ptr() not in owned]
        for f in flats:
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group)
        if not flats:
            self.pack()
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.unpack()
        elif rest:
            sizes = [p.numel() for p in rest]
            small = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            off = 0
            for p, n in zip(rest, sizes):
                if p.grad is not None:
                    small[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            dist.all_reduce(small, op=dist.ReduceOp.SUM, group=group)
            off = 0
            for p, n in zip(rest, sizes):
                seg = small[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = seg.clone()
                else:
                    p.grad.copy_(seg)
                off += n


def broadcast_last_view_stats(tensors: Iterable[torch.Tensor], num_views: int, group=None):
    """training_statis consumes only the LAST view's tensors (train.py:266): the owner of view mv-1
    broadcasts its accumulators so every replica applies the same densification statistics."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    src = owner_of_view(num_views - 1, dist.get_world_size(group))
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def allreduce_count(value: int, device, group=None) -> int:
    """Sum of an integer count (visibility / pruning counters) across ranks; exact in int64."""
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def gather_view_images(local_images, num_views: int, rank: int, world: int, group=None) -> List[torch.Tensor]:
    """All views' rendered images on every rank (SURVEY §8e collective (2)): `local_images[k]` is the image of this
    rank's k-th view (global index shard_views(num_views, rank, world)[k]).  One all_gather of the padded local stack;
    the result is in global view order, remote images are plain (detached) tensors, the local ones are returned as
    given (so autograd still reaches them).  All views must share one [C,H,W] shape."""
    mine = shard_views(num_views, rank, world)
    if len(local_images) != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} of {num_views} views but passed {len(local_images)} images")
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return list(local_images)
    slots = (num_views + world - 1) // world
    if not local_images:
        raise ValueError("gather_view_images: every rank must own at least one view (num_views >= world)")
    ref = local_images[0]
    shape = torch.tensor(list(ref.shape), dtype=torch.int64, device=ref.device)
    hi = shape.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    lo = shape.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    if not torch.equal(lo, hi):
        raise NotImplementedError("gather_view_images: views of different sizes across ranks are not supported")
    stack = torch.zeros((slots,) + tuple(ref.shape), dtype=ref.dtype, device=ref.device)
    for k, img in enumerate(local_images):
        stack[k].copy_(img.detach())
    out = torch.empty((world,) + tuple(stack.shape), dtype=ref.dtype, device=ref.device)
    dist.all_gather(list(out.unbind(0)), stack, group=group)        # list form: supported by both nccl and gloo
    images = []
    for i in range(num_views):
        r, k = owner_of_view(i, world), i // world
        images.append(local_images[k] if r == rank else out[r, k])
    return images


def sharded_consistency_loss(local_gen, all_real, num_views: int, rank: int, world: int, loss_fn=None, group=None):
    """Cross-view consistency term of the mv batch (train.py:199-216) when the views are sharded: every rank gathers the
    other ranks' rendered images (detached — exact, d/d gen_i treats gen_j as a constant), evaluates ALL pairs and
    differentiates w.r.t. its own images only.  The returned value is the full sum over pairs on every rank (do not
    sum it over ranks when logging); after each rank's backward the gradient all-reduce adds every pair's two halves
    exactly once.  `all_real`: the ground-truth images of all num_views views (the camera list is replicated,
    SURVEY §8e); `loss_fn(gens, reals)` defaults to splatco_b200.loss.multiview_consistency_loss."""
    if loss_fn is None:
        from .loss import multiview_consistency_loss as loss_fn
    gens = gather_view_images(local_gen, num_views, rank, world, group)
    return loss_fn(gens, list(all_real))


def render_views_sharded(cams, pc, pipe, bg, loss_fn, rank: int, world: int, bucket: GradBucket = None, group=None):
    """One iteration's render work for the views owned by `rank`: prefilter -> render -> loss for each,
    ONE backward over the summed loss (as train.py:240), then the gradient all-reduce.
    Returns (local_loss_sum tensor, list of render packages)."""
    from .gaussian_renderer import prefilter_voxel, render
    total = None
    pkgs = []
    for i in shard_views(len(cams), rank, world):
        cam = cams[i]
        vm = prefilter_voxel(cam, pc, pipe, bg)
        pkg = render(cam, pc, pipe, bg, visible_mask=vm, retain_grad=True)
        loss = loss_fn(i, pkg)
        total = loss if total is None else total + loss
        pkgs.append(pkg)
    if total is not None:
        total.backward()
    if bucket is not None:
        bucket.allreduce(group)
    return total, pkgs
