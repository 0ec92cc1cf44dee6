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
