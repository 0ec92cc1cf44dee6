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
    """Sums the parameter gradients of a replicated model across ranks.

    The gradients the decode / TriPlaneAttention kernels accumulate (per-anchor rows, planes, MLP weights) live in ONE
    flat zero-filled fp32 buffer per backward pass (`_gradacc`): that buffer is all-reduced IN PLACE with one collective
    and no copies.  Whatever else carries a gradient travels in one small packed bucket.

    No per-step agreement round: replicas run the same program on the same parameter set, so the layout (number and
    sizes of the flat buffers, which other parameters have gradients) is identical on every rank as long as every rank
    rendered at least one view of the iteration -- the precondition of `render_views_sharded` (num_views >= world).
    The layout is verified across ranks when it is first seen and whenever the LOCAL layout changes (densification
    grows the anchors on every rank at once); SPLATCO_CHECK_ALLREDUCE=1 verifies it on every call.

    `params`: a sequence of tensors or a callable returning the current sequence (densification replaces the
    per-anchor Parameters: pass `lambda: [g["params"][0] for g in optimizer.param_groups]`)."""

    def __init__(self, params):
        self._source = params if callable(params) else (lambda p=list(params): p)
        self._agreed = None
        self._small = None
        self.last_bytes = 0

    @property
    def params(self):
        return [p for p in self._source() if p.requires_grad]

    @property
    def total(self) -> int:
        """Number of gradient elements of the current parameter list."""
        return sum(p.numel() for p in self.params)

    def nbytes(self) -> int:
        """Bytes the last allreduce() moved per rank (before the first call: 4 bytes per parameter element)."""
        return self.last_bytes or 4 * self.total

    # ---- packed path (CPU / gloo tests, and parameters whose gradient is not in the flat buffers) ------------------
    @staticmethod
    def _pack(params, flat):
        off = 0
        for p in params:
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n

    @staticmethod
    def _unpack(params, flat):
        off = 0
        for p in params:
            n = p.numel()
            seg = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = seg.clone()
            else:
                p.grad.copy_(seg)
            off += n

    def allreduce(self, group=None, async_op=False):
        """Sum .grad over ranks.  With async_op=True the collectives are only queued (NCCL orders them after the
        kernels already on the current stream) and `wait()` must be called before the gradients are read on another
        stream; on the current stream later kernels are ordered behind them by the process group itself."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        import os
        from . import _gradacc
        params = self.params
        dev = params[0].device if params else torch.device("cpu")
        flats = _gradacc.last_pass_buffers(dev) if dev.type == "cuda" else []
        owned = {f.untyped_storage().data_ptr() for f in flats}
        if flats:
            rest = [p for p in params if p.grad is not None and p.grad.untyped_storage().data_ptr() not in owned]
        else:
            rest = params                      # no shared buffers on this device type: everything is packed (zeros for None)
        sig = (tuple(f.numel() for f in flats), tuple(p.numel() for p in rest))
        if sig != self._agreed or os.environ.get("SPLATCO_CHECK_ALLREDUCE") == "1":
            h = torch.tensor([len(flats), sum(sig[0]), len(rest), sum(sig[1])], dtype=torch.int64, device=dev)
            lo, hi = h.clone(), h.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if not torch.equal(lo, hi):
                raise RuntimeError("GradBucket.allreduce: ranks disagree on the gradient layout "
                                   f"(local {h.tolist()}, min {lo.tolist()}, max {hi.tolist()}): every rank must render "
                                   "at least one view per iteration and hold the same parameter set")
            self._agreed = sig
        small = None
        if rest:
            n = sum(sig[1])
            if self._small is None or self._small.numel() != n or self._small.device != dev:
                self._small = torch.empty(n, dtype=torch.float32, device=dev)
            small = self._small
            self._pack(rest, small)
        bufs = list(flats) + ([small] if small is not None else [])
        self.last_bytes = 4 * sum(b.numel() for b in bufs)
        works = []
        if len(bufs) > 1 and dev.type == "cuda" and dist.get_backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
            # one NCCL group launch for all buffers
            with dist._coalescing_manager(group=group, device=dev, async_ops=True) as cm:
                for b in bufs:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
            works.append(cm)
        else:
            for b in bufs:
                works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group, async_op=True))
        self._pending = (works, rest, small)
        if not async_op:
            self.wait()
        return self

    def wait(self):
        pend, self._pending = getattr(self, "_pending", None), None
        if pend is None:
            return
        works, rest, small = pend
        for w in works:
            if w is not None:
                w.wait()                       # NCCL: orders the current stream behind the collective, no host block
        if small is not None:
            self._unpack(rest, small)


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
