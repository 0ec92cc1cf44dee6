"""Anchor growing / pruning (densification), SURVEY.md §8 row f4.

    anchor_growing(pc, grads, threshold, offset_mask)       drop-in for GaussianModel.anchor_growing
                                                            (scene/gaussian_model.py:832-925)
    adjust_anchor(pc, iteration, check_interval=100, ...)   drop-in for GaussianModel.adjust_anchor (:929-997)
    grow_pass(...)                                          one pass of anchor_growing's loop on plain tensors

Both take the model as first argument so they can be bound in place of the reference methods
(`GaussianModel.anchor_growing = splatco_b200.densify.anchor_growing`).  The voxel arithmetic — candidate positions,
rounding to the grid, de-duplication, the test against every existing anchor, the row order of the new anchors and
the per-voxel feature maximum — runs in libsplatco_b200.so (csrc/densify.cu: a hash set of candidate voxels instead
of unique(dim=0) + an O(U*N) compare); appending rows to the optimizer state stays the model's own method
(`pc.cat_tensors_to_optimizer`, `pc.prune_anchor`), which is bookkeeping outside the kernels.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def grow_pass(anchor, offset, scaling, anchor_feat, cur_size, *, candidate_mask=None, grads=None, threshold=0.0,
              offset_mask=None, rand=None, rand_cut=0.0, div_mode=0):
    """One pass: returns (new_anchor [U,3], new_feat [U,F], n_candidates).

    anchor [N,3], offset [N,K,3], scaling [N,>=3] (activated; only [:, :3] is read, any row stride), anchor_feat [N,F].
    Either `candidate_mask` (bool [n_stat]) or (grads [n_stat], threshold, offset_mask [n_stat], rand [n_stat], rand_cut),
    n_stat <= N*K rows (anchors appended by an earlier pass have no statistics and are never candidates).
    div_mode 0 rounds anchor * (1 / cur_size) like torch's CUDA division by a scalar, 1 rounds anchor / cur_size."""
    L = _lib.lib()
    if not anchor.is_cuda:
        raise RuntimeError("splatco_b200 anchor growing needs CUDA tensors (no CPU fallback)")
    dev = anchor.device
    N, K = int(offset.shape[0]), int(offset.shape[1])
    F = int(anchor_feat.shape[1])
    a = anchor.detach().float().contiguous()
    o = offset.detach().float().contiguous()
    f = anchor_feat.detach().float().contiguous()
    s = scaling.detach()
    if s.dtype != torch.float32 or s.stride(1) != 1:
        s = s.float().contiguous()
    u8 = lambda t: None if t is None else t.detach().reshape(-1).to(torch.uint8).contiguous()
    f32 = lambda t: None if t is None else t.detach().reshape(-1).float().contiguous()
    cm, om, g, r = u8(candidate_mask), u8(offset_mask), f32(grads), f32(rand)
    n_stat = int((cm if cm is not None else g).shape[0])
    sel = (ptr(cm), ptr(g), float(threshold), ptr(om), ptr(r), float(rand_cut))
    with _lib.on_device(dev):
        st = _lib.raw_stream(dev)
        counts = torch.empty(4, dtype=torch.int32, device=dev)
        check(L.splatco_grow_count(n_stat, *sel, ptr(counts), st), "splatco_grow_count")
        n_cand = int(counts[0].item())
        new_anchor = torch.empty(0, 3, dtype=torch.float32, device=dev)
        new_feat = torch.empty(0, F, dtype=torch.float32, device=dev)
        if n_cand == 0:
            return new_anchor, new_feat, 0
        ws = torch.empty(L.splatco_grow_ws_bytes(n_cand), dtype=torch.uint8, device=dev)
        check(L.splatco_grow_unique(N, K, n_stat, ptr(a), ptr(o), ptr(s), int(s.stride(0)), *sel, float(cur_size), int(div_mode),
                                    n_cand, ptr(ws), ptr(counts), st), "splatco_grow_unique")
        c = counts.tolist()
        if c[1] != n_cand:
            raise RuntimeError(f"grow_unique: candidate count changed between passes ({n_cand} -> {c[1]})")
        n_new = c[2]
        if n_new == 0:
            return new_anchor, new_feat, n_cand
        new_anchor = torch.empty(n_new, 3, dtype=torch.float32, device=dev)
        new_feat = torch.empty(n_new, F, dtype=torch.float32, device=dev)
        check(L.splatco_grow_emit(K, F, float(cur_size), n_cand, n_new, ptr(ws), ptr(f), ptr(new_anchor), ptr(new_feat), st),
              "splatco_grow_emit")
    return new_anchor, new_feat, n_cand


def anchor_growing(pc, grads, threshold, offset_mask):
    """scene/gaussian_model.py:832-925.  The random thinning mask is drawn with torch.rand_like on the same device and
    in the same order as the reference draws it, so a seeded run picks the same candidates."""
    K = int(pc.n_offsets)
    n_stat = int(pc.get_anchor.shape[0]) * K
    grads = grads.reshape(-1)
    dev = pc.get_anchor.device
    for i in range(int(pc.update_depth)):
        cur_threshold = threshold * ((pc.update_hierachy_factor // 2) ** i)
        rand = torch.rand_like(grads, dtype=torch.float32)                       # :845 rand_like(candidate_mask.float())
        grown = int(pc.get_anchor.shape[0]) * K - n_stat
        if grown == 0 and i > 0:                                                  # :849-851
            continue
        size_factor = pc.update_init_factor // (pc.update_hierachy_factor ** i)  # :859
        cur_size = pc.voxel_size * size_factor
        new_anchor, new_feat, _ = grow_pass(pc.get_anchor, pc._offset, pc.get_scaling, pc._anchor_feat, cur_size,
                                            grads=grads, threshold=cur_threshold, offset_mask=offset_mask, rand=rand,
                                            rand_cut=0.5 ** (i + 1))
        U = int(new_anchor.shape[0])
        if U == 0:
            continue
        new_scaling = torch.full((U, 6), float(cur_size), dtype=torch.float32, device=dev).log()     # :888-889
        new_rotation = torch.zeros(U, 4, dtype=torch.float32, device=dev)
        new_rotation[:, 0] = 1.0
        tenth = torch.full((U, 1), 0.1, dtype=torch.float32, device=dev)
        new_opacities = torch.log(tenth / (1 - tenth))                                               # :893 inverse_sigmoid(0.1)
        new_offsets = torch.zeros(U, K, 3, dtype=torch.float32, device=dev)
        pc.anchor_demon = torch.cat([pc.anchor_demon, torch.zeros(U, 1, dtype=torch.float32, device=dev)], dim=0)
        pc.opacity_accum = torch.cat([pc.opacity_accum, torch.zeros(U, 1, dtype=torch.float32, device=dev)], dim=0)
        grown_params = pc.cat_tensors_to_optimizer({"anchor": new_anchor, "scaling": new_scaling, "rotation": new_rotation,
                                                    "anchor_feat": new_feat, "offset": new_offsets, "opacity": new_opacities})
        pc._anchor, pc._scaling, pc._rotation = grown_params["anchor"], grown_params["scaling"], grown_params["rotation"]
        pc._anchor_feat, pc._offset, pc._opacity = grown_params["anchor_feat"], grown_params["offset"], grown_params["opacity"]


def adjust_anchor(pc, iteration, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005):
    """scene/gaussian_model.py:929-997: grow from the accumulated offset gradients, reset the statistics of the offsets
    that were eligible, prune anchors whose accumulated opacity stayed low, resize the statistics."""
    K = int(pc.n_offsets)
    dev = pc.get_anchor.device
    with torch.no_grad():
        grads = pc.offset_gradient_accum / pc.offset_denom
        grads[grads.isnan()] = 0.0
        grads_norm = torch.norm(grads, dim=-1)
        offset_mask = (pc.offset_denom > check_interval * success_threshold * 0.5).squeeze(dim=1)
        if iteration % 3000 == 0 or iteration == 1600:                            # :936-945 curvature densification
            curv = pc.compute_curvature(pc.get_anchor).view(pc.get_anchor.shape[0], -1)
            offset_mask = torch.logical_or(offset_mask, torch.cat([(curv <= 0.1).squeeze()] * K, dim=0))

        anchor_growing(pc, grads_norm, grad_threshold, offset_mask)

        n_slots = int(pc.get_anchor.shape[0]) * K
        for name in ("offset_denom", "offset_gradient_accum"):                    # :950-961 (the padding is int32 there,
            t = getattr(pc, name)                                                 #  torch.cat promotes it back to fp32)
            t[offset_mask] = 0
            setattr(pc, name, torch.cat([t, torch.zeros(n_slots - t.shape[0], 1, dtype=t.dtype, device=dev)], dim=0))

        prune_mask = (pc.opacity_accum < min_opacity * pc.anchor_demon).squeeze(dim=1)
        anchors_mask = (pc.anchor_demon > check_interval * success_threshold).squeeze(dim=1)
        prune_mask = torch.logical_and(prune_mask, anchors_mask)
        keep = ~prune_mask
        pc.offset_denom = pc.offset_denom.view(-1, K)[keep].view(-1, 1)
        pc.offset_gradient_accum = pc.offset_gradient_accum.view(-1, K)[keep].view(-1, 1)
        pc.opacity_accum[anchors_mask] = 0.0
        pc.anchor_demon[anchors_mask] = 0.0
        pc.opacity_accum = pc.opacity_accum[keep]
        pc.anchor_demon = pc.anchor_demon[keep]
        if prune_mask.shape[0] > 0:
            pc.prune_anchor(prune_mask)
        pc.max_radii2D = torch.zeros(pc.get_anchor.shape[0], device=dev)
