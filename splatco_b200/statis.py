"""Drop-in for GaussianModel.training_statis (reference scene/gaussian_model.py:761-782).

    training_statis(pc, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask)

Same arguments as the reference method with the model as first argument (so it can be bound in its place:
`GaussianModel.training_statis = splatco_b200.statis.training_statis`), same in-place updates of
`pc.opacity_accum [N,1]`, `pc.anchor_demon [N,1]`, `pc.offset_gradient_accum [N*K,1]`, `pc.offset_denom [N*K,1]`.
One kernel in libsplatco_b200.so instead of ~15 masked torch ops over [N*K] tensors; no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def training_statis(pc, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
    L = _lib.lib()
    acc = [pc.opacity_accum, pc.anchor_demon, pc.offset_gradient_accum, pc.offset_denom]
    dev = acc[0].device
    if not acc[0].is_cuda:
        raise RuntimeError("splatco_b200 training_statis needs CUDA tensors (no CPU fallback)")
    for t in acc:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("splatco_b200 training_statis: accumulators must be contiguous fp32 (as the reference creates them)")
    K = int(pc.n_offsets)
    with torch.no_grad():
        vis_idx = torch.nonzero(anchor_visible_mask).squeeze(1).to(torch.int32)
        V = int(vis_idx.shape[0])
        sel = offset_selection_mask.reshape(-1)
        if int(sel.shape[0]) != V * K or int(opacity.numel()) != V * K:
            raise RuntimeError(f"training_statis: {V} visible anchors x {K} offsets, but mask/opacity have "
                               f"{int(sel.shape[0])}/{int(opacity.numel())} rows")
        sel_u8 = sel.to(torch.uint8).contiguous()
        sel_excl = (torch.cumsum(sel_u8, dim=0, dtype=torch.int32) - sel_u8).contiguous()
        upd = update_filter.reshape(-1).to(torch.uint8).contiguous()
        grad = viewspace_point_tensor.grad
        if grad is None and int(upd.shape[0]) > 0:
            raise RuntimeError("training_statis: viewspace_point_tensor.grad is None (call after backward, as train.py does)")
        grad = None if grad is None else grad.detach().float().contiguous()
        nopac = opacity.detach().reshape(-1).float().contiguous()
        with _lib.on_device(dev):
            check(L.splatco_training_statis(V, K, ptr(vis_idx), ptr(nopac), ptr(sel_u8), ptr(sel_excl), ptr(upd), ptr(grad),
                                            *[ptr(t) for t in acc], _lib.raw_stream(dev)), "splatco_training_statis")
