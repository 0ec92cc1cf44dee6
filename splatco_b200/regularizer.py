"""Plane total-variation regulariser, gradient added in place.

    total_variation_add_grad(plane_grid, w)   drop-in for PlaneGrid.total_variation_add_grad (scene/grids.py:240-250)
    tv_loss(learner, w)                       drop-in for GaussianLearner.tv_loss (scene/gaussian_model.py:217-220),
                                              called as `gaussians.feat_planes.tv_loss(opt.tv_weight_a)` at train.py:242-243

The reference builds six smooth-L1 sums per PlaneGrid and calls backward() on them; here one kernel per plane
(`splatco_tv_add_grad`, csrc/regularizer.cu) adds the same gradient to `plane.grad`.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def _add_plane(plane, w):
    if not plane.is_cuda:
        raise RuntimeError("splatco_b200 total_variation_add_grad needs CUDA planes (no CPU fallback)")
    if plane.dim() != 4 or plane.shape[0] != 1 or plane.dtype != torch.float32 or not plane.is_contiguous():
        raise RuntimeError(f"total_variation_add_grad: expected a contiguous fp32 [1,C,H,W] plane, got {tuple(plane.shape)} {plane.dtype}")
    if plane.grad is None:                       # autograd would create it; backward() on a fresh plane does the same
        plane.grad = torch.zeros_like(plane)
    g = plane.grad
    if g.dtype != torch.float32 or not g.is_contiguous():
        raise RuntimeError("total_variation_add_grad: plane.grad must be contiguous fp32")
    _, C, H, W = (int(s) for s in plane.shape)
    dev = plane.device
    with _lib.on_device(dev):
        check(_lib.lib().splatco_tv_add_grad(C, H, W, ptr(plane.detach()), ptr(g), float(w), _lib.raw_stream(dev)), "splatco_tv_add_grad")


def total_variation_add_grad(plane_grid, w):
    """scene/grids.py:240-250 — wx = wy = wz = w, every term divided by 6."""
    with torch.no_grad():
        for plane in (plane_grid.xy_plane, plane_grid.xz_plane, plane_grid.yz_plane):
            _add_plane(plane, w)


def tv_loss(learner, w):
    """scene/gaussian_model.py:217-220 — levels 0..activate_level, weight w * 0.5 ** (2 - level)."""
    feat = learner._feat
    for level in range(int(feat.activate_level) + 1):
        total_variation_add_grad(feat.k0s[level], w * (0.5 ** (2 - level)))
