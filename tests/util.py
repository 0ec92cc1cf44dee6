"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from splatco_b200.synthetic import random_gaussians, ring_cameras


def scene(M, W, H, seed, sigma_px=(0.5, 4.0), cam_index=1, n_cams=3):
    cam = ring_cameras(n_cams, W, H)[cam_index]
    means, colors, opac, scales, rots = random_gaussians(M, cam, seed, sigma_px=sigma_px)
    return cam, means, colors, opac, scales, rots


def tans(cam):
    return math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)


def oracle_forward(cam, means, colors, opac, scales, rots, bg, scale_mod=1.0):
    from oracle import raster as R
    tx, ty = tans(cam)
    return R.rasterize_forward(means.numpy(), colors.numpy(), opac.numpy(), scales.numpy(), rots.numpy(),
                               scale_mod, cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(),
                               tx, ty, cam.image_height, cam.image_width, np.asarray(bg, np.float32))


def oracle_backward(fw, cam, means, colors, scales, rots, bg, dL, scale_mod=1.0):
    from oracle import raster as R
    tx, ty = tans(cam)
    return R.rasterize_backward(fw, means.numpy(), colors.numpy(), scales.numpy(), rots.numpy(), scale_mod,
                                cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), tx, ty,
                                cam.image_height, cam.image_width, np.asarray(bg, np.float32), dL)


def settings_for(cam, bg, device="cuda", scale_mod=1.0, debug=False):
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizationSettings
    tx, ty = tans(cam)
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=tx, tanfovy=ty,
        bg=torch.as_tensor(bg, dtype=torch.float32, device=device), scale_modifier=scale_mod,
        viewmatrix=cam.world_view_transform.to(device), projmatrix=cam.full_proj_transform.to(device),
        sh_degree=1, campos=cam.camera_center.to(device), prefiltered=False, debug=debug)


def chunk(buf: torch.Tensor, offset: int, dtype, count: int) -> torch.Tensor:
    """Typed view of a chunk inside a uint8 workspace tensor."""
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[offset: offset + nbytes].view(dtype)


def layout(kind, *args):
    from splatco_b200 import _lib
    L = _lib.lib()
    offs = (C.c_size_t * 8)()
    n = getattr(L, f"splatco_{kind}_layout")(*args, offs, 8)
    return [int(offs[i]) for i in range(n)]


def rel_err(a, b, floor_frac=1e-3):
    """|a-b| / max(|b|, floor) with floor = floor_frac * max|b| (atomics reorder fp32 sums)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = max(float(np.abs(b).max()) * floor_frac, 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


def full_path_grad_errors(a, b):
    """Agreement of two gradients that went through the WHOLE path (decode + rasterizer) on both sides.

    The rasterizer is discontinuous in its inputs (alpha = 1/255 cut, T = 1e-4 stop, integer radii / tile rectangles) and
    the two sides' decode outputs differ by fp32 rounding (~1e-6), so a handful of (pixel, splat) pairs are kept by one
    side and dropped by the other.  Each such pair moves a few gradient entries by up to ~1e-3 of the tensor's largest
    entry -- measured with two torch evaluations of the same decode (fp32 vs fp64) in front of the SAME rasterizer
    (tools/debug_grad_diff.py), i.e. it is a property of the function, not of an implementation.  Element-wise relative
    error with a small floor is therefore meaningless here; what must hold is agreement in norm and a bounded worst entry:
        l2   = |a - b|_2 / |b|_2                 (north star: 1e-3 relative)
        amax = max|a - b| / max|b|
    Stage-level tests on IDENTICAL inputs (test_raster_gpu, test_decode_gpu) keep the element-wise 1e-3 bar."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    nb = float(np.linalg.norm(b))
    mb = float(np.abs(b).max()) if b.size else 0.0
    if nb == 0.0:
        return {"l2": float(np.linalg.norm(a)), "amax": float(np.abs(a).max()) if a.size else 0.0}
    return {"l2": float(np.linalg.norm(a - b)) / nb, "amax": float(np.abs(a - b).max()) / mb}


def worst_entry(a, b, floor_frac=1e-3):
    """(rel err, flat index, got, want, max|want|) of the entry with the largest rel_err: for assertion messages."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    floor = max(float(np.abs(b).max()) * floor_frac, 1e-30)
    e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    i = int(e.argmax())
    return float(e[i]), i, float(a[i]), float(b[i]), float(np.abs(b).max())
