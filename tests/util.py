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
