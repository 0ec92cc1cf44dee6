"""CVPM pruning mask: drop-in for GaussianModel.compute_fast_loss_with_key_points
(scene/gaussian_model.py:1112-1219), called per view pair inside the mv loop (train.py:218-234) with
`existing_point_cloud = gaussians.get_anchor`, `distance_threshold = gaussians.voxel_size`; the returned mask goes to
`prune_anchor`.

    compute_fast_loss_with_key_points(pc, real_img1, real_img2, gen_img1, gen_img2, K1, R1, t1, K2, R2, t2,
                                      existing_point_cloud, distance_threshold=0.01, overall_ssim_threshold=0.6,
                                      sigma_threshold=3.0, min_cam_distance=0.5, device='cuda', pts_flag=True)
        -> (weighted_gen_l1_loss, weighted_cross_l1_loss, intersecting_points, combined_mask)

Same arguments (the model first, so it can be bound in place of the method) and the same 4-tuple.  The SSIM of the two
ground-truth images comes from the fused loss kernels, the mask from `splatco_cvpm_mask` (csrc/cvpm.cu); the SSIM gate is
applied on the device, so — unlike the reference's `if overall_ssim < threshold` — nothing here waits for the GPU unless
`pts_flag` asks for the gathered points.  Differences a caller can see: when the gate is closed the two losses are
zero-valued device scalars (the reference returns the Python float 0.0), and the losses are detached (train.py calls
this under no_grad and discards them).  K1, R1, K2, R2 are accepted and unused, as in the reference.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .loss import l1_ssim_loss


def _crop4(a, b, c, d):
    H = min(int(t.shape[1]) for t in (a, b, c, d))
    W = min(int(t.shape[2]) for t in (a, b, c, d))
    return [t[:, :H, :W] for t in (a, b, c, d)]


def cvpm_mask(points, t1, t2, ssim=None, distance_threshold=0.01, overall_ssim_threshold=0.6, sigma_threshold=3.0,
              min_cam_distance=0.5):
    """-> (bool mask [N], int32 count [1]) on the device; `ssim` a device scalar (or None: gate open)."""
    L = _lib.lib()
    if not points.is_cuda:
        raise RuntimeError("splatco_b200 cvpm_mask needs CUDA tensors (no CPU fallback)")
    dev = points.device
    f = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    p, a, b, s = points.detach().float().contiguous(), f(t1), f(t2), f(ssim)
    N = int(p.shape[0])
    with _lib.on_device(dev):
        mask = torch.empty(N, dtype=torch.bool, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        ws = torch.empty(L.splatco_cvpm_ws_bytes(), dtype=torch.uint8, device=dev)
        check(L.splatco_cvpm_mask(N, ptr(p), ptr(a), ptr(b), ptr(s), float(overall_ssim_threshold), float(distance_threshold),
                                  float(sigma_threshold), float(min_cam_distance), ptr(ws), ptr(mask), ptr(count),
                                  _lib.raw_stream(dev)), "splatco_cvpm_mask")
    return mask, count


def compute_fast_loss_with_key_points(pc, real_img1, real_img2, gen_img1, gen_img2, K1, R1, t1, K2, R2, t2,
                                      existing_point_cloud, distance_threshold=0.01, overall_ssim_threshold=0.6,
                                      sigma_threshold=3.0, min_cam_distance=0.5, device="cuda", pts_flag=True):
    dev = existing_point_cloud.device
    with torch.no_grad():
        r1, r2, g1, g2 = [t.detach().to(dev) for t in _crop4(real_img1, real_img2, gen_img1, gen_img2)]
        s = l1_ssim_loss(r1, r2, 1.0, return_parts=True)[1][2]                     # overall SSIM, :1161
        gate = (s >= overall_ssim_threshold).to(torch.float32)
        # (|a - b| * s).mean() == s * l1(a, b), :1170-1176
        gen_l1 = gate * s * l1_ssim_loss(g1, g2, 0.0)
        cross_l1 = gate * s * l1_ssim_loss(r1, r2, 0.0)
        N = int(existing_point_cloud.shape[0])
        if not pts_flag:
            return gen_l1, cross_l1, torch.empty((0, 3), device=dev), torch.zeros(N, dtype=torch.bool, device=dev)
        mask, _ = cvpm_mask(existing_point_cloud, t1, t2, s, distance_threshold, overall_ssim_threshold, sigma_threshold,
                            min_cam_distance)
        return gen_l1, cross_l1, existing_point_cloud[mask], mask
