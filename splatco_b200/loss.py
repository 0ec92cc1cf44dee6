"""Fused image loss of the training loop.

    l1_ssim_loss(image, gt, lambda_dssim=0.2) -> scalar
        == (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
           with the reference's l1_loss / ssim (utils/loss_utils.py:17-18, 33-63; combined at train.py:192-196)
    l1_loss(a, b), ssim(a, b): the reference's names, evaluated by the same kernels.

Two kernels forward, one backward in libsplatco_b200.so (csrc/loss.cu) instead of 5 cuDNN grouped convolutions and
~15 elementwise launches each way; the backward writes dL/dimage, which is what the blend backward consumes.
Gradients flow to `image` only (the ground truth never requires grad in train.py); no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .profiling import stage


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        L = _lib.lib()
        if not image.is_cuda or not gt.is_cuda:
            raise RuntimeError("splatco_b200 l1_ssim_loss needs CUDA tensors (no CPU fallback)")
        if image.shape != gt.shape or image.dim() != 3:
            raise RuntimeError(f"l1_ssim_loss: expected two [C,H,W] images, got {tuple(image.shape)} and {tuple(gt.shape)}")
        dev = image.device
        img = image.detach().float().contiguous()
        g = gt.detach().float().contiguous()
        C, H, W = (int(s) for s in img.shape)
        with _lib.on_device(dev):
            ws = torch.empty(L.splatco_loss_ws_bytes(C, H, W), dtype=torch.uint8, device=dev)
            out3 = torch.empty(3, dtype=torch.float32, device=dev)
            with stage("l1_ssim_fwd"):
                check(L.splatco_l1_ssim_fwd(C, H, W, ptr(img), ptr(g), float(lambda_dssim), ptr(ws), ptr(out3),
                                            _lib.raw_stream(dev)), "splatco_l1_ssim_fwd")
        ctx.keep = (img, g, ws, float(lambda_dssim), (C, H, W))
        ctx.mark_non_differentiable(out3)
        return out3[0], out3

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        L = _lib.lib()
        img, g, ws, lam, (C, H, W) = ctx.keep
        dev = img.device
        gl = g_loss if (g_loss.dtype == torch.float32 and g_loss.is_contiguous()) else g_loss.detach().float().contiguous()
        with _lib.on_device(dev):
            dimg = torch.empty_like(img)
            with stage("l1_ssim_bwd"):
                check(L.splatco_l1_ssim_bwd(C, H, W, ptr(img), ptr(g), lam, ptr(ws), ptr(gl), ptr(dimg),
                                            _lib.raw_stream(dev)), "splatco_l1_ssim_bwd")
        return dimg, None, None


def l1_ssim_loss(image, gt, lambda_dssim=0.2, return_parts=False):
    """(1 - lambda) * L1 + lambda * (1 - SSIM).  With return_parts=True also returns the detached device tensor
    (loss, l1, ssim) the reference logs separately (train.py:193-194, Ll1 / ssim_loss)."""
    if gt.requires_grad:
        raise NotImplementedError("splatco_b200 l1_ssim_loss differentiates w.r.t. the rendered image only")
    loss, parts = _L1SSIM.apply(image, gt, lambda_dssim)
    return (loss, parts) if return_parts else loss


def l1_loss(network_output, gt):
    """utils/loss_utils.py:17-18."""
    return l1_ssim_loss(network_output, gt, 0.0)


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:33-42 (window 11, size_average=True: the only form train.py uses)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("splatco_b200 ssim: window_size=11, size_average=True only")
    return 1.0 - l1_ssim_loss(img1, img2, 1.0)
