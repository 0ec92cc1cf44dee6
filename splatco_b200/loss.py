"""Fused image loss of the training loop.

    l1_ssim_loss(image, gt, lambda_dssim=0.2) -> scalar
        == (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
           with the reference's l1_loss / ssim (utils/loss_utils.py:17-18, 33-63; combined at train.py:192-196)
    l1_loss(a, b), ssim(a, b): the reference's names, evaluated by the same kernels.

    scaling_reg(scaling)                                 == scaling.prod(dim=1).mean()   (train.py:195; no host sync in backward)
    multiview_consistency_loss(gen_imgs, real_imgs, 0.6) == `muiti_con_loss` of train.py:199-216 (all view pairs, one pass)

Two kernels forward, one backward in libsplatco_b200.so (csrc/loss.cu) instead of 5 cuDNN grouped convolutions and
~15 elementwise launches each way; the backward writes dL/dimage, which is what the blend backward consumes.
Gradients flow to `image` only (the ground truth never requires grad in train.py); no CPU fallback.
"""
from __future__ import annotations

import ctypes as _C

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


class _ScalingReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scaling):
        L = _lib.lib()
        if not scaling.is_cuda or scaling.dim() != 2 or scaling.shape[1] != 3:
            raise RuntimeError(f"splatco_b200 scaling_reg needs a CUDA [M,3] tensor (no CPU fallback), got {tuple(scaling.shape)} on {scaling.device}")
        s = scaling.detach()
        if s.dtype != torch.float32 or not s.is_contiguous():
            s = s.float().contiguous()
        M = int(s.shape[0])
        dev = s.device
        with _lib.on_device(dev):
            ws = torch.empty(16, dtype=torch.uint8, device=dev)
            out = torch.empty(1, dtype=torch.float32, device=dev)
            with stage("scaling_reg_fwd"):
                check(L.splatco_scaling_reg_fwd(M, ptr(s), ptr(ws), ptr(out), _lib.raw_stream(dev)), "splatco_scaling_reg_fwd")
        ctx.keep = (s, M)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        s, M = ctx.keep
        dev = s.device
        gl = g if (g.dtype == torch.float32 and g.is_contiguous()) else g.detach().float().contiguous()
        with _lib.on_device(dev):
            ds = torch.empty_like(s)
            with stage("scaling_reg_bwd"):
                check(L.splatco_scaling_reg_bwd(M, ptr(s), ptr(gl), ptr(ds), _lib.raw_stream(dev)), "splatco_scaling_reg_bwd")
        return ds


def scaling_reg(scaling):
    """`scaling.prod(dim=1).mean()` of train.py:195 (the per-view loss adds 0.01 * this) for the [M,3] `render()["scaling"]`.
    torch's prod backward reads a zero count back to the host at the start of every view's backward; this one does not."""
    return _ScalingReg.apply(scaling)


# ---- cross-view consistency term of the mv batch (train.py:199-216, summed in at :237-239) ----------------------

def _pair_list(n):
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def pair_ssim(real_imgs):
    """ssim(real_i, real_j) of every pair i < j on the pair's common crop (align_images, train.py:79-96), as a device
    tensor [n(n-1)/2] (no host sync).  It depends on the ground-truth images only, so a caller that keeps its cameras
    can cache it per camera pair."""
    vals = []
    with torch.no_grad():
        for i, j in _pair_list(len(real_imgs)):
            H = min(int(real_imgs[i].shape[1]), int(real_imgs[j].shape[1]))
            W = min(int(real_imgs[i].shape[2]), int(real_imgs[j].shape[2]))
            vals.append(l1_ssim_loss(real_imgs[i][:, :H, :W], real_imgs[j][:, :H, :W], 1.0, return_parts=True)[1][2:3])
    return torch.cat(vals)


class _MVConsistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, pssim, n, *imgs):
        L = _lib.lib()
        gen, real = imgs[:n], imgs[n:]
        dev = gen[0].device
        for t in imgs:
            if not t.is_cuda or t.dim() != 3 or t.shape[0] != gen[0].shape[0]:
                raise RuntimeError("splatco_b200 multiview_consistency_loss needs [C,H,W] CUDA images (no CPU fallback)")
        gen_c = [t.detach().float().contiguous() for t in gen]
        real_c = [t.detach().float().contiguous() for t in real]
        C = int(gen_c[0].shape[0])
        for g_, r_ in zip(gen_c, real_c):
            if g_.shape != r_.shape:
                raise RuntimeError("multiview_consistency_loss: a view's rendered and ground-truth images differ in size")
        PA = _C.c_void_p * n
        IA = _C.c_int * n
        views = dict(gen=PA(*[t.data_ptr() for t in gen_c]), real=PA(*[t.data_ptr() for t in real_c]),
                     h=IA(*[int(t.shape[1]) for t in gen_c]), w=IA(*[int(t.shape[2]) for t in gen_c]))
        with _lib.on_device(dev):
            ws = torch.empty(L.splatco_mv_consistency_ws_bytes(n), dtype=torch.uint8, device=dev)
            out = torch.empty(1 + n * (n - 1) // 2, dtype=torch.float32, device=dev)
            ps = pssim.detach().float().contiguous()
            with stage("mv_consistency_fwd"):
                check(L.splatco_mv_consistency_fwd(n, C, views["gen"], views["real"], views["h"], views["w"], ptr(ps),
                                                   float(gate), ptr(ws), ptr(out), _lib.raw_stream(dev)), "splatco_mv_consistency_fwd")
        ctx.keep = (gen_c, real_c, views, ws, n, C)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        L = _lib.lib()
        gen_c, real_c, views, ws, n, C = ctx.keep
        dev = gen_c[0].device
        gl = g_loss if (g_loss.dtype == torch.float32 and g_loss.is_contiguous()) else g_loss.detach().float().contiguous()
        with _lib.on_device(dev):
            dgen = [torch.empty_like(t) for t in gen_c]
            DA = (_C.c_void_p * n)(*[t.data_ptr() for t in dgen])
            with stage("mv_consistency_bwd"):
                check(L.splatco_mv_consistency_bwd(n, C, views["gen"], views["real"], DA, views["h"], views["w"], ptr(ws),
                                                   ptr(gl), _lib.raw_stream(dev)), "splatco_mv_consistency_bwd")
        return (None, None, None) + tuple(dgen) + (None,) * n


def multiview_consistency_loss(gen_imgs, real_imgs, ssim_threshold=0.6, pair_ssim_values=None, return_parts=False):
    """`muiti_con_loss` of train.py:199-236 (without the CVPM pruning call inside that loop): the sum over view pairs
    i < j of  ssim(real_i, real_j) * |l1_loss(real_i - real_j, gen_i - gen_j)|  for pairs whose ground-truth SSIM
    exceeds `ssim_threshold`, each pair on its common crop.  train.py adds 0.05 * this to the summed loss.
    Gradients flow to the generated images only.  2 <= len(gen_imgs) <= 8."""
    n = len(gen_imgs)
    if n != len(real_imgs):
        raise RuntimeError("multiview_consistency_loss: need as many ground-truth as generated images")
    if n < 2:
        z = gen_imgs[0].new_zeros(())
        return (z, z.reshape(1)) if return_parts else z
    if n > 8:
        raise NotImplementedError("splatco_b200 multiview_consistency_loss: at most 8 views per batch")
    ps = pair_ssim(real_imgs) if pair_ssim_values is None else pair_ssim_values
    loss, parts = _MVConsistency.apply(float(ssim_threshold), ps, n, *gen_imgs, *real_imgs)
    return (loss, parts) if return_parts else loss
