"""Fused Adam: drop-in for the reference's `torch.optim.Adam(l, lr=0.0, eps=1e-15)`
(scene/gaussian_model.py:519-572), stepped at train.py:310-312.

    FusedAdam(param_groups, lr=0.0, eps=1e-15, betas=(0.9, 0.999))

Same constructor arguments, same `param_groups` (the reference rewrites `group["lr"]` every iteration,
scene/gaussian_model.py:574-606, and swaps `group["params"][0]` when it grows / prunes anchors, :733-758,784-815) and the
same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`), so `cat_tensors_to_optimizer`, `_prune_anchor_optimizer`
and `optimizer.state_dict()` checkpoints keep working.  step() updates every parameter that has a gradient in ONE pass per
element (`splatco_adam_step`, csrc/optim.cu; up to 64 tensors per launch) instead of torch's ~10 multi-tensor passes.
fp32 CUDA parameters only; no CPU fallback; weight decay / amsgrad / maximize are not supported (the reference never uses them).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check
from .profiling import stage


class _AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("step", C.c_int64), ("lr", C.c_float), ("reserved", C.c_float)]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("splatco_b200 FusedAdam: weight_decay / amsgrad are not supported")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._table = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        by_cfg = {}
        keep = []
        touched = []
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr = float(group["lr"])
            items = None
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("splatco_b200 FusedAdam needs CUDA parameters (no CPU fallback)")
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or g.is_sparse:
                    raise RuntimeError("splatco_b200 FusedAdam: contiguous fp32 parameters and dense fp32 gradients only")
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                # `step` is kept as a Python int (torch.optim.Adam accepts and converts such checkpoints); a tensor
                # loaded from a torch checkpoint is converted here
                step = state["step"] = int(state["step"]) + 1
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous()):
                    m = state["exp_avg"] = m.contiguous()
                    v = state["exp_avg_sq"] = v.contiguous()
                if items is None:
                    items = by_cfg.setdefault((p.device, float(beta1), float(beta2), float(group["eps"])), [])
                items.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), step, lr))
                touched += (p, m, v)
        for (dev, beta1, beta2, eps), items in by_cfg.items():
            n = len(items)
            if self._table is None or len(self._table) < n:
                self._table = (_AdamTensor * max(n, 128))()
            tab = self._table
            for i, it in enumerate(items):
                e = tab[i]
                e.param, e.grad, e.exp_avg, e.exp_avg_sq, e.numel, e.step, e.lr = it
            with _lib.on_device(dev):
                with stage("adam_step"):
                    check(L.splatco_adam_step(n, C.byref(tab), beta1, beta2, eps, _lib.raw_stream(dev)), "splatco_adam_step")
        # the kernel writes through raw pointers: tell autograd (and every cache keyed on `_version` -- the decode's
        # TriPlaneAttention / channel-last plane copies, the renderer's exp(_scaling) stash) that these tensors changed,
        # as torch.optim.Adam's in-place ops would
        if touched:
            torch.autograd.graph.increment_version(touched)
        return loss
