#!/usr/bin/env python
"""Wall-clock time the HOST spends inside the main functions of a bench step (perf_counter wrappers, no profiler):
which part of the forward's ~1 ms of Python per view goes where.  Times are inclusive (callee time is inside the
caller's); `Event.synchronize` is the wait for the GPU.    python tools/host_timeline.py [--steps 30]      diagnostic only"""
import argparse
import os
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

ACC = defaultdict(lambda: [0, 0.0])


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            e = ACC[label]
            e[0] += 1
            e[1] += time.perf_counter() - t
    setattr(obj, name, w)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    import splatco_b200.decode as dec
    import splatco_b200.diff_gaussian_rasterization as dgr
    import splatco_b200.gaussian_renderer as gr
    import splatco_b200.loss as loss
    from splatco_b200 import _gradacc
    wrap(dec, "collect_model"); wrap(dec, "_fill_desc")
    wrap(dec._FusedDecode, "forward", "_FusedDecode.forward"); wrap(dec._FusedDecode, "backward", "_FusedDecode.backward")
    wrap(dgr, "preprocess_speculative"); wrap(dgr, "_queue_binning_blend"); wrap(dgr, "publish_speculated")
    wrap(dgr, "visible_mask_compact"); wrap(dgr, "rasterize_backward_state")
    wrap(dgr._RasterizeGaussians, "forward", "_RasterizeGaussians.forward")
    wrap(gr, "generate_neural_gaussians"); wrap(gr, "_settings"); wrap(gr, "_get_scaling")
    wrap(_gradacc, "acquire", "_gradacc.acquire")
    wrap(torch.cuda.Event, "synchronize", "Event.synchronize")
    wrap(loss._L1SSIM, "forward", "_L1SSIM.forward") if hasattr(loss, "_L1SSIM") else None
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.loss import l1_ssim_loss, scaling_reg
    dev = torch.device("cuda", 0)
    cfg = bench.WORKLOADS[a.workload]
    pc = bench.build_model(cfg, dev)
    pc.feat_planes.Q0 = 0.03
    cams, gts = bench.build_views(cfg)
    cams = [c.to(dev) for c in cams]
    gts = [g.to(dev) for g in gts]
    bg = torch.ones(3, device=dev)
    params = pc.parameters()
    T = defaultdict(float)

    def step():
        for p in params:
            p.grad = None
        total = None
        for v in range(cfg["mv"]):
            t0 = time.perf_counter()
            vm = prefilter_voxel(cams[v], pc, bench.PIPE, bg)
            t1 = time.perf_counter()
            pkg = render(cams[v], pc, bench.PIPE, bg, visible_mask=vm, retain_grad=True)
            t2 = time.perf_counter()
            l1 = l1_ssim_loss(pkg["render"], gts[v], 0.2)
            t3 = time.perf_counter()
            sr = scaling_reg(pkg["scaling"])
            t4 = time.perf_counter()
            loss_v = l1 + 0.01 * sr
            total = loss_v if total is None else total + loss_v
            t5 = time.perf_counter()
            T["prefilter_voxel"] += t1 - t0; T["render"] += t2 - t1; T["l1_ssim_loss"] += t3 - t2
            T["scaling_reg"] += t4 - t3; T["torch loss arithmetic"] += t5 - t4
        t6 = time.perf_counter()
        total.backward()
        T["backward (host)"] += time.perf_counter() - t6

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ACC.clear(); T.clear()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    nv = a.steps * cfg["mv"]
    print(f"{1e3 * t_all / a.steps:.2f} ms/step wall, host returns after {1e3 * t_host / a.steps:.2f} ms/step; per VIEW, microseconds:")
    for k, v in T.items():
        print(f"  {k:28s} {1e6 * v / nv:8.1f}")
    print("  inside (inclusive, per view):")
    for k, (n, v) in sorted(ACC.items(), key=lambda kv: -kv[1][1]):
        print(f"    {k:30s} {1e6 * v / nv:8.1f}   ({n / nv:.1f} calls/view)")


if __name__ == "__main__":
    main()
