#!/usr/bin/env python
"""cProfile of the host side of bench steps (which Python functions the step's host time goes to).
    python tools/host_profile.py [--steps 20]          diagnostic only"""
import argparse
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
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

    def step():
        for p in params:
            p.grad = None
        total = None
        for v in range(cfg["mv"]):
            vm = prefilter_voxel(cams[v], pc, bench.PIPE, bg)
            pkg = render(cams[v], pc, bench.PIPE, bg, visible_mask=vm, retain_grad=True)
            loss = l1_ssim_loss(pkg["render"], gts[v], 0.2) + 0.01 * scaling_reg(pkg["scaling"])
            total = loss if total is None else total + loss
        total.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime")
    print(f"per step = totals / {a.steps}")
    st.print_stats(45)
    st.sort_stats("cumtime")
    st.print_stats(70)


if __name__ == "__main__":
    main()
