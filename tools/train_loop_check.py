#!/usr/bin/env python
"""Short training loop over the drop-in path (C2 scene, mv=4): the loss must decrease, memory must stay flat, and the
per-iteration wall time is printed for the reference's optimizer (torch.optim.Adam) and for splatco_b200.optim.FusedAdam.
    python tools/train_loop_check.py [--iters 60]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from splatco_b200.gaussian_renderer import prefilter_voxel, render
from splatco_b200.loss import l1_ssim_loss, scaling_reg
from splatco_b200.optim import FusedAdam


def run(kind, iters):
    dev = torch.device("cuda", 0)
    cfg = bench.WORKLOADS["c2"]
    pc = bench.build_model(cfg, dev)
    pc.feat_planes.Q0 = 0.03
    cams, gts = bench.build_views(cfg)
    cams = [c.to(dev) for c in cams]
    gts = [g.to(dev) for g in gts]
    bg = torch.ones(3, device=dev)
    groups = [{"params": [p], "lr": 1e-4, "name": f"p{i}"} for i, p in enumerate(pc.parameters()) if p.requires_grad]
    opt = FusedAdam(groups, lr=0.0, eps=1e-15) if kind == "fused" else torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    times, losses = [], {}
    for it in range(iters + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        total = None
        for v in range(cfg["mv"]):
            vm = prefilter_voxel(cams[v], pc, bench.PIPE, bg)
            pkg = render(cams[v], pc, bench.PIPE, bg, visible_mask=vm, retain_grad=True)
            loss = l1_ssim_loss(pkg["render"], gts[v], 0.2) + 0.01 * scaling_reg(pkg["scaling"])
            total = loss if total is None else total + loss
        total.backward()
        t1 = time.perf_counter()
        opt.step()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1, t2 - t0))
        if it in (5, iters // 3, iters):
            losses[it] = total.item()
            print(f"[{kind}] it {it} loss {losses[it]:.5f} alloc {torch.cuda.memory_allocated() / 1e6:.0f} MB reserved {torch.cuda.memory_reserved() / 1e6:.0f} MB", flush=True)
    tail = times[10:]
    med = lambda k: sorted(t[k] for t in tail)[len(tail) // 2] * 1e3
    print(f"[{kind}] median per iteration: host fwd+bwd {med(0):.2f} ms, optimizer step + drain {med(1):.2f} ms, total {med(2):.2f} ms; "
          f"max total {max(t[2] for t in tail) * 1e3:.2f} ms", flush=True)
    ks = sorted(losses)
    assert losses[ks[-1]] < losses[ks[0]], "loss did not decrease"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    for kind in ("fused", "torch"):
        run(kind, a.iters)
