#!/usr/bin/env python
"""Where one bench step spends its time (torch.profiler): GPU kernel totals split into this
library's kernels vs torch's, GPU idle time, and the host-side cost of the top CPU ops.
    python tools/step_profile.py [--workload c2] [--steps 5]
Diagnostic only: numbers printed here are taken under a profiler and are never bench values."""
import argparse
import os
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=5)
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
        t_f = time.perf_counter()
        total.backward()
        return t_f

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # un-profiled wall clock, forward / backward split on the host
    t0 = time.perf_counter(); fw = 0.0
    for _ in range(a.steps):
        s0 = time.perf_counter(); tf = step(); fw += tf - s0
    host_done = time.perf_counter()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"un-profiled: {1e3 * (t1 - t0) / a.steps:.2f} ms/step wall; host returns after {1e3 * (host_done - t0) / a.steps:.2f} ms/step "
          f"(forward part {1e3 * fw / a.steps:.2f} ms/step)")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    kern = defaultdict(lambda: [0, 0.0])
    cpu = defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = kern[e.name[:90]]; k[0] += 1; k[1] += e.device_time
        else:
            c = cpu[e.name[:60]]; c[0] += 1; c[1] += e.self_cpu_time_total
    n = a.steps
    # GPU idle gaps: sort device events by start time, report what ran before/after the large gaps
    dev_ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    gaps = defaultdict(lambda: [0, 0.0])
    idle = 0.0
    end = None
    prev = None
    for e in dev_ev:
        st, en = e.time_range.start, e.time_range.end
        if end is not None and st > end:
            g = st - end
            idle += g
            if g > 15:
                k = gaps[(prev.name.split("(")[0][-40:], e.name.split("(")[0][-40:])]; k[0] += 1; k[1] += g
        if end is None or en > end:
            end, prev = en, e
    print(f"GPU idle between kernels: {idle / n / 1e3:.2f} ms/step; gaps > 15 us by (previous kernel -> next kernel):")
    for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"  {v[1] / n:8.1f} us/step n/step={v[0] / n:5.1f}  {k[0]}  ->  {k[1]}")
    tot = sum(v[1] for v in kern.values())
    ours = sum(v[1] for k, v in kern.items() if "splatco" in k or k.startswith(("dec_", "blend_", "sort_", "preprocess", "visible_filter", "duplicate", "identify", "scan_block", "sgemm", "ta_", "loss_", "stat")))
    print(f"GPU kernel time {tot / n / 1e3:.2f} ms/step: library {ours / n / 1e3:.2f}, other {(tot - ours) / n / 1e3:.2f}")
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"  {v[1] / n:9.1f} us/step n/step={v[0] / n:6.1f}  {k}")
    print("host self time, top ops:")
    for k, v in sorted(cpu.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"  {v[1] / n:9.1f} us/step n/step={v[0] / n:6.1f}  {k}")


if __name__ == "__main__":
    main()
