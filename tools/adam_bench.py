#!/usr/bin/env python
"""Host and device cost of one FusedAdam.step() over the C2 model's parameters (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from splatco_b200.optim import FusedAdam

dev = torch.device("cuda", 0)
pc = bench.build_model(bench.WORKLOADS["c2"], dev)
params = [p for p in pc.parameters() if p.requires_grad]
flat = torch.randn(sum(p.numel() for p in params) + 64 * len(params), device=dev) * 1e-3
o = 0
for p in params:
    p.grad = flat[o:o + p.numel()].view(p.shape)
    o += (p.numel() + 63) // 64 * 64
print("tensors", len(params), "elements", sum(p.numel() for p in params))
for name, opt in (("fused", FusedAdam([{"params": [p], "lr": 1e-6} for p in params], lr=0.0, eps=1e-15)),
                  ("torch foreach", torch.optim.Adam([{"params": [p], "lr": 1e-6} for p in params], lr=0.0, eps=1e-15)),
                  ("torch fused", torch.optim.Adam([{"params": [p], "lr": 1e-6} for p in params], lr=0.0, eps=1e-15, fused=True))):
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    e0.record()
    for _ in range(20):
        t = time.perf_counter(); opt.step(); host += time.perf_counter() - t
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: host {host / 20 * 1e3:.3f} ms/step, device span {e0.elapsed_time(e1) / 20:.3f} ms/step")
