"""Phase timing of the decode MLP kernels (splatco_decode_profile(2)): CTA 0 stamps clock64 at its phase boundaries.
    python tools/decode_trace.py [workload]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(name="c2"):
    from splatco_b200 import _lib
    from splatco_b200.gaussian_renderer import generate_neural_gaussians, prefilter_voxel
    L = _lib.lib()
    dev = torch.device("cuda")
    cfg = bench.WORKLOADS[name]
    pc = bench.build_model(cfg, dev)
    cams, _ = bench.build_views(cfg)
    cam = cams[0].to(dev)
    bg = torch.ones(3, device=dev)
    L.splatco_decode_profile(2)
    for _ in range(3):
        for p in pc.parameters():
            p.grad = None
        vm = prefilter_voxel(cam, pc, bench.PIPE, bg)
        outs = generate_neural_gaussians(cam, pc, vm, is_training=True)
        sum(o.sum() for o in outs[:5]).backward()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)()
    assert L.splatco_decode_trace_read(buf) == 0
    t = list(buf)
    FWD = ["wait U", "split", "sync A + stage 1", "epilogue 1", "sync B + stage 2", "epilogue 2"]
    for it in range(4):
        s = t[8 * it: 8 * it + 7]
        if s[6] == 0:
            break
        print(f"fwd tile {it}: " + ", ".join(f"{n} {(b - a) / 1.965e3:.2f}us" for n, a, b in zip(FWD, s[:-1], s[1:])) + f" | total {(s[6] - s[0]) / 1.965e3:.2f}us")
    CT = ["slice0 landed", "slice0 issued", "slice1 landed", "slice1 issued", "slice2 landed (after slice0 done + its copy)", "slice2 issued", "W2 copy queued"]
    for it in range(4):
        c = t[32 + 8 * it: 32 + 8 * it + 8]
        if c[7] == 0:
            break
        print(f"fwd control tile {it}: " + ", ".join(f"{n} +{(b - a) / 1.965e3:.2f}us" for n, a, b in zip(CT, c[:-1], c[1:])))
    BWD = ["P0 compute", "wait prev g", "write dZ", "sync1 + b1", "w2 q0", "w2 q1", "w2 q2", "w2 q3 (+dH regs)", "wait last w2", "write dH", "sync6 + b2",
           "epilogue b2", "g q0", "g q1", "g q2"]
    for it in range(4):
        s = t[64 + 16 * it: 64 + 16 * it + 15]
        if len(s) < 15 or s[14] == 0:
            break
        print(f"bwd tile {it}: " + ", ".join(f"{n} {(b - a) / 1.965e3:.2f}us" for n, a, b in zip(BWD, s[:-1], s[1:])))
        if it + 1 < 4 and t[64 + 16 * (it + 1)]:
            print(f"          g q3 + loop {(t[64 + 16 * (it + 1)] - s[14]) / 1.965e3:.2f}us | tile total {(t[64 + 16 * (it + 1)] - s[0]) / 1.965e3:.2f}us")


if __name__ == "__main__":
    main(*sys.argv[1:])
