#!/usr/bin/env python
"""A/B of the blend kernel implementations on one view of a bench workload (splatco_blend_set_impl):
device time of each variant and the largest difference of its outputs against variant 1.
    python tools/blend_ab.py [--workload c2] [--reps 20]
Diagnostic only (not a bench value)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fwd", default="1")
    ap.add_argument("--bwd", default="1,2,3")
    a = ap.parse_args()
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    from splatco_b200.diff_gaussian_rasterization import rasterize_forward_state
    from splatco_b200.gaussian_renderer import _settings, generate_neural_gaussians, prefilter_voxel
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    cfg = bench.WORKLOADS[a.workload]
    pc = bench.build_model(cfg, dev)
    pc.feat_planes.Q0 = 0.0
    cams, gts = bench.build_views(cfg)
    cam = cams[0].to(dev)
    bg = torch.ones(3, device=dev)
    H, W = cfg["H"], cfg["W"]
    with torch.no_grad():
        vm = prefilter_voxel(cam, pc, bench.PIPE, bg)
        xyz, color, opacity, scaling, rot, _, _ = generate_neural_gaussians(cam, pc, vm, is_training=True)
        settings = _settings(cam, bench.PIPE, bg, 1.0)
        img, radii, st = rasterize_forward_state(xyz, color, opacity, scaling, rot, settings)
    stream = _lib.raw_stream(dev)
    P, R = st.P, st.R
    gt = gts[0].to(dev)
    dL = (torch.sign(img - gt) * 0.8 / img.numel() + torch.randn_like(img) * 1e-7).contiguous()
    col = torch.empty_like(img)

    def time_ms(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {"workload": a.workload, "P": int(P), "R": int(R), "fwd": {}, "bwd": {}}
    ref_img = None
    for f in [int(x) for x in a.fwd.split(",")]:
        check(L.splatco_blend_set_impl(f, 0), "set_impl")
        fn = lambda: check(L.splatco_blend_fwd(R, H, W, ptr(bg), ptr(st.geom), ptr(st.binning), ptr(st.image), ptr(col), stream), "blend_fwd")
        ms = time_ms(fn, a.reps)
        rec = {"ms": round(ms, 4)}
        im = st.image.clone()
        if ref_img is None:
            ref_img, ref_state = col.clone(), im
        else:
            rec["image_max_abs_vs_1"] = float((col - ref_img).abs().max())
            rec["state_bytes_differ"] = int((im.view(torch.uint8) != ref_state.view(torch.uint8)).sum())
        out["fwd"][str(f)] = rec
    check(L.splatco_blend_set_impl(int(a.fwd.split(",")[0]), 0), "set_impl")
    check(L.splatco_blend_fwd(R, H, W, ptr(bg), ptr(st.geom), ptr(st.binning), ptr(st.image), ptr(col), stream), "blend_fwd")
    g = [torch.zeros(P, c, device=dev) for c in (3, 3, 1, 3)]
    t_zero = time_ms(lambda: [t.zero_() for t in g], a.reps)
    ref = None
    for b in [int(x) for x in a.bwd.split(",")]:
        check(L.splatco_blend_set_impl(0, b), "set_impl")

        def fn():
            for t in g:
                t.zero_()
            check(L.splatco_blend_bwd(P, R, H, W, ptr(bg), ptr(st.geom), ptr(st.binning), ptr(st.image), ptr(dL), *[ptr(t) for t in g], stream), "blend_bwd")

        ms = time_ms(fn, a.reps) - t_zero
        rec = {"ms": round(ms, 4)}
        cur = [t.double().clone() for t in g]
        if ref is None:
            ref = cur
        else:
            for name, x, y in zip(("mean2D", "conic", "opacity", "color"), cur, ref):
                scale = y.abs().max().item() + 1e-30
                rec[name + "_max_abs_over_max"] = float((x - y).abs().max().item() / scale)
                rel = ((x - y).abs() / (y.abs() + 1e-3 * scale))
                rec[name + "_max_rel(floor 1e-3 max)"] = float(rel.max().item())
        out["bwd"][str(b)] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
