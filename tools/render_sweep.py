#!/usr/bin/env python
"""Render-only sweep of BASELINE.json configs[4]: M in {1, 2, 5, 10, 20} M Gaussians fed straight
to the rasterizer at 3840x2160 — means uniform in the frustum at depth U(2,6), log-uniform scales giving sigma_px in
[0.5, 4], opacity U(0.05, 0.9), random colours / unit quaternions (splatco_b200.synthetic.random_gaussians).
Forward only (preprocess + binning + blend) through the public GaussianRasterizer, CUDA-event timed, one JSON line per M:
M, visible P, R, R/M, mean contributors per pixel, ms/view, Gaussians/s, instances/s and the stage split.
Views shard across GPUs with no collective (render.py renders independent views): `--gpus N` runs N independent
replicas, so the multi-GPU number is N x this one.

    python tools/render_sweep.py [--sizes 1,2,5,10,20] [--iters 10] [--width 3840 --height 2160]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from splatco_b200 import profiling
from splatco_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from splatco_b200.synthetic import random_gaussians, ring_cameras


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,2,5,10,20")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    W, H = a.width, a.height
    cam = ring_cameras(3, W, H)[1]
    st = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=torch.ones(3, device=dev), scale_modifier=1.0, viewmatrix=cam.world_view_transform.to(dev),
        projmatrix=cam.full_proj_transform.to(dev), sh_degree=1, campos=cam.camera_center.to(dev), prefiltered=False, debug=False)
    rast = GaussianRasterizer(raster_settings=st)
    for m in [float(s) for s in a.sizes.split(",")]:
        M = int(m * 1e6)
        means, colors, opac, scales, rots = [t.to(dev) for t in random_gaussians(M, cam, 500 + int(m))]
        means2d = torch.zeros_like(means)

        def view():
            return rast(means3D=means, means2D=means2d, shs=None, colors_precomp=colors, opacities=opac, scales=scales,
                        rotations=rots, cov3D_precomp=None)
        with torch.no_grad():
            for _ in range(3):
                img, radii = view()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                img, radii = view()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            with profiling.collect() as prof:
                for _ in range(3):
                    view()
                torch.cuda.synchronize()
            stages = {k: round(tot / max(n, 1), 4) for k, (n, tot) in prof.summary().items()}
            from splatco_b200.diff_gaussian_rasterization import rasterize_forward_state
            _, _, state = rasterize_forward_state(means, colors, opac, scales, rots, st)
            R = int(state.R)
            ncontrib = None
            try:
                from tests.util import chunk, layout
                off = layout("image", H, W)
                ncontrib = float(chunk(state.image, off[2], torch.int32, H * W).float().mean().item())
            except Exception:
                pass
        print(json.dumps({"workload": f"C5 render-only {W}x{H}", "gaussians_M": M, "visible_P": int((radii > 0).sum().item()),
                          "instances_R": R, "R_per_M": round(R / M, 3), "mean_contributors_per_pixel": ncontrib,
                          "ms_per_view": round(ms, 4), "gaussians_per_s": round(M / (ms * 1e-3), 1),
                          "instances_per_s": round(R / (ms * 1e-3), 1), "stages_ms": stages,
                          "image_in_unit_range": bool(img.min().item() >= 0.0 and img.max().item() <= 1.0 + 1e-5),
                          "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
        del means, colors, opac, scales, rots, means2d, img, radii, state
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
