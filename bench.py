#!/usr/bin/env python
"""bench.py — fwd+bwd ms/view of SplatCo's differentiable render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1]

A "step" is one training iteration's render work for the `--mv` multi-view batch (BASELINE.json
configs[1]: Tanks&Temples-shaped, ~1M Gaussians, 980x545, mv=4): for every view, forward
(prefilter -> decode -> preprocess -> binning/sort -> blend), L1 loss gradient, backward.  Views
are sharded across ranks when N > 1 (weak scaling: every rank renders its own mv views; per-Gaussian
gradients are all-reduced with NCCL after the local backward).

Printed JSON (one line, rank 0): metric fwd_bwd_ms_per_view (lower is better) = max-over-ranks step
time / views rendered by all ranks; `e2e` = the same through the public API with host inputs (pinned
ground-truth images copied H2D every view, loss read back D2H every step); `roofline` for the
dominant kernel; `cpu_baseline` = the oracle port on the host cores over one view.

`--impl reference` times the reference's own CPU implementation of the path: the reference rasterizer
is CUDA-only and absent and its Python decode cannot travel to the GPU box, so this
arm runs the oracle port (oracle/) on all host threads, one view per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (anchors N, K, W, H, mv)
    "c1": dict(N=10_000, K=10, W=256, H=256, mv=1, desc="C1 synthetic 10k anchors x10, 256x256, 1 view"),
    "c2": dict(N=100_000, K=10, W=980, H=545, mv=4, desc="C2 Tanks&Temples-shaped ~1M Gaussians, 980x545, mv=4"),
    "c3": dict(N=500_000, K=10, W=1152, H=864, mv=4, desc="C3 Mill19-Rubble-shaped ~5M Gaussians, 1152x864, mv=4"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, n in enumerate(names):
                    if any("Active" == r[5 + k].strip() for r in rows if len(r) >= 9):
                        out["reasons"].append(n)
                out["samples"] = len(sm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def build_scene(cfg, seed, device):
    """Synthetic Gaussians of the workload's shape.  Until the fused decode is wired in, the cloud is
    the decode's *output* shape: anchors U([-1,1]^3), K offsets each, ~55 % kept (opacity mask)."""
    from splatco_b200.synthetic import ring_cameras
    g = torch.Generator().manual_seed(20240 + seed)
    N, K = cfg["N"], cfg["K"]
    anchors = torch.rand(N, 3, generator=g) * 2 - 1
    s0 = 2.0 / N ** (1.0 / 3.0)
    ascale = s0 * torch.exp(torch.randn(N, 6, generator=g) * 0.3)
    offs = torch.randn(N, K, 3, generator=g) * 0.5
    keep = torch.rand(N * K, generator=g) < 0.55
    xyz = (anchors[:, None, :] + offs * ascale[:, None, :3]).reshape(-1, 3)[keep]
    M = xyz.shape[0]
    scales = (ascale[:, None, 3:].expand(N, K, 3).reshape(-1, 3)[keep] * torch.sigmoid(torch.randn(M, 3, generator=g))) * 0.5
    q = torch.randn(M, 4, generator=g)
    rots = q / q.norm(dim=1, keepdim=True)
    opac = torch.empty(M, 1).uniform_(0.02, 0.95, generator=g)
    colors = torch.rand(M, 3, generator=g)
    cams = ring_cameras(cfg["mv"], cfg["W"], cfg["H"])
    gts = [torch.rand(3, cfg["H"], cfg["W"], generator=g) for _ in cams]
    t = dict(xyz=xyz, colors=colors, opac=opac, scales=scales.contiguous(), rots=rots)
    return {k: v.to(device) for k, v in t.items()}, cams, gts


def make_settings(cam, bg, device):
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=math.tan(cam.FoVx * 0.5),
        tanfovy=math.tan(cam.FoVy * 0.5), bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform.to(device),
        projmatrix=cam.full_proj_transform.to(device), sh_degree=1, campos=cam.camera_center.to(device),
        prefiltered=False, debug=False)


def run_ours(args):
    import torch.distributed as dist
    from splatco_b200 import _lib, profiling
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = WORKLOADS[args.workload]
    L = _lib.lib()
    params, cams, gts = build_scene(cfg, seed=1, device=device)
    # each rank renders its own mv views (weak scaling): rotate the camera ring per rank
    if world > 1:
        from splatco_b200.synthetic import ring_cameras
        cams = ring_cameras(cfg["mv"], cfg["W"], cfg["H"], phase=2 * math.pi * rank / (world * cfg["mv"]))
    bg = torch.ones(3, device=device)
    leaves = {k: params[k].clone().requires_grad_() for k in ("xyz", "colors", "opac", "scales", "rots")}
    rasts = [GaussianRasterizer(make_settings(c, bg, device)) for c in cams]
    gts_dev = [g.to(device) for g in gts]
    gts_pinned = [g.pin_memory() for g in gts]
    H, W, mv = cfg["H"], cfg["W"], cfg["mv"]
    M = leaves["xyz"].shape[0]
    flat_grads = None

    def step(host_inputs: bool):
        total = None
        for p in leaves.values():
            p.grad = None
        for v in range(mv):
            gt = gts_pinned[v].to(device, non_blocking=True) if host_inputs else gts_dev[v]
            m2d = torch.zeros_like(leaves["xyz"], requires_grad=True)
            img, radii = rasts[v](means3D=leaves["xyz"], means2D=m2d, shs=None, colors_precomp=leaves["colors"],
                                  opacities=leaves["opac"], scales=leaves["scales"], rotations=leaves["rots"],
                                  cov3D_precomp=None)
            loss = (img - gt).abs().mean()
            total = loss if total is None else total + loss
        total.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in leaves.values()])
            dist.all_reduce(flat)
        return total.item() if host_inputs else total

    def timed(n, host_inputs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(host_inputs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.splatco_launch_count()
    with profiling.collect() as prof:
        ms_dev = timed(args.steps, host_inputs=False)
    launches = (L.splatco_launch_count() - launches0)
    stage_sum = prof.summary()
    ms_e2e = timed(args.steps, host_inputs=True)
    clocks = sampler.stop() if rank == 0 else None

    views = mv * world
    ms_step = ms_dev / args.steps
    value = ms_step / views
    e2e_value = (ms_e2e / args.steps) / views
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # roofline of the dominant kernel (algorithmic bytes per launch / mean launch time)
    with torch.no_grad():
        R_list = []
        from splatco_b200.diff_gaussian_rasterization import rasterize_forward_state
        for v in range(mv):
            _, _, st = rasterize_forward_state(leaves["xyz"], leaves["colors"], leaves["opac"], leaves["scales"],
                                               leaves["rots"], rasts[v].raster_settings)
            R_list.append(st.R)
    R_mean = float(np.mean(R_list))
    HW = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    passes = math.ceil((32 + max(1, (T - 1).bit_length())) / 8)
    alg_bytes = {
        "preprocess_fwd": 104.0 * M,
        "binning": 12.0 * R_mean + passes * 24.0 * R_mean + 8.0 * R_mean + 8.0 * R_mean + 8.0 * T,
        "blend_fwd": 40.0 * R_mean + 20.0 * HW,
        "blend_bwd": 76.0 * R_mean + 32.0 * HW,
        "preprocess_bwd": 200.0 * M,
    }
    peak, peak_src = peaks()
    stages = {}
    for k, (n, tot) in stage_sum.items():
        avg = tot / max(n, 1)
        stages[k] = {"calls": n, "avg_ms": round(avg, 4), "share": round(tot / ms_dev, 4),
                     "alg_gbs": round(alg_bytes.get(k, 0.0) / (avg * 1e-3) / 1e9, 1) if avg > 0 else None}
    dom = max(stage_sum.items(), key=lambda kv: kv[1][1])[0] if stage_sum else None
    roof = None
    if dom:
        a = stages[dom]["alg_gbs"]
        roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                "frac": round(a / peak, 4), "peak_source": peak_src, "traffic": None,
                "pairs_per_s": None}
    out = {
        "metric": "fwd_bwd_ms_per_view", "value": round(value, 4), "unit": "ms/view", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "path": "rasterizer (preprocess+binning+blend fwd/bwd); decode pending",
                   "gaussians": int(M), "instances_R": int(R_mean), "mv": mv, "views_per_step": views,
                   "loss": "L1 (torch elementwise)", "l2": "inputs > L2 not guaranteed; L2 not flushed between views",
                   "parallelism": f"view-sharded dp{world}"},
        "it_per_s": round(1000.0 / ms_step, 3),
        "e2e": {"value": round(e2e_value, 4), "unit": "ms/view", "h2d_bytes_per_step": int(mv * 3 * HW * 4),
                "d2h_bytes_per_step": 4 + 4 * mv},
        "gpu_launches": int(launches),
        "roofline": roof, "stages": stages, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(cfg, params, cams, gts, views=1)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_view(params_np, cam, gt, bg):
    """One view, forward + backward, on the CPU oracle (all host threads)."""
    from oracle import raster as R
    tx, ty = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
    H, W = cam.image_height, cam.image_width
    view, proj = cam.world_view_transform.numpy(), cam.full_proj_transform.numpy()
    fw = R.rasterize_forward(params_np["xyz"], params_np["colors"], params_np["opac"], params_np["scales"],
                             params_np["rots"], 1.0, view, proj, tx, ty, H, W, bg)
    dL = (np.sign(fw["image"] - gt) / (3.0 * H * W)).astype(np.float32)
    R.rasterize_backward(fw, params_np["xyz"], params_np["colors"], params_np["scales"], params_np["rots"], 1.0,
                         view, proj, tx, ty, H, W, bg, dL)
    return fw["bn"].R


def cpu_baseline(cfg, params, cams, gts, views=1):
    from oracle import raster as R
    pn = {k: v.detach().cpu().numpy() for k, v in params.items()}
    bg = np.ones(3, np.float32)
    t0 = time.perf_counter()
    for v in range(views):
        cpu_view(pn, cams[v], gts[v].numpy(), bg)
    dt = (time.perf_counter() - t0) / views
    return {"value": round(dt * 1e3, 2), "unit": "ms/view", "cores": int(R.lib().oracle_get_threads()),
            "kind": "port", "sample": f"{views} view(s) of the same workload, fwd+bwd, oracle/raster_oracle.c on all host threads"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path = the oracle port (see module doc)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import raster as R
    cfg = WORKLOADS[args.workload]
    params, cams, gts = build_scene(cfg, seed=1, device="cpu")
    pn = {k: v.numpy() for k, v in params.items()}
    bg = np.ones(3, np.float32)
    budget_s = 240.0
    t_start = time.perf_counter()
    for i in range(min(args.warmup, 1)):
        cpu_view(pn, cams[0], gts[0].numpy(), bg)
    times = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        cpu_view(pn, cams[i % len(cams)], gts[i % len(cams)].numpy(), bg)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    ms = float(np.mean(times)) * 1e3
    cores = int(R.lib().oracle_get_threads())
    out = {"impl": "reference", "metric": "fwd_bwd_ms_per_view", "value": round(ms, 2), "unit": "ms/view",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(times), "warmup": min(args.warmup, 1),
           "ms_per_step": round(ms, 2), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["desc"], "path": "rasterizer (preprocess+binning+blend fwd/bwd); decode pending",
                      "gaussians": int(pn["xyz"].shape[0]), "mv": cfg["mv"]},
           "cpu_baseline": {"value": round(ms, 2), "unit": "ms/view", "cores": cores, "kind": "port",
                            "sample": "one view of the workload per step, fwd+bwd, oracle port on all host threads"},
           "e2e": {"value": round(ms, 2), "unit": "ms/view", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
