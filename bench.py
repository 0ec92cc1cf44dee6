#!/usr/bin/env python
"""bench.py — fwd+bwd ms/view of SplatCo's differentiable render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one training iteration's render work for the `--mv` multi-view batch (BASELINE.json
configs[1]: Tanks&Temples-shaped, 100k anchors x 10 offsets ~ 1M candidate Gaussians, 980x545, mv=4):
for every view   prefilter_voxel -> generate_neural_gaussians (fused decode) -> preprocess ->
binning/sort -> blend,   the reference's per-view loss terms that touch the hot path's outputs
(0.8*L1 + 0.2*(1-SSIM) + 0.01*mean(prod(scaling)), train.py:192-196; the image part is the fused
l1_ssim_loss of SURVEY §8 row f3),   then ONE
backward over the summed loss (train.py:240), through the drop-in `gaussian_renderer.render()`.
With N > 1 every rank renders its own mv views (weak scaling, views sharded by rank) and the
parameter gradients are all-reduced with NCCL.

Printed JSON (one line, rank 0): metric fwd_bwd_ms_per_view (lower is better) = max-over-ranks step
time / views rendered by all ranks, inputs resident in HBM; `e2e` = the same with host inputs (each
view's ground-truth image copied H2D from pinned memory, the loss read back D2H every step);
`roofline` for the stage with the largest share; `cpu_baseline` = the oracle port on the host cores
over one view.

`--impl reference`: the reference's rasterizer is CUDA-only and absent from the mount and its Python
decode cannot travel to the GPU box, so this arm times the oracle port of the
same path (oracle/decode_oracle.py on torch CPU threads + oracle/raster_oracle.c on all host
threads), one view of the workload per step.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    "c1": dict(N=10_000, K=10, W=256, H=256, mv=1, plane_size=512, C=15,
               desc="C1 synthetic 10k anchors x10 offsets, 256x256, 1 view"),
    "c2": dict(N=100_000, K=10, W=980, H=545, mv=4, plane_size=2500, C=15,
               desc="C2 Tanks&Temples-shaped 100k anchors x10 offsets (~1M Gaussians), 980x545, mv=4"),
    "c3": dict(N=500_000, K=10, W=1152, H=864, mv=4, plane_size=2800, C=15,
               desc="C3 Mill19-Rubble-shaped 500k anchors x10 (~5M Gaussians), 1152x864, plane_size 2800, mv=4"),
    "c4": dict(N=1_000_000, K=10, W=1920, H=1080, mv=1, plane_size=2800, C=15,
               desc="C4 MatrixCity-Aerial-shaped 1M anchors x10 (~10M Gaussians), 1920x1080, one view per GPU (mv=8 over 8 GPUs)"),
}
SCALE_FACTOR = 0.5       # anchor scale = 0.5 / N^(1/3): projected sigma ~0.5-4 px at these resolutions
LEVEL = 2                # activate_level in steady state (train.py:305-307)


def ncu_evidence():
    """Per-kernel numbers of the latest committed `ncu --set full` capture (profiles/*_traffic.json, written by
    tools/collect_profiles.py): dram bytes per launch, issue-active, tensor-pipe-active."""
    d = os.path.join(ROOT, "profiles")
    files = sorted(f for f in os.listdir(d) if f.endswith("_traffic.json")) if os.path.isdir(d) else []
    if not files:
        return {}
    ev = json.load(open(os.path.join(d, files[-1])))
    ev["_file"] = "profiles/" + files[-1]
    return ev


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            rows = [r for r in rows if len(r) >= 9]
            if rows:
                out["sm_mhz"] = float(np.median([float(r[1]) for r in rows]))
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
                for k, n in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                    if any(r[5 + k].strip() == "Active" for r in rows):
                        out["reasons"].append(n)
                out["samples"] = len(rows)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def build_model(cfg, device, seed=1):
    from splatco_b200.model import AnchorModel
    pc = AnchorModel(cfg["N"], n_offsets=cfg["K"], plane_size=cfg["plane_size"], num_channels=cfg["C"],
                     device=device, seed=20240 + seed, scale_factor=SCALE_FACTOR)
    pc.feat_planes._feat.activate_level = LEVEL
    pc.train()
    return pc


def build_views(cfg, rank=0, world=1, seed=1):
    from splatco_b200.synthetic import ring_cameras
    phase = 2 * math.pi * rank / (world * cfg["mv"]) if world > 1 else 0.0
    cams = ring_cameras(cfg["mv"], cfg["W"], cfg["H"], phase=phase)
    g = torch.Generator().manual_seed(777 + seed + rank)
    gts = [torch.rand(3, cfg["H"], cfg["W"], generator=g) for _ in cams]
    return cams, gts


PIPE = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)


def run_ours(args):
    import torch.distributed as dist
    from splatco_b200 import _lib, profiling
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.loss import l1_ssim_loss, scaling_reg
    from splatco_b200.multiview import GradBucket

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cfg = WORKLOADS[args.workload]
    L = _lib.lib()
    pc = build_model(cfg, device)
    pc.feat_planes.Q0 = 0.03                       # training-time plane-feature noise (timing config)
    cams, gts = build_views(cfg, rank, world)
    cams = [c.to(device) for c in cams]
    bg = torch.ones(3, device=device)
    gts_dev = [g.to(device) for g in gts]
    gts_pinned = [g.pin_memory() for g in gts]
    params = pc.parameters()
    bucket = GradBucket(params) if world > 1 else None
    H, W, mv = cfg["H"], cfg["W"], cfg["mv"]
    info = {}

    # e2e: each step's ground-truth images leave pinned host memory inside the timed region, on a copy stream into
    # fixed staging buffers (what a data loader does); a view waits for its own image only when it computes its loss
    copy_stream = torch.cuda.Stream(device)
    gts_stage = [torch.empty_like(g) for g in gts_dev]
    copied = [torch.cuda.Event() for _ in range(mv)]

    def step(host_inputs: bool):
        for p in params:
            p.grad = None
        total = None
        Ms, Vs = [], []
        if host_inputs:
            copy_stream.wait_stream(torch.cuda.current_stream(device))     # the previous step no longer reads the buffers
            with torch.cuda.stream(copy_stream):
                for v in range(mv):
                    gts_stage[v].copy_(gts_pinned[v], non_blocking=True)
                    copied[v].record(copy_stream)
        for v in range(mv):
            vm = prefilter_voxel(cams[v], pc, PIPE, bg)
            pkg = render(cams[v], pc, PIPE, bg, visible_mask=vm, retain_grad=True)
            if host_inputs:
                torch.cuda.current_stream(device).wait_event(copied[v])
                gt = gts_stage[v]
            else:
                gt = gts_dev[v]
            # the reference's per-view loss (train.py:192-196, lambda_dssim = 0.2 from arguments/__init__.py), image part fused
            loss = l1_ssim_loss(pkg["render"], gt, 0.2) + 0.01 * (pkg["scaling"].prod(dim=1).mean() if args.torch_scaling_reg else scaling_reg(pkg["scaling"]))
            total = loss if total is None else total + loss
            Ms.append(pkg["radii"].shape[0]); Vs.append(pkg["selection_mask"].shape[0] // cfg["K"])
        total.backward()
        if bucket is not None:
            bucket.allreduce()
        info["M"], info["V"] = float(np.mean(Ms)), float(np.mean(Vs))
        return float(total.item()) if host_inputs else total

    def timed(n, host_inputs):
        return _timed_fn(step, n, host_inputs)

    host_ms = {}

    def _timed_fn(fn, n, host_inputs):
        # the cyclic collector is paused inside the timed region (a generation-2 pass over the autograd graphs of a
        # step costs tens of ms and lands on a random step); per-step host times are kept as a diagnostic
        gc.collect()
        gc.disable()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = [time.perf_counter()]
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        e0.record()
        for i in range(n):
            fn(host_inputs)
            marks[i].record()
            ts.append(time.perf_counter())
        e1.record()
        torch.cuda.synchronize()
        gc.enable()
        d = sorted((b - a) * 1e3 for a, b in zip(ts[:-1], ts[1:]))
        gd = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))
        host_ms[getattr(fn, "__name__", "fn") + ("_host_inputs" if host_inputs else "")] = {
            "host_median": round(d[len(d) // 2], 3), "host_max": round(d[-1], 3),
            "gpu_median": round(gd[len(gd) // 2], 3), "gpu_max": round(gd[-1], 3)}
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    # the clock sampler starts BEFORE the warm-up: nvidia-smi's start-up (NVML initialisation) stalls the GPU for tens
    # of ms on some boxes, which otherwise lands inside the first timed loop; it keeps sampling through the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(warm):
        step(False)
    launches0 = L.splatco_launch_count()
    ms_dev = timed(args.steps, host_inputs=False)
    launches = L.splatco_launch_count() - launches0
    ms_e2e = timed(args.steps, host_inputs=True)
    clocks = sampler.stop() if rank == 0 else None
    # per-stage CUDA-event times come from a separate pass (two event records per stage would otherwise
    # sit inside the headline number); shares are relative to this pass's own step time
    with profiling.collect() as prof:
        ms_stage = timed(args.steps, host_inputs=False)
    stage_sum = prof.summary()
    # whole training iteration (BASELINE metric "train it/s"): the same step + the optimizer update the reference performs
    # at train.py:310-312 (Adam over every parameter group, eps 1e-15, scene/gaussian_model.py:519-572), fused
    from splatco_b200.optim import FusedAdam
    opt = FusedAdam([{"params": [p], "lr": 1e-6, "name": f"p{i}"} for i, p in enumerate(params) if p.requires_grad], lr=0.0, eps=1e-15)

    def train_iter(host_inputs):
        loss = step(host_inputs)
        opt.step()
        return loss
    for _ in range(2):
        train_iter(False)
    launches_t0 = L.splatco_launch_count()
    ms_train = _timed_fn(train_iter, args.steps, True)
    launches_train = L.splatco_launch_count() - launches_t0

    views = mv * world
    ms_step = ms_dev / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # instance count R of each view (for the algorithmic-byte roofline figures)
    from splatco_b200.diff_gaussian_rasterization import rasterize_forward_state
    from splatco_b200.gaussian_renderer import _settings, generate_neural_gaussians
    R_list = []
    with torch.no_grad():
        for v in range(mv):
            vm = prefilter_voxel(cams[v], pc, PIPE, bg)
            xyz, color, opacity, scaling, rot, _, _ = generate_neural_gaussians(cams[v], pc, vm, is_training=True)
            _, _, st = rasterize_forward_state(xyz, color, opacity, scaling, rot, _settings(cams[v], PIPE, bg, 1.0))
            R_list.append(st.R)
    R, M, V, N = float(np.mean(R_list)), info["M"], info["V"], cfg["N"]
    HW = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    passes = math.ceil((32 + max(1, (T - 1).bit_length())) / 8)
    dec_bytes = V * (284 + 240 * (LEVEL + 2) + 50) + 56 * M          # SURVEY §8d decode fwd per visible anchor
    alg_bytes = {                                                      # SURVEY §8d / BASELINE.md §3.4
        "visible_filter": 48.0 * N,
        "decode_fwd": dec_bytes, "decode_emit": 56.0 * M, "decode_bwd": 2.0 * dec_bytes,
        "preprocess_fwd": 104.0 * M,
        "binning": 12.0 * R + passes * 24.0 * R + 8.0 * R + 8.0 * R + 8.0 * T,
        "blend_fwd": 40.0 * R + 20.0 * HW,
        "blend_bwd": 76.0 * R + 32.0 * HW,
        "preprocess_bwd": 200.0 * M,
    }
    peak, peak_src = peaks()
    stages = {}
    for k, (n, tot) in stage_sum.items():
        avg = tot / max(n, 1)
        st = {"calls": n, "avg_ms": round(avg, 4), "share": round(tot / ms_stage, 4)}
        if k in alg_bytes and avg > 0:
            st["alg_gbs"] = round(alg_bytes[k] / (avg * 1e-3) / 1e9, 1)
        stages[k] = st
    timed_k = [k for k in stage_sum if k in alg_bytes]
    dom = max(timed_k, key=lambda k: stage_sum[k][1]) if timed_k else None
    roof = None
    ncu = ncu_evidence()
    if dom:
        a = stages[dom]["alg_gbs"]
        ev = ncu.get(dom + "_kernel")
        roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s",
                "frac": round(a / peak, 4), "peak_source": peak_src,
                "traffic": int(ev["dram_bytes"]) if ev else None,      # dram read+write per launch, ncu --set full
                "traffic_source": ncu.get("_file") if ev else None,
                "alg_bytes_per_launch": int(alg_bytes[dom])}
        if dom.startswith("blend"):
            # the blend is FP32-issue bound, not HBM bound: also report (pixel, splat) pairs/s
            roof["pairs_per_s"] = round(R * 256 / (stages[dom]["avg_ms"] * 1e-3), 1)
    # the decode MLP runs on tcgen05 (3xTF32): algorithmic FLOPs counted once, against half the measured bf16 peak
    mlp = None
    ev = ncu.get("dec_tc_fwd_kernel")
    if ev:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] / 2 if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1100.0
        flops = V * (32.5e3 + 5.5e3 * LEVEL)
        ach = flops / (ev["time_us"] * 1e-6) / 1e12
        mlp = {"kernel": "dec_tc_fwd_kernel", "bound": "tensor", "achieved": round(ach, 2), "peak": round(pk, 1),
               "unit": "TFLOP/s", "frac": round(ach / pk, 4), "tensor_pipe_active_pct": ev["tensor_active_pct"],
               "source": ncu.get("_file"), "note": "kernel time from the committed ncu capture, not from this run"}
    out = {
        "metric": "fwd_bwd_ms_per_view", "value": round(ms_step / views, 4), "unit": "ms/view", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "path": "prefilter_voxel + render() drop-in: decode, preprocess, binning, blend, fwd+bwd",
                   "anchors": N, "visible_anchors": int(V), "gaussians": int(M), "instances_R": int(R),
                   "activate_level": LEVEL, "plane_size": cfg["plane_size"], "num_channels": cfg["C"], "Q0": 0.03,
                   "mv": mv, "views_per_step": views, "loss": "0.8*L1 + 0.2*(1-SSIM) (fused kernels) + 0.01*mean(prod(scaling)) (" + ("torch ops" if args.torch_scaling_reg else "splatco scaling_reg kernel, no host sync in its backward") + "), train.py:192-196",
                   "l2": "per-step working set (planes + workspaces) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"view-sharded dp{world}, NCCL grad all-reduce" if world > 1 else "single GPU"},
        "it_per_s": round(1000.0 / ms_step, 3),
        "e2e": {"value": round((ms_e2e / args.steps) / views, 4), "unit": "ms/view",
                "h2d_bytes_per_step": int(mv * 3 * HW * 4), "d2h_bytes_per_step": 4 + 8 * mv},
        "gpu_launches": int(launches),
        "train": {"it_per_s": round(1000.0 * args.steps / ms_train, 3), "ms_per_iter": round(ms_train / args.steps, 4),
                  "gpu_launches": int(launches_train),
                  "includes": "the e2e step (H2D ground truth, mv views fwd+bwd, loss read-back" + (", NCCL grad all-reduce" if world > 1 else "") +
                              ") + FusedAdam update of every parameter (train.py:310-312), lr 1e-6"},
        "roofline": roof, "decode_mlp": mlp, "stages": stages, "clocks": clocks,
        "per_step_ms": host_ms,
    }
    if bucket is not None:
        out["config"]["allreduce_bytes_per_step"] = bucket.nbytes()
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(cfg)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---- CPU oracle port (cpu_baseline leg and --impl reference) -------------------------------------------
class _CpuScene:
    def __init__(self, cfg):
        self.cfg = cfg
        self.pc = build_model(cfg, "cpu")
        self.pc.feat_planes.Q0 = 0.0
        self.cams, self.gts = build_views(cfg)
        self.p = {"feat." + k: v for k, v in self.pc.feat_planes._feat.state_dict().items()}
        for name in ("mlp_opacity", "mlp_cov", "mlp_color"):
            self.p.update({f"{name}.{k}": v for k, v in getattr(self.pc, name).state_dict().items()})
        for k, v in self.p.items():
            if v.dtype.is_floating_point and "running" not in k and "xyz_m" not in k:
                v.requires_grad_(True)

    def view(self, i):
        """prefilter + decode + rasterize forward, L1 gradient, rasterize + decode backward; returns seconds."""
        from oracle import decode_oracle as D
        from oracle import raster as R
        cfg, pc = self.cfg, self.pc
        cam, gt = self.cams[i % len(self.cams)], self.gts[i % len(self.gts)].numpy()
        H, W = cfg["H"], cfg["W"]
        tx, ty = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        view, proj = cam.world_view_transform.numpy(), cam.full_proj_transform.numpy()
        bg = np.ones(3, np.float32)
        t0 = time.perf_counter()
        scaling = torch.exp(pc._scaling)
        radii = R.visible_filter(pc._anchor.detach().numpy(), scaling.detach().numpy()[:, :3],
                                 torch.nn.functional.normalize(pc._rotation).numpy(), 1.0, view, proj, tx, ty, H, W)
        vis = torch.from_numpy(radii > 0)
        xyz, color, opacity, scl, rot, _, _ = D.decode(self.p, pc._anchor_feat, pc._anchor, pc._offset, scaling, vis,
                                                       cam.camera_center, LEVEL, cfg["K"])
        a = [t.detach().numpy() for t in (xyz, color, opacity, scl, rot)]
        fw = R.rasterize_forward(a[0], a[1], a[2], a[3], a[4], 1.0, view, proj, tx, ty, H, W, bg)
        dL = (np.sign(fw["image"] - gt) / (3.0 * H * W)).astype(np.float32)
        g = R.rasterize_backward(fw, a[0], a[1], a[3], a[4], 1.0, view, proj, tx, ty, H, W, bg, dL)
        outs = [xyz, color, opacity, scl, rot]
        grads = [torch.from_numpy(g[k]) for k in ("means3D", "colors", "opacities", "scales", "rotations")]
        torch.autograd.backward(outs, grads)
        for t in [pc._anchor_feat, pc._anchor, pc._offset, pc._scaling] + list(self.p.values()):
            t.grad = None
        return time.perf_counter() - t0


def cpu_baseline(cfg):
    from oracle import raster as R
    torch.set_num_threads(os.cpu_count() or 1)
    sc = _CpuScene(cfg)
    dt = sc.view(0)
    return {"value": round(dt * 1e3, 2), "unit": "ms/view", "cores": int(R.lib().oracle_get_threads()),
            "kind": "port", "sample": "1 view of the same workload, fwd+bwd: oracle/decode_oracle.py (torch CPU, "
                                      "autograd) + oracle/raster_oracle.c on all host threads"}


def run_reference(args):
    """Reference arm = the oracle port of the path on the host cores (see module doc)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import raster as R
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = WORKLOADS[args.workload]
    sc = _CpuScene(cfg)
    budget_s, t_start = 240.0, time.perf_counter()
    nwarm = min(args.warmup, 1)
    for i in range(nwarm):
        sc.view(i)
    times = []
    for i in range(args.steps):
        times.append(sc.view(i))
        if time.perf_counter() - t_start > budget_s:
            break
    ms = float(np.mean(times)) * 1e3
    cores = int(R.lib().oracle_get_threads())
    out = {"impl": "reference", "metric": "fwd_bwd_ms_per_view", "value": round(ms, 2), "unit": "ms/view",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(times), "warmup": nwarm,
           "ms_per_step": round(ms, 2), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["desc"], "path": "oracle port: prefilter + decode + rasterize, fwd+bwd, one view per step",
                      "anchors": cfg["N"], "mv": cfg["mv"], "activate_level": LEVEL},
           "cpu_baseline": {"value": round(ms, 2), "unit": "ms/view", "cores": cores, "kind": "port",
                            "sample": "one view of the workload per step (the reference's CUDA rasterizer is absent and "
                                      "has no CPU path; its Python decode cannot travel): oracle port on all host threads"},
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
    ap.add_argument("--level", type=int, default=2, choices=[0, 1, 2],
                    help="activate_level of the feature planes: 2 = steady state after iteration 21 000 (default, the headline), "
                         "0 = the first 12 000 iterations (train.py:305-307, SURVEY §8d times both)")
    ap.add_argument("--torch-scaling-reg", action="store_true",
                    help="scaling regulariser with torch ops as train.py:195 writes it (its prod backward syncs with the host every view)")
    args = ap.parse_args()
    global LEVEL
    LEVEL = args.level
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
