#!/usr/bin/env python
"""Per-tensor gradient errors of the large decode cases (tests/test_decode_gpu.py::test_decode_matches_oracle_large):
ours (impl 1 / 2) and the fp32 torch oracle, both against the SAME oracle evaluated in fp64 on the CPU.
    python tools/debug_decode_large.py --level 0 --rc 4 --N 30000
Diagnostic only."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--rc", type=int, default=4)
    ap.add_argument("--N", type=int, default=30000)
    a = ap.parse_args()
    from oracle import decode_oracle as D
    from splatco_b200 import _lib
    from splatco_b200.gaussian_renderer import generate_neural_gaussians
    from splatco_b200.model import AnchorModel
    from tests.test_decode_gpu import NAMES, Cam
    from tests.util import full_path_grad_errors
    K = 10
    res = {}

    def setup():
        pc = AnchorModel(a.N, n_offsets=K, plane_size=512, num_channels=3 * a.rc, device="cuda", seed=3)
        pc.feat_planes.Q0 = 0.0
        pc.feat_planes._feat.activate_level = a.level
        with torch.no_grad():
            pc._anchor.data[: a.N // 10] *= 2.5
            for m in pc.feat_planes.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.2, 0.2)
        return pc

    g = torch.Generator().manual_seed(9)
    vis = (torch.rand(a.N, generator=g) < 0.7)
    cam = Cam(torch.tensor([2.5, -1.5, 0.7]).cuda(), 0)

    def oracle(pc, dtype):
        p = {"feat." + k: v.detach().cpu() for k, v in pc.feat_planes._feat.state_dict().items()}
        for name in ("mlp_opacity", "mlp_cov", "mlp_color"):
            p.update({f"{name}.{k}": v.detach().cpu() for k, v in getattr(pc, name).state_dict().items()})
        cast = lambda v: v.to(dtype) if v.dtype.is_floating_point else v
        leaves = {k: cast(getattr(pc, k).detach().cpu().clone()).requires_grad_() for k in ("_anchor", "_offset", "_anchor_feat", "_scaling")}
        pw = {k: (cast(v.clone()).requires_grad_() if v.dtype.is_floating_point and "running" not in k and "xyz_m" not in k else cast(v))
              for k, v in p.items()}
        ref = D.decode(pw, leaves["_anchor_feat"], leaves["_anchor"], leaves["_offset"], torch.exp(leaves["_scaling"]), vis,
                       cam.camera_center.cpu().to(dtype), a.level, K)
        gl = torch.Generator().manual_seed(11)
        loss = 0
        for nm, b in zip(NAMES, ref[:5]):
            w = torch.randn(b.shape, generator=gl)
            loss = loss + (b * w.to(dtype)).sum()
        loss.backward()
        grads = {k: v.grad.double().numpy() for k, v in leaves.items()}
        grads.update({k: v.grad.double().numpy() for k, v in pw.items() if getattr(v, "grad", None) is not None})
        return grads, ref[6].numpy()

    pc = setup()
    g64, m64 = oracle(pc, torch.float64)
    g32, m32 = oracle(pc, torch.float32)
    res["oracle32_vs_64"] = {k: full_path_grad_errors(g32[k], g64[k]) for k in g64 if k in g32}
    res["mask32_eq_64"] = bool(np.array_equal(m32, m64))
    for impl in (1, 2):
        _lib.check(_lib.lib().splatco_decode_set_impl(impl), "set_impl")
        pc = setup()
        outs = generate_neural_gaussians(cam, pc, vis.cuda(), is_training=True)
        res[f"mask_eq_impl{impl}"] = bool(np.array_equal(outs[6].cpu().numpy(), m64))
        gl = torch.Generator().manual_seed(11)
        loss = 0
        for nm, t in zip(NAMES, outs[:5]):
            w = torch.randn(t.shape, generator=gl)
            loss = loss + (t * w.cuda()).sum()
        loss.backward()
        got = {k: getattr(pc, k).grad.double().cpu().numpy() for k in ("_anchor", "_offset", "_anchor_feat", "_scaling")}
        for k, v in pc.feat_planes._feat.named_parameters():
            if v.grad is not None:
                got["feat." + k] = v.grad.double().cpu().numpy()
        for name in ("mlp_opacity", "mlp_cov", "mlp_color"):
            for k, v in getattr(pc, name).named_parameters():
                got[f"{name}.{k}"] = v.grad.double().cpu().numpy()
        res[f"impl{impl}_vs_64"] = {k: full_path_grad_errors(got[k], g64[k]) for k in g64 if k in got}
        res[f"impl{impl}_vs_32"] = {k: full_path_grad_errors(got[k], g32[k]) for k in g32 if k in got}
    rows = sorted(res["impl2_vs_64"], key=lambda k: -res["impl2_vs_64"][k]["l2"])
    print(json.dumps({k: v for k, v in res.items() if not isinstance(v, dict)}))
    print(f"{'tensor':45s} {'impl2/64 l2':>12s} {'impl1/64 l2':>12s} {'orc32/64 l2':>12s} {'impl2/32 l2':>12s} {'impl2/64 amax':>13s} {'orc32/64 amax':>13s}")
    for k in rows[:25]:
        print(f"{k:45s} {res['impl2_vs_64'][k]['l2']:12.3e} {res['impl1_vs_64'][k]['l2']:12.3e} {res['oracle32_vs_64'][k]['l2']:12.3e} "
              f"{res['impl2_vs_32'][k]['l2']:12.3e} {res['impl2_vs_64'][k]['amax']:13.3e} {res['oracle32_vs_64'][k]['amax']:13.3e}")


if __name__ == "__main__":
    main()
