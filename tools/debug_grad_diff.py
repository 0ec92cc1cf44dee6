"""Debug aid: where do full-path gradients of the product path and a torch-autograd decode differ?
Both sides use the SAME product rasterizer; side A decodes with oracle/decode_oracle.py on cuda (torch autograd, fp32 or
fp64), side B with the product decode.  Prints, per leaf, max|b|, the error quantiles and the worst element."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def main(N=20000, W=320, H=200, level=2, C=15, loss_kind="l1", dtype=torch.float32):
    from oracle import decode_oracle as D
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizer
    from splatco_b200.gaussian_renderer import _settings, generate_neural_gaussians, prefilter_voxel
    from splatco_b200.model import AnchorModel
    from splatco_b200.synthetic import ring_cameras
    dev = "cuda"
    pc = AnchorModel(N, n_offsets=10, plane_size=512, num_channels=C, device=dev, seed=3, scale_factor=0.5)
    pc.feat_planes._feat.activate_level = level
    pc.feat_planes.Q0 = 0.0
    pc.train()
    cam = ring_cameras(3, W, H)[1].to(dev)
    bg = torch.ones(3, device=dev)
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    wimg = torch.randn(3, H, W, generator=torch.Generator().manual_seed(6)).to(dev) / (3 * H * W)
    settings = _settings(cam, pipe, bg, 1.0)
    with torch.no_grad():
        vm = prefilter_voxel(cam, pc, pipe, bg)

    def raster_loss(xyz, color, opacity, scl, rot):
        m2d = torch.zeros_like(xyz, requires_grad=True)
        img, radii = GaussianRasterizer(settings)(means3D=xyz, means2D=m2d, shs=None, colors_precomp=color, opacities=opacity,
                                                  scales=scl, rotations=rot, cov3D_precomp=None)
        if loss_kind == "l1":
            return (img - gt).abs().mean(), img
        return (img * wimg).sum(), img

    # side B: product decode
    names = ["_anchor", "_offset", "_anchor_feat", "_scaling"]
    for p in pc.parameters():
        p.grad = None
    outs = generate_neural_gaussians(cam, pc, vm, is_training=True)
    lb, img_b = raster_loss(*outs[:5])
    lb.backward()
    gb = {k: getattr(pc, k).grad.detach().double().cpu().numpy() for k in names}
    for name in ("mlp_opacity", "mlp_cov", "mlp_color"):
        for k, v in getattr(pc, name).named_parameters():
            gb[f"{name}.{k}"] = v.grad.detach().double().cpu().numpy()
    for k, v in pc.feat_planes._feat.named_parameters():
        if v.grad is not None:
            gb[f"feat.{k}"] = v.grad.detach().double().cpu().numpy()
    # side A: torch decode on cuda
    p = {"feat." + k: v.detach() for k, v in pc.feat_planes._feat.state_dict().items()}
    for name in ("mlp_opacity", "mlp_cov", "mlp_color"):
        p.update({f"{name}.{k}": v.detach() for k, v in getattr(pc, name).state_dict().items()})
    pw = {k: (v.to(dtype).clone().requires_grad_() if v.dtype.is_floating_point and "running" not in k and "xyz_m" not in k else v)
          for k, v in p.items()}
    leaves = {k: getattr(pc, k).detach().to(dtype).clone().requires_grad_() for k in names}
    ref = D.decode(pw, leaves["_anchor_feat"], leaves["_anchor"], leaves["_offset"], torch.exp(leaves["_scaling"]), vm,
                   cam.camera_center.to(dtype), level, 10)
    mask_equal = bool(torch.equal(ref[6], outs[6]))
    ra = [t.float() for t in ref[:5]]
    la, img_a = raster_loss(*ra)
    la.backward()
    ga = {k: leaves[k].grad.detach().double().cpu().numpy() for k in names}
    for k, v in pw.items():
        if torch.is_tensor(v) and v.requires_grad and v.grad is not None:
            ga[k] = v.grad.detach().double().cpu().numpy()
    print(f"== N={N} {W}x{H} level={level} C={C} loss={loss_kind} dtype={dtype} V={int(vm.sum())} M={outs[0].shape[0]} mask_equal={mask_equal} "
          f"img_diff={float((img_a - img_b).abs().max()):.2e}")
    rows = []
    for k in gb:
        if k not in ga:
            continue
        a, b = ga[k].ravel(), gb[k].ravel()          # a: torch reference, b: product
        mx = np.abs(a).max()
        if mx == 0:
            continue
        e = np.abs(a - b) / np.maximum(np.abs(a), 3e-3 * mx)
        i = int(e.argmax())
        rows.append((float(e.max()), k, mx, float(np.quantile(e, 0.999)), a[i], b[i], i, a.size))
    for r in sorted(rows, reverse=True)[:10]:
        print("  max_rel %.4f  %-28s max|g| %.3e  q99.9 %.2e  worst: ref %.6e ours %.6e idx %d of %d" % r)


if __name__ == "__main__":
    main(loss_kind="l1")
    main(loss_kind="dense")
    main(loss_kind="l1", dtype=torch.float64)
    main(N=30000, level=0, C=12, loss_kind="dense")
