"""Small end-to-end target for compute-sanitizer (racecheck / memcheck on small configs): a few iterations
of a train.py-shaped loop on a tiny scene -- prefilter, decode (tcgen05 MLP kernels, gather / scatter), rasterizer
(tile-segmented binning with its shared-memory radix sort and atomics, blend forward / backward with the per-warp
transposition buffer), losses, training_statis, one adjust_anchor pass (grow_* hash set) and FusedAdam.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py      (see tools/sanitizer_run.sh)"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(iters=3):
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.loss import l1_ssim_loss, multiview_consistency_loss, scaling_reg
    from splatco_b200.model import AnchorModel
    from splatco_b200.regularizer import tv_loss
    from splatco_b200.synthetic import ring_cameras
    torch.manual_seed(0)
    K, W, H, mv = 10, 96, 64, 2
    pc = AnchorModel(1500, n_offsets=K, plane_size=64, num_channels=15, device="cuda", seed=2, scale_factor=1.0)
    pc.feat_planes._feat.activate_level = 2
    pc.feat_planes.Q0 = 0.03
    pc.train()
    opt = pc.training_setup(voxel_size=0.02)
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    cams = [c.to("cuda") for c in ring_cameras(4, W, H)]
    g = torch.Generator().manual_seed(5)
    base = torch.rand(3, H, W, generator=g)
    gts = [(base + 0.05 * torch.randn(3, H, W, generator=g)).clamp(0, 1).cuda() for _ in cams]
    bg = torch.ones(3, device="cuda")
    for it in range(1, iters + 1):
        total, gens, reals = None, [], []
        for v in range(mv):
            i = (it * mv + v) % len(cams)
            vm = prefilter_voxel(cams[i], pc, pipe, bg)
            pkg = render(cams[i], pc, pipe, bg, visible_mask=vm, retain_grad=True)
            loss = l1_ssim_loss(pkg["render"], gts[i], 0.2) + 0.01 * scaling_reg(pkg["scaling"])
            total = loss if total is None else total + loss
            gens.append(pkg["render"]); reals.append(gts[i])
        total = total + 0.05 * multiview_consistency_loss(gens, reals, 0.6)
        total.backward()
        tv_loss(pc.feat_planes, 4e-7)
        with torch.no_grad():
            pc.training_statis(pkg["viewspace_points"], pkg["neural_opacity"], pkg["visibility_filter"], pkg["selection_mask"], vm)
            if it == iters:
                pc.adjust_anchor(iteration=it, check_interval=1, success_threshold=0.0, grad_threshold=1e-9, min_opacity=0.005)
        opt.step()
        opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    print("sanitize target done: anchors", int(pc.get_anchor.shape[0]), "loss", float(total))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
