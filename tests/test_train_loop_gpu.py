"""End-to-end: a train.py-shaped loop (train.py:165-312) on a small synthetic scene with EVERY drop-in in place —
prefilter_voxel / render, fused L1+SSIM + scaling regulariser, the cross-view consistency term, one backward over the
mv views, tv_loss every 4th iteration, training_statis on the last view, adjust_anchor (anchor growing + pruning, which
swaps the per-anchor Parameters and resizes the Adam moments) and FusedAdam.  Checks the bookkeeping invariants after
densification and that optimisation keeps working on the resized model."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_training_loop_with_densification():
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.loss import l1_ssim_loss, multiview_consistency_loss, scaling_reg
    from splatco_b200.model import AnchorModel
    from splatco_b200.regularizer import tv_loss
    from splatco_b200.synthetic import ring_cameras
    torch.manual_seed(0)
    K, W, H, mv = 10, 160, 112, 2
    pc = AnchorModel(4000, n_offsets=K, plane_size=128, num_channels=15, device="cuda", seed=2, scale_factor=1.0)
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
    losses, sizes = [], []
    for it in range(1, 41):
        total, gens, reals = None, [], []
        for v in range(mv):
            i = (it * mv + v) % len(cams)
            vm = prefilter_voxel(cams[i], pc, pipe, bg)
            pkg = render(cams[i], pc, pipe, bg, visible_mask=vm, retain_grad=True)
            loss = l1_ssim_loss(pkg["render"], gts[i], 0.2) + 0.01 * scaling_reg(pkg["scaling"])
            total = loss if total is None else total + loss
            gens.append(pkg["render"])
            reals.append(gts[i])
        total = total + 0.05 * multiview_consistency_loss(gens, reals, 0.6)
        total.backward()
        if it % 4 == 0:
            tv_loss(pc.feat_planes, 4e-7)
        with torch.no_grad():
            pc.training_statis(pkg["viewspace_points"], pkg["neural_opacity"], pkg["visibility_filter"], pkg["selection_mask"], vm)
            if it in (20, 30):
                before = int(pc.get_anchor.shape[0])
                pc.adjust_anchor(iteration=it, check_interval=10, success_threshold=0.8, grad_threshold=1e-7, min_opacity=0.005)
                N = int(pc.get_anchor.shape[0])
                sizes.append((before, N))
                # every per-anchor tensor, statistic and Adam moment follows the new anchor count
                for t in (pc._offset, pc._anchor_feat, pc._scaling, pc._rotation, pc._opacity, pc.opacity_accum, pc.anchor_demon):
                    assert t.shape[0] == N
                assert pc.offset_denom.shape == (N * K, 1) and pc.offset_gradient_accum.shape == (N * K, 1)
                for name in ("_anchor", "_offset", "_anchor_feat", "_scaling"):
                    p = getattr(pc, name)
                    st = opt.state.get(p)
                    assert st is not None and st["exp_avg"].shape == p.shape and st["exp_avg_sq"].shape == p.shape, name
                    assert any(gr["params"][0] is p for gr in opt.param_groups), name
                assert torch.isfinite(pc._anchor).all() and torch.isfinite(pc._anchor_feat).all()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(float(total.item()))
    assert all(l == l and l < 1e3 for l in losses), losses
    assert sizes[0][1] != sizes[0][0], f"anchor count did not change at the first adjust_anchor: {sizes}"
    assert max(n for _, n in sizes) > 4000, f"no anchors were grown: {sizes}"
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5, (losses[:5], losses[-5:])
