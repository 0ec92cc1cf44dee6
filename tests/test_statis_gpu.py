"""training_statis kernel against tests/golden/statis.npz, produced by the reference's own
GaussianModel.training_statis (tests/golden/make_statis_golden.py).  Counters bit-exact, float sums to 1e-6."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "statis.npz")


def test_training_statis_matches_reference():
    from splatco_b200.statis import training_statis
    d = np.load(GOLD)
    N, K = int(d["N"]), int(d["K"])
    dev = "cuda"
    pc = SimpleNamespace(n_offsets=K, opacity_accum=torch.zeros(N, 1, device=dev), anchor_demon=torch.zeros(N, 1, device=dev),
                         offset_gradient_accum=torch.zeros(N * K, 1, device=dev), offset_denom=torch.zeros(N * K, 1, device=dev))
    for call in range(2):
        t = lambda k: torch.from_numpy(d[f"c{call}.{k}"]).to(dev)
        vp = SimpleNamespace(grad=t("grad"))
        training_statis(pc, vp, t("nopac"), t("upd"), t("sel"), t("vis"))
        assert np.array_equal(pc.anchor_demon.cpu().numpy(), d[f"c{call}.anchor_demon"]), "anchor_demon must be bit-exact"
        assert np.array_equal(pc.offset_denom.cpu().numpy(), d[f"c{call}.offset_denom"]), "offset_denom must be bit-exact"
        for k in ("opacity_accum", "offset_gradient_accum"):
            got, want = getattr(pc, k).cpu().numpy(), d[f"c{call}.{k}"]
            assert np.abs(got - want).max() <= 1e-6 * max(np.abs(want).max(), 1.0), k


def test_training_statis_after_render_backward():
    """End to end on render() outputs (the call train.py:264-266 makes), against the reference formulas in torch."""
    from tests.test_render_gpu import PIPE, _cams, _model
    from splatco_b200.gaussian_renderer import prefilter_voxel, render
    from splatco_b200.statis import training_statis
    pc = _model(N=2000, seed=9)
    cam = _cams(160, 112)[1]
    bg = torch.ones(3, device="cuda")
    vm = prefilter_voxel(cam, pc, PIPE, bg)
    pkg = render(cam, pc, PIPE, bg, visible_mask=vm, retain_grad=True)
    pkg["render"].mean().backward()
    N, K = pc._anchor.shape[0], pc.n_offsets
    for name, n in (("opacity_accum", N), ("anchor_demon", N), ("offset_gradient_accum", N * K), ("offset_denom", N * K)):
        setattr(pc, name, torch.zeros(n, 1, device="cuda"))
    training_statis(pc, pkg["viewspace_points"], pkg["neural_opacity"], pkg["visibility_filter"], pkg["selection_mask"], vm)
    # reference formulas (scene/gaussian_model.py:761-782)
    tmp = pkg["neural_opacity"].detach().view(-1).clamp_min(0).view(-1, K)
    oa = torch.zeros(N, 1, device="cuda"); oa[vm] += tmp.sum(dim=1, keepdim=True)
    ad = torch.zeros(N, 1, device="cuda"); ad[vm] += 1
    comb = torch.zeros(N * K, dtype=torch.bool, device="cuda")
    comb[vm.unsqueeze(1).repeat(1, K).view(-1)] = pkg["selection_mask"]
    tmpm = comb.clone(); comb[tmpm] = pkg["visibility_filter"]
    gn = torch.norm(pkg["viewspace_points"].grad[pkg["visibility_filter"], :2], dim=-1, keepdim=True)
    og = torch.zeros(N * K, 1, device="cuda"); og[comb] += gn
    od = torch.zeros(N * K, 1, device="cuda"); od[comb] += 1
    assert torch.equal(pc.anchor_demon, ad) and torch.equal(pc.offset_denom, od)
    assert (pc.opacity_accum - oa).abs().max() <= 1e-6 * max(oa.abs().max().item(), 1.0)
    assert (pc.offset_gradient_accum - og).abs().max() <= 1e-6 * max(og.abs().max().item(), 1e-12)
    assert od.sum() > 0
