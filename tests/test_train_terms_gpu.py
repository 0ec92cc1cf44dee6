"""GPU parity of the training-loop terms around the render path, through the C ABI:
  * against tests/golden/train_terms.npz (outputs of the reference's own Python, make_train_terms_golden.py)
  * against oracle/train_terms_oracle.py (pinned to the same fixtures on CPU) at larger sizes and on edge cases.
Bars: anchor growing bit-exact (integer voxel work; max is order independent); TV gradient 1e-6 of max|g|;
cross-view loss value 2e-6, gradient 1e-3 relative."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_terms.npz")
PLANES = ("xy_plane", "xz_plane", "yz_plane")


# ---- total variation -------------------------------------------------------------------------------------------------
def _grid(d, n, grad_key="grad0"):
    g = SimpleNamespace()
    for name in PLANES:
        p = nn.Parameter(torch.from_numpy(d[f"tv.g{n}.{name}"]).cuda())
        p.grad = torch.from_numpy(d[f"tv.g{n}.{name}.{grad_key}"].copy()).cuda()
        setattr(g, name, p)
    return g


def test_tv_matches_reference():
    from splatco_b200.regularizer import total_variation_add_grad, tv_loss
    d = np.load(GOLD)
    g1 = _grid(d, 1)
    total_variation_add_grad(g1, float(d["tv.w_direct"]))
    for name in PLANES:
        want = d[f"tv.g1.{name}.grad_direct"]
        assert np.abs(getattr(g1, name).grad.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()
    grids = [_grid(d, n) for n in range(3)]
    tv_loss(SimpleNamespace(_feat=SimpleNamespace(activate_level=2, k0s=grids)), float(d["tv.w_tvloss"]))
    for n in range(3):
        for name in PLANES:
            want = d[f"tv.g{n}.{name}.grad_tvloss"]
            assert np.abs(getattr(grids[n], name).grad.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max(), (n, name)


def test_tv_large_plane_and_fresh_grad():
    from oracle import train_terms_oracle as O
    from splatco_b200.regularizer import total_variation_add_grad
    g = torch.Generator().manual_seed(2)
    pg = SimpleNamespace()
    for name, (H, W) in zip(PLANES, [(625, 625), (301, 1250), (1, 37)]):       # odd sizes, a single-row plane
        setattr(pg, name, nn.Parameter((torch.randn(1, 5, H, W, generator=g) * 0.8).cuda()))
    total_variation_add_grad(pg, 0.01)                                           # .grad is None: created like autograd would
    for name in PLANES:
        p = getattr(pg, name)
        want = O.tv_grad(p.detach().cpu().numpy(), 0.01)
        assert np.abs(p.grad.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()
    with pytest.raises(RuntimeError):
        total_variation_add_grad(SimpleNamespace(**{n: nn.Parameter(torch.zeros(1, 5, 4, 4)) for n in PLANES}), 0.1)   # CPU tensors


# ---- cross-view consistency --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [0, 1])
def test_mv_consistency_matches_reference(case):
    from splatco_b200.loss import multiview_consistency_loss, pair_ssim
    d = np.load(GOLD)
    n = int(d[f"mvc.c{case}.n"])
    gens = [torch.from_numpy(d[f"mvc.c{case}.gen{v}"]).cuda().requires_grad_() for v in range(n)]
    reals = [torch.from_numpy(d[f"mvc.c{case}.real{v}"]).cuda() for v in range(n)]
    assert np.abs(pair_ssim(reals).cpu().numpy() - d[f"mvc.c{case}.ssim"]).max() < 2e-6
    loss, parts = multiview_consistency_loss(gens, reals, 0.6, return_parts=True)
    (0.05 * loss).backward()
    assert abs(loss.item() - float(d[f"mvc.c{case}.loss"])) < 2e-6
    assert np.abs(parts[1:].cpu().numpy() - d[f"mvc.c{case}.parts"]).max() < 2e-6
    for v in range(n):
        want = d[f"mvc.c{case}.grad{v}"]
        got = gens[v].grad.cpu().numpy()
        assert got.shape == want.shape
        assert (np.abs(got - want) <= 1e-3 * np.abs(want) + 1e-12).all(), (v, np.abs(got - want).max())


def test_mv_consistency_full_size_properties():
    from oracle import train_terms_oracle as O
    from splatco_b200.loss import multiview_consistency_loss, pair_ssim
    g = torch.Generator(device="cuda").manual_seed(8)
    base = torch.rand(3, 545, 980, device="cuda", generator=g)
    reals = [(base + 0.05 * torch.randn(3, 545, 980, device="cuda", generator=g)).clamp(0, 1) for _ in range(3)] + [torch.rand(3, 545, 980, device="cuda", generator=g)]
    gens = [(r + 0.1 * torch.randn_like(r)).clamp(0, 1).requires_grad_() for r in reals]
    ps = pair_ssim(reals)
    loss = multiview_consistency_loss(gens, reals, 0.6, pair_ssim_values=ps)
    loss.backward()
    g1 = [t.grad.clone() for t in gens]
    want, parts, ssims, grads = O.mv_consistency([t.detach().cpu().numpy() for t in gens], [t.cpu().numpy() for t in reals], 0.6)
    assert np.abs(ps.cpu().numpy() - ssims).max() < 5e-6
    assert abs(loss.item() - want) < 1e-6 * max(1.0, abs(want))
    for v in range(4):
        assert (np.abs(g1[v].cpu().numpy() - grads[v]) <= 1e-3 * np.abs(grads[v]) + 1e-12).all()
    assert g1[3].abs().max().item() == 0.0                       # the unrelated view is gated out of every pair
    for t in gens:
        t.grad = None
    (2.5 * multiview_consistency_loss(gens, reals, 0.6, pair_ssim_values=ps)).backward()      # linear in the upstream gradient
    for v in range(4):
        assert torch.allclose(gens[v].grad, 2.5 * g1[v], rtol=1e-6, atol=0)
    # the sum of the gradients over the views of one pair cancels: d/dgen_i = -d/dgen_j
    assert torch.allclose(g1[0] + g1[1] + g1[2], torch.zeros_like(g1[0]), atol=1e-12)


# ---- anchor growing ----------------------------------------------------------------------------------------------------
class FakeModel:
    """What anchor_growing / adjust_anchor touch on a GaussianModel: parameters, statistics, and the two optimizer
    bookkeeping methods (cat rows / mask rows; _prune_anchor_optimizer also clamps scaling[:, 3:] at 0.05,
    scene/gaussian_model.py:796-815)."""
    NAMES = {"anchor": "_anchor", "offset": "_offset", "anchor_feat": "_anchor_feat", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}

    def __init__(self, d, p, voxel_size):
        for k in self.NAMES.values():
            setattr(self, k, torch.from_numpy(d[f"{p}.in.{k}"]).cuda())
        for k in ("opacity_accum", "anchor_demon", "offset_denom", "offset_gradient_accum"):
            setattr(self, k, torch.from_numpy(d[f"{p}.in.{k}"]).cuda())
        self.n_offsets, self.feat_dim, self.voxel_size = 10, 32, voxel_size
        self.update_depth, self.update_init_factor, self.update_hierachy_factor = 3, 16, 4

    get_anchor = property(lambda self: self._anchor)
    get_scaling = property(lambda self: 1.0 * torch.exp(self._scaling))

    def cat_tensors_to_optimizer(self, td):
        return {n: torch.cat([getattr(self, a), td[n]], dim=0) for n, a in self.NAMES.items()}

    def prune_anchor(self, mask):
        keep = ~mask
        for a in self.NAMES.values():
            setattr(self, a, getattr(self, a)[keep])
        self._scaling[:, 3:] = self._scaling[:, 3:].clamp(max=0.05)


class _ReplayRand:
    def __init__(self, rands):
        self.rands, self.saved = list(rands), torch.rand_like

    def __enter__(self):
        torch.rand_like = lambda t, *a, **k: self.rands.pop(0).to(t.device)
        return self

    def __exit__(self, *exc):
        torch.rand_like = self.saved


def _check_model(pc, d, p, keys):
    for k in keys:
        got = getattr(pc, k).detach().cpu().numpy()
        if k == "_offset":
            assert np.array_equal(got.astype(np.float64).sum(axis=(1, 2)), d[f"{p}._offset.rowsum"])
        else:
            want = d[f"{p}.{k}"]
            assert got.shape == want.shape, (k, got.shape, want.shape)
            if k in ("_scaling", "_opacity"):          # rows of new anchors hold log() values: device log vs the fixture's CPU log
                assert np.allclose(got, want, rtol=1e-6, atol=0), k
            else:
                assert np.array_equal(got, want), k


@pytest.mark.parametrize("case", [0, 1])
def test_anchor_growing_matches_reference(case):
    from splatco_b200 import densify
    d = np.load(GOLD)
    p = f"grow.c{case}"
    rands = [torch.from_numpy(d[f"{p}.rand{i}"]) for i in range(int(d[f"{p}.n_rand"]))]
    pc = FakeModel(d, p, float(d[f"{p}.voxel_size"]))
    with _ReplayRand(rands):
        densify.anchor_growing(pc, torch.from_numpy(d[f"{p}.grads_norm"]).cuda(), 0.0002, torch.from_numpy(d[f"{p}.offset_mask"]).cuda())
    _check_model(pc, d, f"{p}.grown", ["_anchor", "_anchor_feat", "_scaling", "_rotation", "_opacity", "_offset", "opacity_accum", "anchor_demon"])

    pc = FakeModel(d, p, float(d[f"{p}.voxel_size"]))
    with _ReplayRand(rands):
        densify.adjust_anchor(pc, iteration=1700, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005)
    _check_model(pc, d, f"{p}.adjusted", ["_anchor", "_anchor_feat", "_scaling", "_rotation", "_opacity", "_offset", "opacity_accum", "anchor_demon",
                                          "offset_denom", "offset_gradient_accum"])


@pytest.mark.parametrize("div_mode", [0, 1])
def test_grow_pass_large_vs_oracle(div_mode):
    """200 k anchors, a voxel size that is not a power of two (reciprocal and true division differ on ties), negative
    coordinates, anchors sitting exactly on voxel centres; both forms of the candidate selection."""
    from oracle import train_terms_oracle as O
    from splatco_b200.densify import grow_pass
    g = torch.Generator().manual_seed(77 + div_mode)
    N, K, F = 200_000, 10, 32
    cur = 0.037
    anchor = torch.rand(N, 3, generator=g) * 6 - 3
    anchor[: N // 3] = torch.round(anchor[: N // 3] / cur) * cur
    anchor[N // 3: N // 2] = (torch.round(anchor[N // 3: N // 2] / cur) + 0.5) * cur          # exact ties of round()
    offset = torch.randn(N, K, 3, generator=g) * 0.5
    scaling = 0.06 * torch.exp(torch.randn(N, 6, generator=g) * 0.3)
    feat = torch.randn(N, F, generator=g)
    feat[::7] = 0.0
    feat[::11] = -feat[::11].abs()
    grads = torch.rand(N * K, generator=g) * 1e-3
    omask = torch.rand(N * K, generator=g) < 0.7
    rand = torch.rand(N * K, generator=g)
    cand = (grads >= 0.0004) & omask & (rand > 0.5)
    want_a, want_f = O.grow_pass(anchor.numpy(), offset.numpy(), scaling.numpy(), feat.numpy(), cand.numpy(), cur, div_mode)
    assert want_a.shape[0] > 10_000
    dev = lambda t: t.cuda()
    a1, f1, n1 = grow_pass(dev(anchor), dev(offset), dev(scaling), dev(feat), cur, candidate_mask=dev(cand), div_mode=div_mode)
    assert n1 == int(cand.sum())
    assert np.array_equal(a1.cpu().numpy(), want_a) and np.array_equal(f1.cpu().numpy(), want_f)
    a2, f2, n2 = grow_pass(dev(anchor), dev(offset), dev(scaling)[:, :3], dev(feat), cur, grads=dev(grads), threshold=0.0004,
                           offset_mask=dev(omask), rand=dev(rand), rand_cut=0.5, div_mode=div_mode)     # strided scaling slice
    assert n2 == n1 and torch.equal(a2, a1) and torch.equal(f2, f1)


def test_grow_pass_edge_cases():
    from splatco_b200.densify import grow_pass
    N, K, F = 64, 10, 32
    anchor = torch.zeros(N, 3).cuda()
    offset = torch.zeros(N, K, 3).cuda()
    scaling = torch.ones(N, 6).cuda()
    feat = torch.arange(N * F, dtype=torch.float32).reshape(N, F).cuda()
    none = torch.zeros(N * K, dtype=torch.bool).cuda()
    a, f, n = grow_pass(anchor, offset, scaling, feat, 0.5, candidate_mask=none)
    assert n == 0 and a.shape == (0, 3) and f.shape == (0, F)
    allc = torch.ones(N * K, dtype=torch.bool).cuda()
    a, f, n = grow_pass(anchor, offset, scaling, feat, 0.5, candidate_mask=allc)       # every candidate sits in an occupied voxel
    assert n == N * K and a.shape == (0, 3)
    offset[:, 3] = torch.tensor([-1.0, 2.0, 0.26]).cuda()                               # one new voxel shared by all anchors
    a, f, n = grow_pass(anchor, offset, scaling, feat, 0.5, candidate_mask=allc)
    assert a.cpu().tolist() == [[-1.0, 2.0, 0.5]] and torch.equal(f[0], feat[N - 1])
    with pytest.raises(RuntimeError):
        grow_pass(anchor.cpu(), offset.cpu(), scaling.cpu(), feat.cpu(), 0.5, candidate_mask=allc.cpu())


# ---- CVPM pruning mask -----------------------------------------------------------------------------------------------
def test_cvpm_matches_reference():
    from splatco_b200.cvpm import compute_fast_loss_with_key_points
    d = np.load(GOLD)
    c = lambda k: torch.from_numpy(d[k]).cuda()
    cloud, eye = c("cvpm.cloud"), torch.eye(3)
    for name, r2 in (("open", "cvpm.real2"), ("tight", "cvpm.real2"), ("gated", "cvpm.other")):
        a, b, pts, mask = compute_fast_loss_with_key_points(None, c("cvpm.real1"), c(r2), c("cvpm.gen1"), c("cvpm.gen2"), eye, eye,
                                                            torch.from_numpy(d["cvpm.t1"]), eye, eye, torch.from_numpy(d["cvpm.t2"]), cloud,
                                                            distance_threshold=float(d[f"cvpm.{name}.thr"]), overall_ssim_threshold=0.6)
        want = np.unpackbits(d[f"cvpm.{name}.mask"])[: cloud.shape[0]].astype(bool)
        assert int(mask.sum()) == int(want.sum()) == int(pts.shape[0])               # the count train.py prunes by
        assert np.array_equal(mask.cpu().numpy(), want), name
        assert abs(float(a) - float(d[f"cvpm.{name}.gen_l1"])) < 2e-6 and abs(float(b) - float(d[f"cvpm.{name}.cross_l1"])) < 2e-6


def test_cvpm_large_vs_oracle_and_edges():
    from oracle import train_terms_oracle as O
    from splatco_b200.cvpm import cvpm_mask
    g = torch.Generator().manual_seed(5)
    N = 1_000_000
    t1, t2 = torch.tensor([0.5, -2.0, 1.0]), torch.tensor([-1.0, 1.5, 0.3])
    d = (t2 - t1) / (t2 - t1).norm()
    cloud = torch.rand(N, 3, generator=g) * 6 - 3
    cloud[:100_000] = t1 + d * (torch.rand(100_000, 1, generator=g) * 20 - 10) + torch.randn(100_000, 3, generator=g) * 0.01
    want = O.cvpm_mask(cloud.numpy(), t1.numpy(), t2.numpy(), None, 0.01)
    mask, count = cvpm_mask(cloud.cuda(), t1, t2, None, 0.01)
    got = mask.cpu().numpy()
    assert int(count.item()) == int(got.sum())
    # fp32 distances within an ulp of a threshold may fall either side (the dot product is an FMA chain here, a BLAS call
    # in torch): allow a handful of flips out of a million, none of them far from a threshold
    assert int((got != want).sum()) <= 3 and int(want.sum()) > 10_000
    one = cloud[:1].cuda()
    m1, c1 = cvpm_mask(one, t1, t2, None, 100.0)           # a single point: std is NaN -> every point is an "outlier", like torch
    assert bool(m1[0]) and int(c1.item()) == 1
    m0, c0 = cvpm_mask(cloud[:0].cuda(), t1, t2, None, 0.01)
    assert m0.shape == (0,) and int(c0.item()) == 0
    mg, cg = cvpm_mask(cloud[:1000].cuda(), t1, t2, torch.tensor(0.3).cuda(), 100.0)
    assert int(cg.item()) == 0 and not bool(mg.any())
