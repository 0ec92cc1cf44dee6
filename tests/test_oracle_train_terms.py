"""CPU: the numpy restatements in oracle/train_terms_oracle.py against tests/golden/train_terms.npz, which holds outputs
of the reference's own Python (PlaneGrid.total_variation_add_grad / tv_loss, loss_utils + the train.py pair loop,
GaussianModel.anchor_growing).  This pins the oracle the GPU tests then use at larger sizes."""
import os

import numpy as np

from oracle import train_terms_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_terms.npz")
PLANES = ("xy_plane", "xz_plane", "yz_plane")


def test_tv_oracle_matches_reference():
    d = np.load(GOLD)
    for name in PLANES:
        want = d[f"tv.g1.{name}.grad_direct"]
        got = d[f"tv.g1.{name}.grad0"] + O.tv_grad(d[f"tv.g1.{name}"], float(d["tv.w_direct"]))
        assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    w = float(d["tv.w_tvloss"])
    for n in range(3):                                   # tv_loss: level weights w * 0.5 ** (2 - level)
        for name in PLANES:
            want = d[f"tv.g{n}.{name}.grad_tvloss"]
            got = d[f"tv.g{n}.{name}.grad0"] + O.tv_grad(d[f"tv.g{n}.{name}"], w * 0.5 ** (2 - n))
            assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), (n, name)
    assert np.abs(O.tv_grad(d["tv.g1.xy_plane"], 1.0)).max() > 0.3      # the clamped branch is exercised (|diff| > 1)


def test_mv_consistency_oracle_matches_reference():
    d = np.load(GOLD)
    for case in (0, 1):
        n = int(d[f"mvc.c{case}.n"])
        gens = [d[f"mvc.c{case}.gen{v}"] for v in range(n)]
        reals = [d[f"mvc.c{case}.real{v}"] for v in range(n)]
        loss, parts, ssims, grads = O.mv_consistency(gens, reals, 0.6)
        assert abs(loss - float(d[f"mvc.c{case}.loss"])) < 2e-6
        assert np.abs(parts - d[f"mvc.c{case}.parts"]).max() < 2e-6
        assert np.abs(ssims - d[f"mvc.c{case}.ssim"]).max() < 2e-6
        assert (parts > 0).any() and (parts == 0).any()                  # both sides of the SSIM gate
        for v in range(n):
            want = d[f"mvc.c{case}.grad{v}"]                             # gradient of 0.05 * loss
            assert np.abs(0.05 * grads[v] - want).max() <= 1e-3 * np.abs(want).max() + 1e-12


def test_anchor_growing_oracle_matches_reference():
    d = np.load(GOLD)
    for case in (0, 1):
        p = f"grow.c{case}"
        rands = [d[f"{p}.rand{i}"] for i in range(int(d[f"{p}.n_rand"]))]
        anchor, offset, ls, feat, added = O.anchor_growing(
            d[f"{p}.in._anchor"], d[f"{p}.in._offset"], d[f"{p}.in._scaling"], d[f"{p}.in._anchor_feat"], d[f"{p}.grads_norm"], 0.0002,
            d[f"{p}.offset_mask"], rands, float(d[f"{p}.voxel_size"]), div_mode=1)
        assert sum(added) > 0 and all(a > 0 for a in added)
        assert np.array_equal(anchor, d[f"{p}.grown._anchor"])
        assert np.array_equal(feat, d[f"{p}.grown._anchor_feat"])
        assert np.array_equal(ls, d[f"{p}.grown._scaling"])
        assert np.array_equal(offset.astype(np.float64).sum(axis=(1, 2)), d[f"{p}.grown._offset.rowsum"])
        # power-of-two voxel sizes: the reciprocal form (torch CUDA) gives the same grid
        a2, _, _, f2, _ = O.anchor_growing(
            d[f"{p}.in._anchor"], d[f"{p}.in._offset"], d[f"{p}.in._scaling"], d[f"{p}.in._anchor_feat"], d[f"{p}.grads_norm"], 0.0002,
            d[f"{p}.offset_mask"], rands, float(d[f"{p}.voxel_size"]), div_mode=0)
        assert np.array_equal(a2, anchor) and np.array_equal(f2, feat)


def test_cvpm_mask_oracle_matches_reference():
    d = np.load(GOLD)
    cloud, t1, t2 = d["cvpm.cloud"], d["cvpm.t1"], d["cvpm.t2"]
    s_open = O.ssim(d["cvpm.real1"][:, :38, :50], d["cvpm.real2"])
    s_gated = O.ssim(d["cvpm.real1"][:, :38, :50], d["cvpm.other"])
    assert s_open > 0.6 > s_gated
    for name, s in (("open", s_open), ("tight", s_open), ("gated", s_gated)):
        want = np.unpackbits(d[f"cvpm.{name}.mask"])[: cloud.shape[0]].astype(bool)
        got = O.cvpm_mask(cloud, t1, t2, s, float(d[f"cvpm.{name}.thr"]))
        assert int(got.sum()) == int(want.sum()) == int(d[f"cvpm.{name}.n_points"])
        assert np.array_equal(got, want), name
    assert int(np.unpackbits(d["cvpm.open.mask"]).sum()) > 500 and int(np.unpackbits(d["cvpm.gated.mask"]).sum()) == 0
