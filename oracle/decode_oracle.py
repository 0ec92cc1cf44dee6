"""CPU restatement of SplatCo's anchor decode (`generate_neural_gaussians`) in plain torch ops.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by splatco_b200/.

PARITY PINNED: tests/test_oracle_decode.py checks this file against the fixtures
tests/golden/decode_*.npz, which were produced by running the reference's own unmodified Python
(`gaussian_renderer.generate_neural_gaussians`, `FeaturePlanes`, `PlaneGrid`, `TriPlaneAttention`)
on CPU with tests/golden/make_decode_golden.py.

Follows (file:line in /root/reference):
  gaussian_renderer/__init__.py:18-116   generate_neural_gaussians
  scene/gaussian_model.py:149-169        FeaturePlanes.forward (levels, models, CTX_models, sum)
  scene/gaussian_model.py:209-215        GaussianLearner.inference (xyz detached, Q = self.Q0)
  scene/grids.py:146-201                 PlaneGrid.forward / compute_planes_feat (bilinear, align_corners)
  scene/grids.py:22-64                   ChannelAttention / SpatialAttention / TriPlaneAttention
  scene/gaussian_model.py:307-337        MLP heads
  scene/embedding.py:53-80               appearance embedding

`params` uses the key names of the fixtures ("feat.k0s.0.xy_plane", "mlp_opacity.0.weight", ...).
BatchNorm is always in train mode (batch statistics over the visible anchors, biased variance,
eps 1e-5), as in the reference where feat_planes never enters eval mode.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def bilinear_plane(plane, u, v):
    """plane [C,A,B]; u indexes A, v indexes B, both in normalised [-1,1] coords (align_corners=True,
    zero padding) — what F.grid_sample(plane[None], grid=(v,u)) computes.  Returns [V,C]."""
    C, A, B = plane.shape
    fu = (u + 1) * 0.5 * (A - 1)
    fv = (v + 1) * 0.5 * (B - 1)
    u0 = torch.floor(fu)
    v0 = torch.floor(fv)
    wu1 = fu - u0
    wv1 = fv - v0
    out = 0
    for du, wu in ((0, 1 - wu1), (1, wu1)):
        for dv, wv in ((0, 1 - wv1), (1, wv1)):
            iu = (u0 + du).long()
            iv = (v0 + dv).long()
            ok = (iu >= 0) & (iu < A) & (iv >= 0) & (iv < B)
            val = plane[:, iu.clamp(0, A - 1), iv.clamp(0, B - 1)]          # [C,V]
            out = out + (val * (wu * wv * ok.to(plane.dtype))[None]).t()
    return out


def triplane_attention(x, p, prefix):
    """x [1,C,E,E] -> TA(x) (scene/grids.py:22-64)."""
    w1 = p[prefix + ".ca.sharedMLP.0.weight"]
    w2 = p[prefix + ".ca.sharedMLP.2.weight"]
    avg = x.mean(dim=(2, 3), keepdim=True)
    mx = x.amax(dim=(2, 3), keepdim=True)
    ca = torch.sigmoid(F.conv2d(F.relu(F.conv2d(avg, w1)), w2) + F.conv2d(F.relu(F.conv2d(mx, w1)), w2))
    x = ca * x
    sa_in = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)
    sa = torch.sigmoid(F.conv2d(sa_in, p[prefix + ".sa.conv.weight"], padding=3))
    return sa * x


def plane_features(p, level_idx, xyz, attended=None):
    """PlaneGrid.forward for k0s[level_idx] with Q = 0 (scene/grids.py:146-201)."""
    pre = f"feat.k0s.{level_idx}"
    mn, mx = p[pre + ".xyz_min"].to(xyz.dtype), p[pre + ".xyz_max"].to(xyz.dtype)
    ind = (xyz - mn) / (mx - mn) * 2 - 1
    X, Y, Z = ind[:, 0], ind[:, 1], ind[:, 2]
    xy, xz, yz = p[pre + ".xy_plane"][0], p[pre + ".xz_plane"][0], p[pre + ".yz_plane"][0]
    f_xy = bilinear_plane(xy, X, Y)      # grid (x=ind[1], y=ind[0]) -> rows follow X, cols follow Y
    f_xz = bilinear_plane(xz, X, Z)
    f_yz = bilinear_plane(yz, Y, Z)
    if level_idx != 0:
        return torch.cat([f_xy, f_xz, f_yz], dim=1)
    if attended is None:
        ta = triplane_attention(torch.cat([xy, xz, yz], dim=0)[None], p, pre + ".TA")[0]
        attended = torch.chunk(ta, 3, dim=0)
    a_xy, a_xz, a_yz = attended
    return torch.cat([f_xy, bilinear_plane(a_xy, X, Y), f_xz, bilinear_plane(a_xz, X, Z),
                      f_yz, bilinear_plane(a_yz, Y, Z)], dim=1)


def batchnorm_train(x, gamma, beta):
    mean = x.mean(dim=0)
    var = x.var(dim=0, unbiased=False)
    return (x - mean) / torch.sqrt(var + BN_EPS) * gamma + beta


def geo_features(p, xyz, g_fea, level):
    """FeaturePlanes.forward: sum over levels of [Linear(BN(planes)) | Linear(BN(g_fea))]."""
    total = 0
    for l in range(level + 1):
        feat = plane_features(p, l, xyz)
        rr = F.linear(batchnorm_train(feat, p[f"feat.models.{l}.0.weight"], p[f"feat.models.{l}.0.bias"]),
                      p[f"feat.models.{l}.1.weight"], p[f"feat.models.{l}.1.bias"])
        rrr = F.linear(batchnorm_train(g_fea, p[f"feat.CTX_models.{l}.0.weight"], p[f"feat.CTX_models.{l}.0.bias"]),
                       p[f"feat.CTX_models.{l}.1.weight"], p[f"feat.CTX_models.{l}.1.bias"])
        total = total + torch.cat([rr, rrr], dim=1)
    return total


def mlp(p, name, x, act):
    h = F.relu(F.linear(x, p[name + ".0.weight"], p[name + ".0.bias"]))
    y = F.linear(h, p[name + ".2.weight"], p[name + ".2.bias"])
    return act(y) if act is not None else y


def decode(p, anchor_feat, anchor, offset, scaling, vis_mask, cam_center, level, n_offsets,
           appearance_dim=0, uid=0, dists=False):
    """Returns (xyz, color, opacity, scaling, rot, neural_opacity, mask) like the reference's training
    branch.  `scaling` is the activated [N,6] tensor (get_scaling = exp(_scaling))."""
    K = n_offsets
    feat = anchor_feat[vis_mask]
    anc = anchor[vis_mask]
    offs = offset[vis_mask]
    scl = scaling[vis_mask]
    V = anc.shape[0]
    g_fea = torch.cat([feat, anc, offs.reshape(V, -1), scl], dim=1)
    geo = geo_features(p, anc.detach(), g_fea, level)
    ob_view = anc - cam_center
    ob_dist = ob_view.norm(dim=1, keepdim=True)
    ob_view = ob_view / ob_dist
    x_wod = torch.cat([feat, ob_view, geo], dim=1)
    x_wd = torch.cat([feat, ob_view, ob_dist, geo], dim=1)
    x = x_wd if dists else x_wod
    neural_opacity = mlp(p, "mlp_opacity", x, torch.tanh).reshape(-1, 1)
    mask = (neural_opacity > 0.0).view(-1)
    opacity = neural_opacity[mask]
    xc = x
    if appearance_dim > 0:
        app = p["embedding_appearance.embedding.weight"][uid][None].expand(V, -1)
        xc = torch.cat([x, app], dim=1)
    color = mlp(p, "mlp_color", xc, torch.sigmoid).reshape(V * K, 3)
    scale_rot = mlp(p, "mlp_cov", x, None).reshape(V * K, 7)
    offsets = offs.reshape(-1, 3)
    rep = torch.cat([scl, anc], dim=1).repeat_interleave(K, dim=0)
    allc = torch.cat([rep, color, scale_rot, offsets], dim=1)[mask]
    scaling_repeat, repeat_anchor, color, scale_rot, offsets = allc.split([6, 3, 3, 7, 3], dim=1)
    out_scaling = scaling_repeat[:, 3:] * torch.sigmoid(scale_rot[:, :3])
    rot = F.normalize(scale_rot[:, 3:7])
    xyz = repeat_anchor + offsets * scaling_repeat[:, :3]
    return xyz, color, opacity, out_scaling, rot, neural_opacity, mask


def bn_running_update(x, running_mean, running_var, momentum=0.1):
    """What nn.BatchNorm1d does to its buffers in train mode (unbiased variance for the buffer)."""
    n = x.shape[0]
    mean = x.mean(dim=0)
    var_unb = x.var(dim=0, unbiased=True) if n > 1 else torch.zeros_like(mean)
    return (1 - momentum) * running_mean + momentum * mean, (1 - momentum) * running_var + momentum * var_unb
