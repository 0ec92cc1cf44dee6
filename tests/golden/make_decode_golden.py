"""Generates tests/golden/decode_*.npz by running the REFERENCE's own Python decode on CPU.

Run here (the authoring container), where /root/reference exists:
    python tests/golden/make_decode_golden.py
It imports, unmodified, from /root/reference:
    gaussian_renderer.generate_neural_gaussians        (gaussian_renderer/__init__.py:18-116)
    scene.gaussian_model.FeaturePlanes                 (scene/gaussian_model.py:97-181)
    scene.grids.PlaneGrid / TriPlaneAttention          (scene/grids.py:22-64,102-201)
with empty sys.modules stubs for the native/absent packages the reference imports at module load
(SURVEY.md §8c).  The only patch: FeaturePlanes.get_offsets hard-codes device='cuda'
(scene/gaussian_model.py:178-179) and is replaced by a CPU twin (its outputs feed the never-executed
Spatial_CTX only).  The MLP heads are constructed exactly as GaussianModel.__init__ does
(scene/gaussian_model.py:307-337) minus the `.cuda()`.

The .npz holds inputs, every parameter, the decode outputs for activate_level 0/1/2, and autograd
gradients of a fixed random linear functional of the outputs — the fixtures the oracle restatement
(oracle/decode_oracle.py) and the CUDA decode are pinned against.  /root/reference is NOT needed to
run the tests.
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ["diff_gaussian_rasterization", "torch_scatter", "plyfile", "simple_knn", "simple_knn._C",
                 "_gridcreater", "_gridencoder", "kornia"]:
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["diff_gaussian_rasterization"].GaussianRasterizationSettings = object
    sys.modules["diff_gaussian_rasterization"].GaussianRasterizer = object
    sys.modules["torch_scatter"].scatter_max = None
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object
    sys.modules["simple_knn._C"].distCUDA2 = None
    sys.modules["kornia"].create_meshgrid = None
    sys.path.insert(0, REF)
    import gaussian_renderer  # noqa
    from scene import gaussian_model as gm

    def get_offsets_cpu(self, resolutions_list, dim=3):
        offsets_list = [0]
        offsets = 0
        for resolution in resolutions_list:
            offsets += resolution ** dim
            offsets_list.append(offsets)
        return torch.tensor(resolutions_list, dtype=torch.int), torch.tensor(offsets_list, dtype=torch.int)

    gm.FeaturePlanes.get_offsets = get_offsets_cpu
    return gaussian_renderer, gm


class Learner(nn.Module):
    """CPU stand-in for GaussianLearner (scene/gaussian_model.py:183-215): same `inference` body."""

    def __init__(self, gm, plane_size, num_channels, Q0):
        super().__init__()
        self.Q0 = Q0
        xyz_min = torch.tensor([-2, -2, -2])
        xyz_max = torch.tensor([2, 2, 2])
        self._feat = gm.FeaturePlanes(world_size=[plane_size] * 3, xyz_min=xyz_min, xyz_max=xyz_max,
                                      feat_dim=num_channels, mlp_width=[168], out_dim=[32], subplane_multiplier=1)

    def inference(self, xyz, g_fea, Q0):
        inputs = xyz.detach()
        return self._feat(inputs, g_fea, self.Q0)


class PC:
    pass


def make_pc(gm, N, K, plane_size, num_channels, seed, appearance_dim=0, dists=False, feat_bank=False, n_cams=4):
    torch.manual_seed(seed)
    pc = PC()
    feat_dim = 32
    pc.n_offsets = K
    pc.use_feat_bank = feat_bank
    pc.appearance_dim = appearance_dim
    pc.add_opacity_dist = pc.add_cov_dist = pc.add_color_dist = dists
    dd = 1 if dists else 0
    # use_feat_bank is unusable in the reference itself: mlp_feature_bank is Linear(3+1, ...)
    # (scene/gaussian_model.py:307-313) but is fed [view(3), dist(1), geo_fea(64)] = 68 columns
    # (gaussian_renderer/__init__.py:42-44) -> shape error.  Not part of the fixtures.
    assert not feat_bank
    pc.mlp_opacity = nn.Sequential(nn.Linear(feat_dim + 3 + dd + 64, feat_dim), nn.ReLU(True),
                                   nn.Linear(feat_dim, K), nn.Tanh())
    pc.mlp_cov = nn.Sequential(nn.Linear(feat_dim + 3 + dd + 64, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, 7 * K))
    pc.mlp_color = nn.Sequential(nn.Linear(feat_dim + 3 + dd + appearance_dim + 64, feat_dim), nn.ReLU(True),
                                 nn.Linear(feat_dim, 3 * K), nn.Sigmoid())
    with torch.no_grad():
        pc.mlp_opacity[2].bias += 0.3          # ~60 % of offsets survive the mask
    pc.feat_planes = Learner(gm, plane_size, num_channels, Q0=0.0)
    # make BN affine / running stats non-trivial so the fixtures pin them
    with torch.no_grad():
        for mod in pc.feat_planes.modules():
            if isinstance(mod, nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    if appearance_dim > 0:
        from scene.embedding import Embedding
        pc.embedding_appearance = Embedding(n_cams, appearance_dim)
    g = torch.Generator().manual_seed(seed + 1)
    pc._anchor = (torch.rand(N, 3, generator=g) * 2 - 1) * 1.2
    pc._anchor[: N // 10] *= 2.2               # a shell outside the [-2,2]^3 plane bbox: zero padding
    pc._anchor.requires_grad_()
    pc._offset = (torch.randn(N, K, 3, generator=g) * 0.5).requires_grad_()
    pc._anchor_feat = (torch.randn(N, 32, generator=g) * 0.3).requires_grad_()
    s0 = 2.0 / N ** (1 / 3)
    pc._scaling = torch.log(s0 * torch.exp(torch.randn(N, 6, generator=g) * 0.3)).requires_grad_()
    pc.rotation_activation = torch.nn.functional.normalize
    type(pc).get_anchor = property(lambda self: self._anchor)
    type(pc).get_scaling = property(lambda self: 1.0 * torch.exp(self._scaling))
    type(pc).get_opacity_mlp = property(lambda self: self.mlp_opacity)
    type(pc).get_cov_mlp = property(lambda self: self.mlp_cov)
    type(pc).get_color_mlp = property(lambda self: self.mlp_color)
    type(pc).get_featurebank_mlp = property(lambda self: self.mlp_feature_bank)
    type(pc).get_appearance = property(lambda self: self.embedding_appearance)
    return pc


class Cam:
    def __init__(self, center, uid):
        self.camera_center = center
        self.uid = uid


def flat_params(pc):
    out = {}
    for name in ("mlp_opacity", "mlp_cov", "mlp_color", "mlp_feature_bank", "embedding_appearance"):
        mod = getattr(pc, name, None)
        if mod is None:
            continue
        for k, v in mod.state_dict().items():
            out[f"{name}.{k}"] = v.detach().numpy().copy()
    for k, v in pc.feat_planes._feat.state_dict().items():
        out[f"feat.{k}"] = v.detach().numpy().copy()
    return out


def named_leaf_params(pc):
    named = {"_anchor": pc._anchor, "_offset": pc._offset, "_anchor_feat": pc._anchor_feat, "_scaling": pc._scaling}
    for name in ("mlp_opacity", "mlp_cov", "mlp_color", "mlp_feature_bank", "embedding_appearance"):
        mod = getattr(pc, name, None)
        if mod is None:
            continue
        for k, v in mod.named_parameters():
            named[f"{name}.{k}"] = v
    for k, v in pc.feat_planes._feat.named_parameters():
        named[f"feat.{k}"] = v
    return named


def run_case(gr, gm, tag, N=200, K=10, plane_size=64, num_channels=15, seed=0, **variant):
    pc = make_pc(gm, N, K, plane_size, num_channels, seed, **variant)
    g = torch.Generator().manual_seed(seed + 2)
    vis = torch.rand(N, generator=g) < 0.8
    cam = Cam(torch.tensor([2.5, -1.5, 0.7]), uid=2)
    data = dict(N=N, K=K, plane_size=plane_size, num_channels=num_channels, vis=vis.numpy(),
                cam_center=cam.camera_center.numpy(), uid=cam.uid,
                appearance_dim=variant.get("appearance_dim", 0), dists=int(variant.get("dists", False)),
                feat_bank=int(variant.get("feat_bank", False)))
    data.update({f"param.{k}": v for k, v in flat_params(pc).items()})
    for k in ("_anchor", "_offset", "_anchor_feat", "_scaling"):
        data[f"in.{k}"] = getattr(pc, k).detach().numpy().copy()
    names = ["xyz", "color", "opacity", "scaling", "rot", "neural_opacity", "mask"]
    for level in (0, 1, 2):
        pc.feat_planes._feat.activate_level = level
        for p in named_leaf_params(pc).values():
            p.grad = None
        outs = gr.generate_neural_gaussians(cam, pc, vis, is_training=True)
        # BN running stats are updated by this call (train mode); record them after level 2 only
        gl = torch.Generator().manual_seed(seed + 10 + level)
        loss = 0
        for nm, t in zip(names[:5], outs[:5]):
            w = torch.randn(t.shape, generator=gl)
            data[f"L{level}.w.{nm}"] = w.numpy()
            loss = loss + (t * w).sum()
            data[f"L{level}.out.{nm}"] = t.detach().numpy().copy()
        data[f"L{level}.out.neural_opacity"] = outs[5].detach().numpy().copy()
        data[f"L{level}.out.mask"] = outs[6].numpy().copy()
        loss.backward()
        for k, p in named_leaf_params(pc).items():
            if p.grad is not None:
                data[f"L{level}.grad.{k}"] = p.grad.detach().numpy().copy()
    for k, v in pc.feat_planes._feat.state_dict().items():
        if "running" in k or "num_batches" in k:
            data[f"after.feat.{k}"] = v.detach().numpy().copy()
    path = os.path.join(OUT, f"decode_{tag}.npz")
    np.savez_compressed(path, **data)
    print(tag, "M per level:", [int(data[f"L{l}.out.mask"].sum()) for l in (0, 1, 2)], os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    gr, gm = import_reference()
    torch.set_num_threads(4)
    run_case(gr, gm, "base", seed=0)
    run_case(gr, gm, "variants", N=120, seed=5, plane_size=48, appearance_dim=8, dists=True)
