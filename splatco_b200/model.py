"""Host-side mirror of the slice of the reference model that the render hot path reads.

The decode kernels take their weights from `nn.Module` parameters with the SAME module tree and
state-dict key names as the reference, so reference checkpoints (`feat_planes.state_dict()`,
train.py:313-316) load unchanged and a reference `GaussianModel` can be passed to
`splatco_b200.gaussian_renderer.render` directly.  These classes exist so tests, bench.py and the
synthetic scenes have a model to hand to the drop-in without the reference tree on the path; they
hold parameters only — the arithmetic is in libsplatco_b200.so (except TriPlaneAttention, see
decode.py).

Mirrors (file:line in /root/reference):
  scene/grids.py:22-64      ChannelAttention / SpatialAttention / TriPlaneAttention
  scene/grids.py:102-131    PlaneGrid parameters (xy/xz/yz planes [1, C//3, E, E], xyz_min/max buffers)
  scene/gaussian_model.py:97-147   FeaturePlanes (k0s = [TA@P/4, P/4, P/2, P], models, CTX_models)
  scene/gaussian_model.py:183-215  GaussianLearner (Q0, _feat, bbox [-2,2]^3)
  scene/gaussian_model.py:307-337  MLP heads;  :403-441 accessors
  scene/gaussian_model.py:510-572  training_setup (statistics + Adam parameter groups), :733-758 / :784-830 the optimizer
                                   bookkeeping of growing / pruning anchors — so the densification drop-ins
                                   (splatco_b200.densify) can be exercised end to end without the reference tree
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class ChannelAttention(nn.Module):
    def __init__(self, in_planes, ratio=5):
        super().__init__()
        self.sharedMLP = nn.Sequential(nn.Conv2d(in_planes, in_planes // ratio, 1, bias=False), nn.ReLU(),
                                       nn.Conv2d(in_planes // ratio, in_planes, 1, bias=False))

    def forward(self, x):
        avg = self.sharedMLP(F.adaptive_avg_pool2d(x, 1))
        mx = self.sharedMLP(F.adaptive_max_pool2d(x, 1))
        return torch.sigmoid(avg + mx)


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=3 if kernel_size == 7 else 1, bias=False)

    def forward(self, x):
        s = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)
        return torch.sigmoid(self.conv(s))


class TriPlaneAttention(nn.Module):
    def __init__(self, planes):
        super().__init__()
        self.ca = ChannelAttention(planes)
        self.sa = SpatialAttention()

    def forward(self, x):
        x = self.ca(x) * x
        return self.sa(x) * x


class PlaneGrid(nn.Module):
    def __init__(self, channels, edge, xyz_min, xyz_max, TAflag=False):
        super().__init__()
        self.channels = channels
        self.TAflag = TAflag
        self.register_buffer("xyz_min", torch.tensor(xyz_min, dtype=torch.float32))
        self.register_buffer("xyz_max", torch.tensor(xyz_max, dtype=torch.float32))
        R = channels // 3
        self.xy_plane = nn.Parameter(torch.randn(1, R, edge, edge) * 0.1)
        self.xz_plane = nn.Parameter(torch.randn(1, R, edge, edge) * 0.1)
        self.yz_plane = nn.Parameter(torch.randn(1, R, edge, edge) * 0.1)
        if TAflag:
            self.TA = TriPlaneAttention(channels)

    def get_dim(self):
        return self.channels * 2 if self.TAflag else self.channels


class FeaturePlanes(nn.Module):
    def __init__(self, plane_size, num_channels, xyz_min=(-2.0, -2.0, -2.0), xyz_max=(2.0, 2.0, 2.0),
                 allocate_unused_full_res=False):
        super().__init__()
        self.activate_level = 0
        self.num_levels = 3
        edges = [int(plane_size * 0.5 ** (2 - i)) for i in range(3)]
        self.k0s = nn.ModuleList()
        for i, e in enumerate(edges):
            if i == 0:
                self.k0s.append(PlaneGrid(num_channels, e, xyz_min, xyz_max, TAflag=True))
            if i < 2 or allocate_unused_full_res:
                # the reference also allocates k0s[3] (full resolution) but never samples it
                self.k0s.append(PlaneGrid(num_channels, e, xyz_min, xyz_max))
        self.models = nn.ModuleList()
        self.CTX_models = nn.ModuleList()
        for i in range(3):
            d = self.k0s[i].get_dim()
            self.models.append(nn.Sequential(nn.BatchNorm1d(d), nn.Linear(d, 32)))
            self.CTX_models.append(nn.Sequential(nn.BatchNorm1d(71), nn.Linear(71, 32)))


class GaussianLearner(nn.Module):
    def __init__(self, plane_size, num_channels, allocate_unused_full_res=False):
        super().__init__()
        self.Q0 = 0.03
        self._feat = FeaturePlanes(plane_size, num_channels, allocate_unused_full_res=allocate_unused_full_res)

    def activate_plane_level(self):
        self._feat.activate_level += 1


class AnchorModel:
    """Duck-type of the reference GaussianModel as read by gaussian_renderer."""

    def __init__(self, N, n_offsets=10, feat_dim=32, plane_size=256, num_channels=15, appearance_dim=0,
                 num_cameras=1, add_opacity_dist=False, add_cov_dist=False, add_color_dist=False,
                 device="cuda", seed=0, scale_factor=2.0):
        g = torch.Generator().manual_seed(seed)
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            self.feat_dim, self.n_offsets = feat_dim, n_offsets
            self.use_feat_bank = False
            self.appearance_dim = appearance_dim
            self.add_opacity_dist, self.add_cov_dist, self.add_color_dist = add_opacity_dist, add_cov_dist, add_color_dist
            od, cd, ld = int(add_opacity_dist), int(add_cov_dist), int(add_color_dist)
            self.mlp_opacity = nn.Sequential(nn.Linear(feat_dim + 3 + od + 64, feat_dim), nn.ReLU(True),
                                             nn.Linear(feat_dim, n_offsets), nn.Tanh())
            self.mlp_cov = nn.Sequential(nn.Linear(feat_dim + 3 + cd + 64, feat_dim), nn.ReLU(True),
                                         nn.Linear(feat_dim, 7 * n_offsets))
            self.mlp_color = nn.Sequential(nn.Linear(feat_dim + 3 + ld + appearance_dim + 64, feat_dim), nn.ReLU(True),
                                           nn.Linear(feat_dim, 3 * n_offsets), nn.Sigmoid())
            with torch.no_grad():
                self.mlp_opacity[2].bias += 0.3      # ~60 % of offsets survive the opacity mask
            self.feat_planes = GaussianLearner(plane_size, num_channels)
            self.embedding_appearance = None
            if appearance_dim > 0:
                emb = nn.Module()
                emb.embedding = nn.Embedding(num_cameras, appearance_dim)
                self.embedding_appearance = emb
        anchor = torch.rand(N, 3, generator=g) * 2 - 1
        s0 = scale_factor / max(N, 1) ** (1.0 / 3.0)     # scale_factor 2.0 = mean anchor spacing
        self._anchor = nn.Parameter(anchor.to(device))
        self._offset = nn.Parameter((torch.randn(N, n_offsets, 3, generator=g) * 0.5).to(device))
        self._anchor_feat = nn.Parameter((torch.randn(N, feat_dim, generator=g) * 0.3).to(device))
        self._scaling = nn.Parameter(torch.log(s0 * torch.exp(torch.randn(N, 6, generator=g) * 0.3)).to(device))
        rot = torch.zeros(N, 4)
        rot[:, 0] = 1
        self._rotation = nn.Parameter(rot.to(device), requires_grad=False)
        self.rotation_activation = F.normalize
        for m in self.modules():
            m.to(device)

    def modules(self):
        ms = [self.mlp_opacity, self.mlp_cov, self.mlp_color, self.feat_planes]
        if self.embedding_appearance is not None:
            ms.append(self.embedding_appearance)
        return ms

    def parameters(self):
        ps = [self._anchor, self._offset, self._anchor_feat, self._scaling]
        for m in self.modules():
            ps += list(m.parameters())
        return ps

    # accessors, scene/gaussian_model.py:403-441
    @property
    def get_anchor(self):
        return self._anchor

    @property
    def get_scaling(self):
        return 1.0 * torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_opacity_mlp(self):
        return self.mlp_opacity

    @property
    def get_cov_mlp(self):
        return self.mlp_cov

    @property
    def get_color_mlp(self):
        return self.mlp_color

    @property
    def get_appearance(self):
        return self.embedding_appearance

    # ---- training-side bookkeeping (what densify.anchor_growing / adjust_anchor and statis.training_statis touch) ----
    PER_ANCHOR = ("anchor", "offset", "anchor_feat", "opacity", "scaling", "rotation")
    _SKIP = ("mlp", "conv", "feat_base", "embedding", "feat_planes")

    def training_setup(self, voxel_size=0.01, update_depth=3, update_init_factor=16, update_hierachy_factor=4, fused=True):
        """Statistics and Adam parameter groups as GaussianModel.training_setup creates them (scene/gaussian_model.py:510-572;
        learning rates = arguments/__init__.py defaults), with splatco_b200.optim.FusedAdam (or torch.optim.Adam)."""
        dev, N = self._anchor.device, int(self._anchor.shape[0])
        self.voxel_size, self.update_depth = voxel_size, update_depth
        self.update_init_factor, self.update_hierachy_factor = update_init_factor, update_hierachy_factor
        if not hasattr(self, "_opacity"):
            self._opacity = nn.Parameter(torch.full((N, 1), -2.1972246, device=dev), requires_grad=False)   # inverse_sigmoid(0.1)
        self.opacity_accum = torch.zeros(N, 1, device=dev)
        self.anchor_demon = torch.zeros(N, 1, device=dev)
        self.offset_gradient_accum = torch.zeros(N * self.n_offsets, 1, device=dev)
        self.offset_denom = torch.zeros(N * self.n_offsets, 1, device=dev)
        self.max_radii2D = torch.zeros(N, device=dev)
        groups = [{"params": [self._anchor], "lr": 0.0, "name": "anchor"}, {"params": [self._offset], "lr": 0.01, "name": "offset"},
                  {"params": [self._anchor_feat], "lr": 0.0075, "name": "anchor_feat"}, {"params": [self._opacity], "lr": 0.02, "name": "opacity"},
                  {"params": [self._scaling], "lr": 0.007, "name": "scaling"}, {"params": [self._rotation], "lr": 0.002, "name": "rotation"},
                  {"params": list(self.mlp_opacity.parameters()), "lr": 0.002, "name": "mlp_opacity"},
                  {"params": list(self.mlp_cov.parameters()), "lr": 0.004, "name": "mlp_cov"},
                  {"params": list(self.mlp_color.parameters()), "lr": 0.008, "name": "mlp_color"}]
        if self.embedding_appearance is not None:
            groups.append({"params": list(self.embedding_appearance.parameters()), "lr": 0.05, "name": "embedding_appearance"})
        feat = self.feat_planes._feat
        for i in range(3):
            active = i == feat.activate_level
            groups.append({"params": list(feat.k0s[i].parameters()), "lr": 0.01 if active else 0.001, "name": f"feat_planes{i}"})
            groups.append({"params": list(feat.models[i].parameters()), "lr": 1e-4 if active else 1e-5, "name": f"fp_mlp_f{i}"})
        if fused:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
        else:
            self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        return self.optimizer

    def _swap_rows(self, new_rows):
        """Replace the per-anchor Parameters (and their Adam moments) by `new_rows[name](old tensor)`."""
        out = {}
        for group in self.optimizer.param_groups:
            if any(k in group["name"] for k in self._SKIP):
                continue
            old = group["params"][0]
            state = self.optimizer.state.pop(old, None)
            fn = new_rows[group["name"]]
            new = nn.Parameter(fn(old.detach(), False), requires_grad=old.requires_grad)
            if state is not None and len(state) > 0:
                state["exp_avg"] = fn(state["exp_avg"], True)
                state["exp_avg_sq"] = fn(state["exp_avg_sq"], True)
                self.optimizer.state[new] = state
            group["params"][0] = new
            out[group["name"]] = new
        self._anchor, self._offset, self._anchor_feat = out["anchor"], out["offset"], out["anchor_feat"]
        self._opacity, self._scaling, self._rotation = out["opacity"], out["scaling"], out["rotation"]
        return out

    def cat_tensors_to_optimizer(self, tensors_dict):
        """scene/gaussian_model.py:733-758: append rows to every per-anchor Parameter; their Adam moments get zero rows."""
        def rows(name):
            ext = tensors_dict[name]
            return lambda t, is_moment: torch.cat([t, torch.zeros_like(ext) if is_moment else ext.to(t.dtype)], dim=0)
        return self._swap_rows({n: rows(n) for n in self.PER_ANCHOR})

    def prune_anchor(self, mask):
        """scene/gaussian_model.py:784-830: drop the masked anchors (and their moments); the reference also clamps the
        kept rows' scaling[:, 3:] at 0.05 here (:796-800), reproduced because the decode reads those columns."""
        keep = ~mask

        def rows(name):
            def fn(t, is_moment):
                t = t[keep]
                if name == "scaling" and not is_moment:
                    t = t.clone()
                    t[:, 3:] = t[:, 3:].clamp(max=0.05)
                return t
            return fn
        return self._swap_rows({n: rows(n) for n in self.PER_ANCHOR})

    def training_statis(self, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
        from .statis import training_statis
        return training_statis(self, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask)

    def anchor_growing(self, grads, threshold, offset_mask):
        from .densify import anchor_growing
        return anchor_growing(self, grads, threshold, offset_mask)

    def adjust_anchor(self, iteration, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005):
        from .densify import adjust_anchor
        return adjust_anchor(self, iteration, check_interval, success_threshold, grad_threshold, min_opacity)

    def eval(self):           # scene/gaussian_model.py:350-357: MLP heads only; feat_planes stays in train mode
        self.mlp_opacity.eval(); self.mlp_cov.eval(); self.mlp_color.eval()

    def train(self):
        self.mlp_opacity.train(); self.mlp_cov.train(); self.mlp_color.train()
