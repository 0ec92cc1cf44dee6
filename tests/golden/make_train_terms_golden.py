"""Generates tests/golden/train_terms.npz by running the REFERENCE's own Python on CPU (authoring container only):

  tv.*     scene/grids.py PlaneGrid.total_variation_add_grad (:240-250) and GaussianLearner.tv_loss
           (scene/gaussian_model.py:217-220) — unmodified methods on reference PlaneGrid modules with a pre-existing .grad
  mvc.*    the cross-view pair loop of train.py:199-216 evaluated with the reference's own utils/loss_utils.py
           (l1_loss, ssim) and its own align_images (train.py:79-96, exec'd from the reference file — train.py itself cannot
           be imported here: lpips / the rasterizer are absent); value + autograd gradients w.r.t. the generated images
  grow.*   GaussianModel.anchor_growing (:832-925) and adjust_anchor (:929-997), unmodified, on a bare GaussianModel
           instance with a real Adam optimizer.  Patches needed to run them without a GPU: Tensor.cuda -> identity,
           device='cuda' keyword dropped from torch.zeros/ones, torch_scatter.scatter_max -> Tensor.scatter_reduce('amax')
           (same result), torch.rand_like recorded so the tests can replay the random thinning masks.
           voxel sizes are powers of two, so torch's CPU division and its CUDA reciprocal-multiply round identically.

    python tests/golden/make_train_terms_golden.py
"""
import os
import re
import sys
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_decode_golden import import_reference, REF  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def tv_fixtures(gm, out):
    from scene.grids import PlaneGrid
    torch.manual_seed(7)
    grids = []
    for n, (edge, amp, ta) in enumerate([(9, 1.0, True), (9, 12.0, False), (18, 6.0, False)]):
        pg = PlaneGrid(15, [edge, edge + 2, edge + 5], [-2, -2, -2], [2, 2, 2], {"factor": 1}, TAflag=ta)
        with torch.no_grad():
            for p in (pg.xy_plane, pg.xz_plane, pg.yz_plane):
                p.mul_(amp)                     # amp 12: |differences| > 1 are common (the clamped branch)
        for name in ("xy_plane", "xz_plane", "yz_plane"):
            p = getattr(pg, name)
            p.grad = torch.randn_like(p) * 1e-3
            out[f"tv.g{n}.{name}"] = p.detach().numpy().copy()
            out[f"tv.g{n}.{name}.grad0"] = p.grad.numpy().copy()
        grids.append(pg)
    # single grid, direct call
    grids[1].total_variation_add_grad(1e-3)
    for name in ("xy_plane", "xz_plane", "yz_plane"):
        out[f"tv.g1.{name}.grad_direct"] = getattr(grids[1], name).grad.numpy().copy()
        getattr(grids[1], name).grad = torch.from_numpy(out[f"tv.g1.{name}.grad0"].copy())
    # tv_loss over the levels (k0s = [TA grid, plain, plain]; activate_level 2 -> weights w/4, w/2, w)
    me = SimpleNamespace(_feat=SimpleNamespace(activate_level=2, k0s=grids))
    gm.GaussianLearner.tv_loss(me, 2e-3)
    for n in range(3):
        for name in ("xy_plane", "xz_plane", "yz_plane"):
            out[f"tv.g{n}.{name}.grad_tvloss"] = getattr(grids[n], name).grad.numpy().copy()
    out["tv.w_direct"], out["tv.w_tvloss"] = 1e-3, 2e-3


def mvc_fixtures(out):
    from utils.loss_utils import l1_loss, ssim
    src = open(os.path.join(REF, "train.py")).read()
    m = re.search(r"^def align_images\(.*?^    return [^\n]*\n", src, re.S | re.M) or re.search(r"def align_images\(.*?return img1_aligned[^\n]*\n", src, re.S)
    ns = {}
    exec(m.group(0), ns)
    align_images = ns["align_images"]
    g = torch.Generator().manual_seed(31)
    for case, sizes in enumerate([[(40, 56)] * 4, [(33, 47), (35, 45), (33, 50)]]):
        n = len(sizes)
        base = torch.rand(3, 64, 64, generator=g)
        reals, gens = [], []
        for v, (H, W) in enumerate(sizes):
            # views 0..n-2 are near-copies of one image (pair SSIM > 0.6), the last one is unrelated (gate closed)
            r = (base[:, :H, :W] + 0.03 * torch.randn(3, H, W, generator=g)).clamp(0, 1) if v < n - 1 else torch.rand(3, H, W, generator=g)
            reals.append(r)
            gens.append((r + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1).requires_grad_())
        imgs = [[gens[i], reals[i]] for i in range(n)]
        total, parts, ssims = 0, [], []
        for i in range(n):                                   # train.py:203-216,236
            for j in range(i + 1, n):
                gen_img1, real_img1, gen_img2, real_img2 = imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1]
                gen_img1, gen_img2, real_img1, real_img2 = align_images(gen_img1, gen_img2, real_img1, real_img2)
                s = ssim(real_img1, real_img2)
                if s > 0.6:
                    loss_tmp = ssim(real_img1, real_img2) * torch.abs(l1_loss(real_img1 - real_img2, gen_img1 - gen_img2))
                else:
                    loss_tmp = 0
                total = total + loss_tmp
                parts.append(float(loss_tmp.detach()) if torch.is_tensor(loss_tmp) else 0.0)
                ssims.append(float(s))
        (0.05 * total).backward()
        out[f"mvc.c{case}.n"] = n
        out[f"mvc.c{case}.loss"] = float(total)
        out[f"mvc.c{case}.parts"] = np.array(parts, np.float64)
        out[f"mvc.c{case}.ssim"] = np.array(ssims, np.float64)
        for v in range(n):
            out[f"mvc.c{case}.real{v}"] = reals[v].numpy()
            out[f"mvc.c{case}.gen{v}"] = gens[v].detach().numpy()
            out[f"mvc.c{case}.grad{v}"] = (gens[v].grad if gens[v].grad is not None else torch.zeros_like(gens[v])).numpy()


class _CpuPatches:
    """Lets the reference's hard-coded .cuda() / device='cuda' run on CPU and records torch.rand_like draws."""

    def __enter__(self):
        self.saved = (torch.Tensor.cuda, torch.zeros, torch.ones, torch.rand_like)
        self.rands = []
        torch.Tensor.cuda = lambda t, *a, **k: t
        strip = lambda fn: (lambda *a, **k: fn(*a, **{kk: vv for kk, vv in k.items() if not (kk == "device" and str(vv).startswith("cuda"))}))
        torch.zeros, torch.ones = strip(self.saved[1]), strip(self.saved[2])
        rl = self.saved[3]

        def rand_like(t, *a, **k):
            r = rl(t, *a, **k)
            self.rands.append(r.clone())
            return r
        torch.rand_like = rand_like
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda, torch.zeros, torch.ones, torch.rand_like = self.saved


def scatter_max(src, index, dim=0):
    n = int(index.max()) + 1
    o = torch.zeros(n, src.shape[1], dtype=src.dtype).scatter_reduce(0, index, src, "amax", include_self=False)
    return o, None


def bare_model(gm, N, K, F, voxel_size, seed, clustered):
    g = torch.Generator().manual_seed(seed)
    me = object.__new__(gm.GaussianModel)
    me.n_offsets, me.feat_dim, me.voxel_size = K, F, voxel_size
    me.update_depth, me.update_init_factor, me.update_hierachy_factor = 3, 16, 4
    me.scaling_activation = torch.exp
    anchor = torch.rand(N, 3, generator=g) * 2 - 1
    if clustered:                                    # many anchors exactly on voxel centres of the coarsest grid
        anchor[: N // 2] = torch.round(anchor[: N // 2] / (voxel_size * 16)) * (voxel_size * 16)
    me._anchor = nn.Parameter(anchor)
    me._offset = nn.Parameter(torch.randn(N, K, 3, generator=g) * 0.5)
    me._anchor_feat = nn.Parameter(torch.randn(N, F, generator=g) * 0.3)
    me._scaling = nn.Parameter(torch.log(0.08 * torch.exp(torch.randn(N, 6, generator=g) * 0.3)))
    me._rotation = nn.Parameter(torch.zeros(N, 4))
    me._opacity = nn.Parameter(torch.zeros(N, 1))
    groups = [{"params": [getattr(me, "_" + n)], "lr": 1e-3, "name": n} for n in ("anchor", "offset", "anchor_feat", "opacity", "scaling", "rotation")]
    me.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    me.opacity_accum = torch.rand(N, 1, generator=g) * 2.0
    me.anchor_demon = torch.randint(60, 120, (N, 1), generator=g).float()
    me.offset_denom = torch.randint(0, 100, (N * K, 1), generator=g).float()
    me.offset_gradient_accum = me.offset_denom * torch.rand(N * K, 1, generator=g) * 2e-3
    return me


MODEL_KEYS = ("_anchor", "_offset", "_anchor_feat", "_scaling", "_rotation", "_opacity")
STAT_KEYS = ("opacity_accum", "anchor_demon", "offset_denom", "offset_gradient_accum")


def snapshot(me, out, prefix, full=True):
    for k in MODEL_KEYS + STAT_KEYS:
        v = getattr(me, k).detach().numpy().copy()
        if not full and k == "_offset":            # rows of grown anchors are zeros, pruning drops rows: the row sums pin both
            out[f"{prefix}.{k}.rowsum"] = v.astype(np.float64).sum(axis=(1, 2))
            continue
        out[f"{prefix}.{k}"] = v


def grow_fixtures(gm, out):
    gm.scatter_max = scatter_max
    for case, (N, clustered, vs) in enumerate([(700, False, 2.0 ** -7), (1200, True, 2.0 ** -8)]):
        K, F = 10, 32
        me = bare_model(gm, N, K, F, vs, 100 + case, clustered)
        snapshot(me, out, f"grow.c{case}.in")
        out[f"grow.c{case}.voxel_size"] = vs
        # (a) anchor_growing alone
        grads = (me.offset_gradient_accum / me.offset_denom)
        grads[grads.isnan()] = 0.0
        grads_norm = torch.norm(grads, dim=-1)
        offset_mask = (me.offset_denom > 40).squeeze(1)
        torch.manual_seed(5 + case)
        with _CpuPatches() as pt, torch.no_grad():
            gm.GaussianModel.anchor_growing(me, grads_norm, 0.0002, offset_mask)
        out[f"grow.c{case}.grads_norm"], out[f"grow.c{case}.offset_mask"] = grads_norm.numpy(), offset_mask.numpy()
        for i, r in enumerate(pt.rands):
            out[f"grow.c{case}.rand{i}"] = r.numpy()
        out[f"grow.c{case}.n_rand"] = len(pt.rands)
        snapshot(me, out, f"grow.c{case}.grown", full=False)
        # (b) the whole adjust_anchor on a fresh copy (iteration 1700: no curvature pass)
        me = bare_model(gm, N, K, F, vs, 100 + case, clustered)
        torch.manual_seed(5 + case)
        with _CpuPatches() as pt, torch.no_grad():
            gm.GaussianModel.adjust_anchor(me, iteration=1700, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005)
        snapshot(me, out, f"grow.c{case}.adjusted", full=False)
        print(f"grow case {case}: N {N} -> grown {out[f'grow.c{case}.grown._anchor'].shape[0]} -> adjusted {me._anchor.shape[0]}")


def cvpm_fixtures(gm, out):
    """GaussianModel.compute_fast_loss_with_key_points (:1112-1219), unmodified (it never touches self), device='cpu'."""
    g = torch.Generator().manual_seed(61)
    t1, t2 = torch.tensor([-1.5, 0.2, 0.1]), torch.tensor([1.2, -0.3, 0.4])
    d = (t2 - t1) / (t2 - t1).norm()
    N = 30000
    cloud = torch.rand(N, 3, generator=g) * 4 - 2
    s = torch.rand(4000, 1, generator=g) * 16 - 8                       # along the line, far beyond both cameras (3-sigma outliers)
    cloud[:4000] = t1 + d * s + torch.randn(4000, 3, generator=g) * 0.03
    cloud[4000:4500] = (t1 if True else t2) + torch.randn(500, 3, generator=g) * 0.2
    cloud[4500:5000] = t2 + torch.randn(500, 3, generator=g) * 0.2
    base = torch.rand(3, 40, 52, generator=g)
    real1 = (base + 0.03 * torch.randn(3, 40, 52, generator=g)).clamp(0, 1)
    real2 = (base[:, :38, :50] + 0.03 * torch.randn(3, 38, 50, generator=g)).clamp(0, 1)
    other = torch.rand(3, 38, 50, generator=g)
    gen1, gen2 = (real1 + 0.1 * torch.randn(3, 40, 52, generator=g)).clamp(0, 1), (real2 + 0.1 * torch.randn(3, 38, 50, generator=g)).clamp(0, 1)
    K = R = torch.eye(3)
    out.update({"cvpm.cloud": cloud.numpy(), "cvpm.t1": t1.numpy(), "cvpm.t2": t2.numpy(), "cvpm.real1": real1.numpy(), "cvpm.real2": real2.numpy(),
                "cvpm.other": other.numpy(), "cvpm.gen1": gen1.numpy(), "cvpm.gen2": gen2.numpy()})
    for name, r2, thr in (("open", real2, 0.05), ("tight", real2, 0.004), ("gated", other, 0.05)):
        a, b, pts, mask = gm.GaussianModel.compute_fast_loss_with_key_points(None, real1, r2, gen1, gen2, K, R, t1.clone(), K, R, t2.clone(), cloud,
                                                                             distance_threshold=thr, overall_ssim_threshold=0.6, device="cpu")
        out[f"cvpm.{name}.thr"] = thr
        out[f"cvpm.{name}.gen_l1"], out[f"cvpm.{name}.cross_l1"] = float(a), float(b)
        out[f"cvpm.{name}.mask"] = np.packbits(mask.numpy())
        out[f"cvpm.{name}.n_points"] = int(pts.shape[0])
        print(f"cvpm {name}: mask.sum() = {int(mask.sum())}, losses {float(a):.6f} {float(b):.6f}")


def main():
    _, gm = import_reference()
    out = {}
    cvpm_fixtures(gm, out)
    tv_fixtures(gm, out)
    mvc_fixtures(out)
    grow_fixtures(gm, out)
    np.savez_compressed(os.path.join(OUT, "train_terms.npz"), **out)
    print("wrote train_terms.npz:", len(out), "arrays;", {k: v for k, v in out.items() if k.endswith(".loss") or k.endswith(".parts") or k.endswith(".ssim")})


if __name__ == "__main__":
    main()
