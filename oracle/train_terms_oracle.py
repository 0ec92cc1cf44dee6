"""TEST INFRASTRUCTURE — CPU restatements (numpy) of the training-loop terms around the render path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

Pinned against tests/golden/train_terms.npz, which tests/golden/make_train_terms_golden.py produced by running the
reference's own Python (tests/test_oracle_train_terms.py).  Restated from (file:line under /root/reference):
  tv_grad                 scene/grids.py:240-250 (PlaneGrid.total_variation_add_grad), scene/gaussian_model.py:217-220
  ssim / l1               utils/loss_utils.py:17-18, 20-63
  mv_consistency          train.py:79-96 (align_images), :199-216 (pair loop), :237-239
  grow_pass               scene/gaussian_model.py:839-897 (one pass of anchor_growing's loop)
  anchor_growing          scene/gaussian_model.py:832-925
  adjust_anchor_stats     scene/gaussian_model.py:929-997 (statistics resets, prune mask)
"""
from __future__ import annotations

import numpy as np


# ---- total variation ----------------------------------------------------------------------------------------------
def tv_grad(plane: np.ndarray, w: float) -> np.ndarray:
    """d/dplane of  w/6 * (smooth_l1_sum(p[..., 1:, :], p[..., :-1, :]) + smooth_l1_sum(p[..., 1:], p[..., :-1])), beta = 1."""
    p = plane.astype(np.float64)
    g = np.zeros_like(p)
    dr = np.clip(p[..., 1:, :] - p[..., :-1, :], -1.0, 1.0)        # d smooth_l1 / d(input) = clamp(diff, -1, 1)
    g[..., 1:, :] += dr
    g[..., :-1, :] -= dr
    dc = np.clip(p[..., :, 1:] - p[..., :, :-1], -1.0, 1.0)
    g[..., :, 1:] += dc
    g[..., :, :-1] -= dc
    return g * (w / 6.0)


# ---- image terms ----------------------------------------------------------------------------------------------------
def _window():
    x = np.arange(11, dtype=np.float64)
    g = np.exp(-((x - 5) ** 2) / (2 * 1.5 ** 2))
    return g / g.sum()


def _blur(img):
    """conv2d with the 11x11 Gaussian window, zero padding 5, per channel (utils/loss_utils.py:44-46)."""
    g = _window()
    C, H, W = img.shape
    pad = np.zeros((C, H + 10, W + 10))
    pad[:, 5:-5, 5:-5] = img
    tmp = sum(g[k] * pad[:, :, k:k + W] for k in range(11))
    return sum(g[k] * tmp[:, k:k + H, :] for k in range(11))


def ssim(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    mu1, mu2 = _blur(a), _blur(b)
    s1, s2, s12 = _blur(a * a) - mu1 * mu1, _blur(b * b) - mu2 * mu2, _blur(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(m.mean())


def mv_consistency(gens, reals, gate=0.6):
    """Returns (sum of pair losses, per-pair losses, per-pair ssim, d(sum)/d gen_i as full-size arrays).
    align_images crops the four images of a PAIR to their common top-left region (train.py:79-96,209)."""
    n = len(gens)
    grads = [np.zeros(g.shape, np.float64) for g in gens]
    parts, ssims = [], []
    for i in range(n):
        for j in range(i + 1, n):
            H = min(gens[i].shape[1], gens[j].shape[1], reals[i].shape[1], reals[j].shape[1])
            W = min(gens[i].shape[2], gens[j].shape[2], reals[i].shape[2], reals[j].shape[2])
            r1, r2 = reals[i][:, :H, :W].astype(np.float64), reals[j][:, :H, :W].astype(np.float64)
            g1, g2 = gens[i][:, :H, :W].astype(np.float64), gens[j][:, :H, :W].astype(np.float64)
            s = ssim(r1, r2)
            ssims.append(s)
            if not s > gate:
                parts.append(0.0)
                continue
            e = (r1 - r2) - (g1 - g2)
            parts.append(s * np.abs(e).mean())
            sg = np.sign(e) * (s / e.size)
            grads[i][:, :H, :W] -= sg
            grads[j][:, :H, :W] += sg
    return float(sum(parts)), np.array(parts), np.array(ssims), grads


# ---- anchor growing -------------------------------------------------------------------------------------------------
def grid_coords(x: np.ndarray, cur_size: float, div_mode: int) -> np.ndarray:
    """round(x / cur_size).int() in fp32.  div_mode 0: x * (1/cur_size) (torch CUDA scalar division), 1: x / cur_size."""
    x = x.astype(np.float32)
    cs = np.float32(cur_size)
    q = x / cs if div_mode else x * (np.float32(1.0) / cs)
    return np.rint(q.astype(np.float32)).astype(np.int32)          # np.rint: half to even, like torch.round


def grow_pass(anchor, offset, scaling, anchor_feat, candidate_mask, cur_size, div_mode=0):
    """One pass (gaussian_model.py:855-897): returns (candidate_anchor [U,3] fp32, new_feat [U,F] fp32)."""
    anchor, offset, scaling = anchor.astype(np.float32), offset.astype(np.float32), scaling.astype(np.float32)
    N, K = offset.shape[:2]
    mask = np.zeros(N * K, bool)
    mask[: candidate_mask.shape[0]] = candidate_mask
    all_xyz = anchor[:, None, :] + offset * scaling[:, None, :3]                       # :855 (mul, then add, fp32)
    existing = grid_coords(anchor, cur_size, div_mode)                                  # :862
    sel = grid_coords(all_xyz.reshape(-1, 3)[mask], cur_size, div_mode)                # :864-865
    F = anchor_feat.shape[1]
    if sel.shape[0] == 0:
        return np.zeros((0, 3), np.float32), np.zeros((0, F), np.float32)
    uniq, inverse = np.unique(sel, axis=0, return_inverse=True)                         # :867 (rows sorted lexicographically)
    inverse = inverse.reshape(-1)
    have = set(map(tuple, existing.tolist()))
    keep = np.array([tuple(u) not in have for u in uniq.tolist()], bool)               # :871-884
    cand_anchor = uniq[keep].astype(np.float32) * np.float32(cur_size)                  # :885
    feat = np.repeat(anchor_feat.astype(np.float32), K, axis=0)[mask]                   # :895
    fmax = np.full((uniq.shape[0], F), -np.inf, np.float32)
    np.maximum.at(fmax, inverse, feat)                                                  # :897 scatter_max
    return cand_anchor, fmax[keep]


def anchor_growing(anchor, offset, log_scaling, anchor_feat, grads, threshold, offset_mask, rands, voxel_size,
                   update_depth=3, update_init_factor=16, update_hierachy_factor=4, div_mode=0):
    """The loop of gaussian_model.py:832-925 with the recorded random draws; returns the grown (anchor, offset,
    log_scaling, anchor_feat) and the number of anchors each pass added."""
    K = offset.shape[1]
    n_stat = anchor.shape[0] * K
    added = []
    for i in range(update_depth):
        cur_threshold = threshold * ((update_hierachy_factor // 2) ** i)
        cand = (grads >= np.float32(cur_threshold)) & offset_mask & (rands[i] > np.float32(0.5 ** (i + 1)))
        if anchor.shape[0] * K - n_stat == 0 and i > 0:
            added.append(0)
            continue
        cur_size = voxel_size * (update_init_factor // (update_hierachy_factor ** i))
        new_anchor, new_feat = grow_pass(anchor, offset, np.exp(log_scaling.astype(np.float32)), anchor_feat, cand, cur_size, div_mode)
        U = new_anchor.shape[0]
        added.append(U)
        if U == 0:
            continue
        anchor = np.concatenate([anchor, new_anchor])
        offset = np.concatenate([offset, np.zeros((U, K, 3), np.float32)])
        log_scaling = np.concatenate([log_scaling, np.log(np.full((U, 6), cur_size, np.float32))])
        anchor_feat = np.concatenate([anchor_feat, new_feat])
    return anchor, offset, log_scaling, anchor_feat, added


# ---- CVPM pruning mask ----------------------------------------------------------------------------------------------
def cvpm_mask(points, t1, t2, ssim_value=None, distance_threshold=0.01, overall_ssim_threshold=0.6, sigma_threshold=3.0,
              min_cam_distance=0.5):
    """scene/gaussian_model.py:1163-1165,1178-1210 in fp32 numpy (same operation order as the reference's torch ops)."""
    P = points.astype(np.float32)
    N = P.shape[0]
    if ssim_value is not None and ssim_value < overall_ssim_threshold:
        return np.zeros(N, bool)
    f = np.float32
    c1, c2 = t1.astype(f).reshape(3), t2.astype(f).reshape(3)
    d1 = c2 - c1
    d2 = c1 - c2
    d1 = d1 / np.sqrt((d1 * d1).sum(dtype=f))
    d2 = d2 / np.sqrt((d2 * d2).sum(dtype=f))

    def line_dist(c, d):
        dots = ((P - c) @ d.reshape(3, 1)).astype(f)
        proj = c + d * dots
        return np.sqrt(((P - proj) ** 2).sum(axis=1, dtype=f))

    valid = (line_dist(c1, d1) < f(distance_threshold)) & (line_dist(c2, d2) < f(distance_threshold))
    cam = lambda c: np.sqrt(((P - c) ** 2).sum(axis=1, dtype=f))
    close = (cam(c1) < f(min_cam_distance)) | (cam(c2) < f(min_cam_distance))
    mean = P.astype(np.float64).mean(axis=0).astype(f)
    std = P.astype(np.float64).std(axis=0, ddof=1).astype(f) if N > 1 else np.full(3, np.nan, f)
    outlier = ~np.all(np.abs(P - mean) < f(sigma_threshold) * std, axis=1)
    return valid & (close | outlier)
