"""Independent naive, tile-aware alpha-composite oracle in PyTorch (autograd-able, any dtype).

TEST INFRASTRUCTURE ONLY (see oracle/raster.py).  It exists to cross-check the C restatement
(oracle/raster_oracle.c): same semantics (SURVEY.md Appendix A / BASELINE.md §3.1 "naive tile-aware
PyTorch alpha-composite oracle"), written a second time, vectorised over pixels and sequential over
depth-sorted Gaussians, so that `torch.autograd` supplies the gradients the hand-derived backward in
the C oracle (and in the CUDA kernels) is compared against.

Reference call-site contract: gaussian_renderer/__init__.py:145-171.
"""
from __future__ import annotations

import torch


def project(means3D, scales, rots, scale_mod, view, proj, tanfovx, tanfovy, H, W):
    """Differentiable projection: returns (xy_pix [P,2], depth [P], conic [P,3], cov2d (a,b,c))."""
    dt = means3D.dtype
    view = view.to(dt).reshape(4, 4)      # row-vector convention: p_view = [p,1] @ view
    proj = proj.to(dt).reshape(4, 4)
    P = means3D.shape[0]
    ph = torch.cat([means3D, torch.ones(P, 1, dtype=dt)], dim=1)
    pv = ph @ view
    hom = ph @ proj
    pw = 1.0 / (hom[:, 3] + 0.0000001)
    ndc = hom[:, :2] * pw[:, None]
    s = scale_mod * scales
    r, x, y, z = rots.unbind(1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(P, 3, 3)
    M = R * s[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = pv[:, 2]
    txtz, tytz = pv[:, 0] / tz, pv[:, 1] / tz
    # reference quirk: where the clamp is active the clamped t.x is treated as a constant
    tx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * tz).detach(), pv[:, 0])
    ty = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * tz).detach(), pv[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz),
                     zero, fy / tz, -(fy * ty) / (tz * tz)], dim=1).reshape(P, 2, 3)
    Rot = view[:3, :3].t()                 # Rot[i][k] = view_flat[4k+i]
    A = J @ Rot
    cov = A @ Sigma @ A.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c = cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], dim=1)
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    return torch.stack([px, py], dim=1), tz, conic, (a, b, c)


def radii_from_cov(a, b, c):
    det = a * c - b * b
    mid = 0.5 * (a + c)
    sq = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    return torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + sq, mid - sq)))


def composite(xy, conic, opacities, colors, bg, rect, order, H, W):
    """Front-to-back blend.  rect [P,4] int (tile units, max exclusive; zero-area = invisible),
    order = Gaussian ids sorted by (fp32 depth bits, id).  Returns image [3,H,W]."""
    dt = xy.dtype
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    tx_, ty_ = xs // 16, ys // 16
    pxf, pyf = xs.to(dt), ys.to(dt)
    T = torch.ones(H, W, dtype=dt)
    C = torch.zeros(3, H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    opac = opacities.reshape(-1)
    for g in order.tolist():
        r0x, r0y, r1x, r1y = rect[g].tolist()
        if (r1x - r0x) * (r1y - r0y) == 0:
            continue
        inside = (tx_ >= r0x) & (tx_ < r1x) & (ty_ >= r0y) & (ty_ < r1y) & ~done
        if not bool(inside.any()):
            continue
        dx, dy = xy[g, 0] - pxf, xy[g, 1] - pyf
        power = -0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) - conic[g, 1] * dx * dy
        ea = opac[g] * torch.exp(power)
        alpha = ea + (torch.clamp(ea, max=0.99) - ea).detach()   # reference keeps the gradient when clamped
        live = inside & (power <= 0) & (alpha >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = live & (test_T < 0.0001)
        done = done | stop
        live = live & ~stop
        w = torch.where(live, alpha * T, torch.zeros_like(T))
        C = C + colors[g][:, None, None] * w[None]
        T = torch.where(live, test_T, T)
    return C + T[None] * bg.to(dt)[:, None, None]
