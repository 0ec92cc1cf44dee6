"""Seeded synthetic cameras and scenes of the BASELINE.json shapes.

Camera matrices follow the reference's conventions exactly: `world_view_transform` and
`full_proj_transform` are the *transposed* (row-vector) matrices built from `getWorld2View2` and
`getProjectionMatrix` (reference utils/graphics_utils.py:38-71, scene/cameras.py:48-58), so the
rasterizer reads them column-major just like it does for the reference's `Camera`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch


def world2view(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """W2C 4x4 with the reference's (R stored transposed, t as given) convention
    (utils/graphics_utils.py:38-51 with translate=0, scale=1)."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt.astype(np.float32)


def projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    """utils/graphics_utils.py:53-71."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class SynthCamera:
    """Duck-type of the fields gaussian_renderer reads from the reference's Camera
    (gaussian_renderer/__init__.py:34,57,142-155)."""
    uid: int
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor
    full_proj_transform: torch.Tensor
    camera_center: torch.Tensor
    R: np.ndarray = field(default=None, repr=False)
    T: np.ndarray = field(default=None, repr=False)

    def to(self, device):
        return SynthCamera(self.uid, self.image_width, self.image_height, self.FoVx, self.FoVy,
                           self.world_view_transform.to(device), self.full_proj_transform.to(device),
                           self.camera_center.to(device), self.R, self.T)


def look_at_camera(uid, eye, target, W, H, fovx, znear=0.01, zfar=100.0) -> SynthCamera:
    eye = np.asarray(eye, np.float64)
    target = np.asarray(target, np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    # camera axes (x right, y down, z forward) as rows of the world->camera rotation
    Rwc = np.stack([right, down, fwd], axis=0)
    t = -Rwc @ eye
    R = Rwc.transpose()                     # reference stores R transposed (COLMAP loader convention)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    wv = torch.tensor(world2view(R, t)).transpose(0, 1).contiguous()
    pm = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(pm.unsqueeze(0))).squeeze(0).contiguous()
    center = wv.inverse()[3, :3].contiguous()
    return SynthCamera(uid, W, H, fovx, fovy, wv, full, center, R, t)


def ring_cameras(n, W, H, radius=3.0, height=0.5, fovx=2.0 * math.atan(0.5 / 1.2), phase=0.0):
    cams = []
    for i in range(n):
        a = phase + 2.0 * math.pi * i / max(n, 1)
        eye = (radius * math.cos(a), radius * math.sin(a), height)
        cams.append(look_at_camera(i, eye, (0.0, 0.0, 0.0), W, H, fovx))
    return cams


def random_gaussians(M, cam: SynthCamera, seed, sigma_px=(0.5, 4.0), depth=(2.0, 6.0), frac_behind=0.02):
    """C5-style direct Gaussian cloud: means uniform in the frustum (a small share behind the
    camera / off-screen so culling is exercised), log-uniform scales giving sigma_px in the given
    range, opacity U(0.05,0.9), colour U(0,1), random unit quaternions.  CPU fp32 tensors."""
    g = torch.Generator().manual_seed(seed)
    W, H = cam.image_width, cam.image_height
    tanx, tany = math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2)
    z = torch.empty(M).uniform_(depth[0], depth[1], generator=g)
    u = torch.empty(M).uniform_(-1.15, 1.15, generator=g)
    v = torch.empty(M).uniform_(-1.15, 1.15, generator=g)
    nb = int(M * frac_behind)
    if nb > 0:
        z[:nb] = torch.empty(nb).uniform_(-1.0, 0.25, generator=g)
    pv = torch.stack([u * tanx * z, v * tany * z, z, torch.ones(M)], dim=1)
    means = (pv @ cam.world_view_transform.inverse())[:, :3].contiguous()
    focal = W / (2 * tanx)
    lo, hi = math.log(sigma_px[0]), math.log(sigma_px[1])
    spx = torch.exp(torch.empty(M, 3).uniform_(lo, hi, generator=g))
    scales = (spx * z.abs().clamp_min(0.3)[:, None] / focal).contiguous()
    q = torch.randn(M, 4, generator=g)
    rots = (q / q.norm(dim=1, keepdim=True)).contiguous()
    opac = torch.empty(M, 1).uniform_(0.05, 0.9, generator=g)
    colors = torch.rand(M, 3, generator=g)
    return means, colors, opac, scales, rots
