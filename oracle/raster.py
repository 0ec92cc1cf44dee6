"""ctypes front-end of the CPU rasterizer oracle (oracle/raster_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``splatco_b200/`` imports this.

PARITY UNPINNED (see raster_oracle.c header): the reference rasterizer source is in the missing
``submodules.zip``; this restates SURVEY.md Appendix A (reference call sites
gaussian_renderer/__init__.py:145-171, 208-242).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_raster.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_raster.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_num_rendered.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


@dataclass
class Projected:
    radii: np.ndarray
    xy: np.ndarray
    depths: np.ndarray
    cov3d: np.ndarray
    conic_opacity: np.ndarray
    rect: np.ndarray
    tiles_touched: np.ndarray


@dataclass
class Binned:
    R: int
    keys_unsorted: np.ndarray
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray


def grid_of(H, W):
    return (W + 15) // 16, (H + 15) // 16


def visible_filter(means3D, scales, rots, scale_mod, view, proj, tanfovx, tanfovy, H, W):
    """scales may be a strided [N,3] view of an [N,6] array, like the reference passes."""
    means3D = _f32(means3D)
    rots = _f32(rots)
    scales = np.asarray(scales, dtype=np.float32)
    if scales.strides[1] != 4:
        scales = np.ascontiguousarray(scales)
    stride = scales.strides[0] // 4
    N = means3D.shape[0]
    radii = np.zeros(N, np.int32)
    view, proj = _f32(view).reshape(-1), _f32(proj).reshape(-1)
    rc = lib().oracle_visible_filter(C.c_int(N), _p(means3D), _p(scales), C.c_int(stride), _p(rots),
                                     C.c_float(scale_mod), _p(view), _p(proj), C.c_float(tanfovx),
                                     C.c_float(tanfovy), C.c_int(H), C.c_int(W), _p(radii))
    assert rc == 0
    return radii


def preprocess(means3D, scales, rots, opacities, scale_mod, view, proj, tanfovx, tanfovy, H, W) -> Projected:
    means3D, scales, rots = _f32(means3D), _f32(scales), _f32(rots)
    opacities = _f32(opacities).reshape(-1)
    P = means3D.shape[0]
    view, proj = _f32(view).reshape(-1), _f32(proj).reshape(-1)
    out = Projected(np.zeros(P, np.int32), np.zeros((P, 2), np.float32), np.zeros(P, np.float32),
                    np.zeros((P, 6), np.float32), np.zeros((P, 4), np.float32),
                    np.zeros((P, 4), np.int32), np.zeros(P, np.uint32))
    rc = lib().oracle_preprocess(C.c_int(P), _p(means3D), _p(scales), C.c_int(3), _p(rots), _p(opacities),
                                 C.c_float(scale_mod), _p(view), _p(proj), C.c_float(tanfovx),
                                 C.c_float(tanfovy), C.c_int(H), C.c_int(W), _p(out.radii), _p(out.xy),
                                 _p(out.depths), _p(out.cov3d), _p(out.conic_opacity), _p(out.rect),
                                 _p(out.tiles_touched))
    assert rc == 0
    return out


def binning(pr: Projected, H, W) -> Binned:
    gx, gy = grid_of(H, W)
    P = pr.radii.shape[0]
    R = int(lib().oracle_num_rendered(C.c_int(P), _p(pr.tiles_touched)))
    ku = np.zeros(max(R, 1), np.uint64)
    ks = np.zeros(max(R, 1), np.uint64)
    vs = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.int32)
    rc = lib().oracle_binning(C.c_int(P), _p(pr.radii), _p(pr.rect), _p(pr.depths), _p(pr.tiles_touched),
                              C.c_int(gx), C.c_int(gy), C.c_int64(R), _p(ku), _p(ks), _p(vs), _p(ranges))
    assert rc == 0, rc
    return Binned(R, ku[:R], ks[:R], vs[:R], ranges)


def blend_fwd(pr: Projected, bn: Binned, colors, bg, H, W):
    colors, bg = _f32(colors), _f32(bg)
    out = np.zeros((3, H, W), np.float32)
    final_T = np.zeros((H, W), np.float32)
    n_contrib = np.zeros((H, W), np.int32)
    fragile = np.zeros((H, W), np.uint8)
    pl = bn.point_list if bn.R > 0 else np.zeros(1, np.uint32)
    rc = lib().oracle_blend_fwd(C.c_int(H), C.c_int(W), _p(bn.ranges), _p(pl), _p(pr.xy),
                                _p(pr.conic_opacity), _p(colors), _p(bg), _p(out), _p(final_T),
                                _p(n_contrib), _p(fragile))
    assert rc == 0
    return out, final_T, n_contrib, fragile.astype(bool)


def blend_bwd(pr: Projected, bn: Binned, colors, bg, final_T, n_contrib, dL_dpix, H, W):
    colors, bg, dL_dpix = _f32(colors), _f32(bg), _f32(dL_dpix)
    P = pr.radii.shape[0]
    g_mean2D = np.zeros((P, 3), np.float32)
    g_conic = np.zeros((P, 3), np.float32)
    g_opac = np.zeros((P, 1), np.float32)
    g_color = np.zeros((P, 3), np.float32)
    pl = bn.point_list if bn.R > 0 else np.zeros(1, np.uint32)
    rc = lib().oracle_blend_bwd(C.c_int(P), C.c_int(H), C.c_int(W), _p(bn.ranges), _p(pl), _p(pr.xy),
                                _p(pr.conic_opacity), _p(colors), _p(bg), _p(_f32(final_T)),
                                _p(np.ascontiguousarray(n_contrib, dtype=np.int32)), _p(dL_dpix),
                                _p(g_mean2D), _p(g_conic), _p(g_opac), _p(g_color))
    assert rc == 0
    return g_mean2D, g_conic, g_opac, g_color


def preprocess_bwd(means3D, scales, rots, scale_mod, view, proj, tanfovx, tanfovy, H, W, radii,
                   g_mean2D, g_conic):
    means3D, scales, rots = _f32(means3D), _f32(scales), _f32(rots)
    view, proj = _f32(view).reshape(-1), _f32(proj).reshape(-1)
    P = means3D.shape[0]
    gm = np.zeros((P, 3), np.float32)
    gs = np.zeros((P, 3), np.float32)
    gq = np.zeros((P, 4), np.float32)
    rc = lib().oracle_preprocess_bwd(C.c_int(P), _p(means3D), _p(scales), C.c_int(3), _p(rots),
                                     C.c_float(scale_mod), _p(view), _p(proj), C.c_float(tanfovx),
                                     C.c_float(tanfovy), C.c_int(H), C.c_int(W),
                                     _p(np.ascontiguousarray(radii, dtype=np.int32)), _p(_f32(g_mean2D)),
                                     _p(_f32(g_conic)), _p(gm), _p(gs), _p(gq))
    assert rc == 0
    return gm, gs, gq


def rasterize_forward(means3D, colors, opacities, scales, rots, scale_mod, view, proj, tanfovx,
                      tanfovy, H, W, bg):
    """Whole forward; returns a dict with every intermediate the parity tests compare."""
    pr = preprocess(means3D, scales, rots, opacities, scale_mod, view, proj, tanfovx, tanfovy, H, W)
    bn = binning(pr, H, W)
    img, final_T, n_contrib, fragile = blend_fwd(pr, bn, colors, bg, H, W)
    return dict(pr=pr, bn=bn, image=img, final_T=final_T, n_contrib=n_contrib, fragile=fragile)


def rasterize_backward(fw, means3D, colors, scales, rots, scale_mod, view, proj, tanfovx, tanfovy,
                       H, W, bg, dL_dpix):
    pr, bn = fw["pr"], fw["bn"]
    g_mean2D, g_conic, g_opac, g_color = blend_bwd(pr, bn, colors, bg, fw["final_T"], fw["n_contrib"],
                                                   dL_dpix, H, W)
    gm, gs, gq = preprocess_bwd(means3D, scales, rots, scale_mod, view, proj, tanfovx, tanfovy, H, W,
                                pr.radii, g_mean2D, g_conic)
    return dict(means3D=gm, means2D=g_mean2D, colors=g_color, opacities=g_opac, scales=gs,
                rotations=gq, conic=g_conic)
