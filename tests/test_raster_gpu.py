"""GPU parity tests of the rasterizer path (preprocess -> binning -> blend fwd/bwd), through the
C ABI, against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): radii, tile keys, sort order, tile ranges bit-exact; image within
1e-4 max abs; gradients within 1e-3 relative (floor 1e-3*max|g| per tensor, atomics reorder sums).
Pixels the oracle flags `fragile` (some evaluated pair within 1e-4 relative of the alpha=1/255 or
T=1e-4 cut, where a 1-ulp exp() difference legitimately moves the pixel by up to ~0.4 %) are held to
1e-2 instead and must stay a tiny share of the image.
"""
import numpy as np
import pytest
import torch

from tests.util import chunk, layout, oracle_backward, oracle_forward, rel_err, scene, settings_for, tans

pytestmark = pytest.mark.gpu


def gpu_forward(cam, means, colors, opac, scales, rots, bg, scale_mod=1.0, debug=False):
    from splatco_b200.diff_gaussian_rasterization import rasterize_forward_state
    st = settings_for(cam, bg, scale_mod=scale_mod, debug=debug)
    d = "cuda"
    color, radii, state = rasterize_forward_state(means.to(d), colors.to(d), opac.to(d), scales.to(d), rots.to(d), st)
    torch.cuda.synchronize()
    return color, radii, state, st


def unpack_state(state):
    from splatco_b200 import _lib
    P, R, H, W = state.P, state.R, state.H, state.W
    out = {}
    go = layout("geom", P)
    rec = chunk(state.geom, go[0], torch.float32, 12 * P).view(P, 12).cpu().numpy()
    out["xy"] = rec[:, 0:2]
    out["conic_opacity"] = rec[:, [2, 3, 4, 5]]
    out["depth"] = rec[:, 9]
    out["depths"] = chunk(state.geom, go[1], torch.float32, P).cpu().numpy()
    out["tiles"] = chunk(state.geom, go[2], torch.int32, P).cpu().numpy().astype(np.uint32)
    if R > 0:
        bo = layout("binning", R)
        s = _lib.lib().splatco_sorted_buffer_index(H, W)
        out["keys"] = chunk(state.binning, bo[0 + s], torch.int64, R).cpu().numpy().view(np.uint64)
        out["point_list"] = chunk(state.binning, bo[2 + s], torch.int32, R).cpu().numpy().view(np.uint32)
    io = layout("image", H, W)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = chunk(state.image, io[0], torch.int32, 2 * T).view(T, 2).cpu().numpy()
    out["final_T"] = chunk(state.image, io[1], torch.float32, H * W).view(H, W).cpu().numpy()
    out["n_contrib"] = chunk(state.image, io[2], torch.int32, H * W).view(H, W).cpu().numpy()
    return out


def assert_forward_parity(fw, color, radii, state, strict_float=True):
    pr, bn = fw["pr"], fw["bn"]
    g = unpack_state(state)
    assert np.array_equal(radii.cpu().numpy(), pr.radii), "radii not bit-exact"
    assert np.array_equal(g["tiles"], pr.tiles_touched), "tiles_touched not bit-exact"
    assert state.R == bn.R, "num_rendered differs"
    vis = pr.radii > 0
    # pinned-rounding chain: the projected floats are bit-identical too
    assert np.array_equal(g["depths"][vis].view(np.uint32), pr.depths[vis].view(np.uint32)), "depth bits differ"
    assert np.array_equal(g["xy"][vis].view(np.uint32), pr.xy[vis].view(np.uint32)), "xy bits differ"
    assert np.array_equal(g["conic_opacity"][vis].view(np.uint32), pr.conic_opacity[vis].view(np.uint32))
    if bn.R > 0:
        assert np.array_equal(g["keys"], bn.keys), "sorted tile|depth keys differ"
        assert np.array_equal(g["point_list"], bn.point_list), "sort permutation differs"
    assert np.array_equal(g["ranges"], bn.ranges), "tile ranges differ"
    img = color.cpu().numpy()
    frag = fw["fragile"]
    err = np.abs(img - fw["image"])
    assert frag.mean() < 0.02
    assert err[:, ~frag].max() <= 1e-4, f"image max abs err {err[:, ~frag].max()}"
    assert err.max() <= 1e-2
    assert np.array_equal(g["n_contrib"][~frag], fw["n_contrib"][~frag])
    assert np.abs(g["final_T"][~frag] - fw["final_T"][~frag]).max() <= 1e-5
    return g


@pytest.mark.parametrize("seed,W,H,M,sig", [
    (1, 64, 48, 300, (1.0, 6.0)),        # tiny
    (2, 250, 130, 5000, (0.5, 4.0)),     # ragged: W,H not multiples of 16
    (3, 256, 256, 100000, (0.5, 4.0)),   # BASELINE configs[0] shape (~100k Gaussians, 256x256): 5 sort passes
    (4, 980, 545, 60000, (0.5, 8.0)),    # C2 resolution (6 sort passes), R spans many sort CTAs
    (5, 97, 33, 2000, (4.0, 40.0)),      # large splats: long per-tile lists, multi-batch tiles, early termination
])
def test_forward_parity(seed, W, H, M, sig):
    cam, means, colors, opac, scales, rots = scene(M, W, H, seed, sigma_px=sig)
    bg = [1.0, 1.0, 1.0] if seed % 2 else [0.1, 0.2, 0.3]
    fw = oracle_forward(cam, means, colors, opac, scales, rots, bg)
    color, radii, state, _ = gpu_forward(cam, means, colors, opac, scales, rots, bg)
    assert_forward_parity(fw, color, radii, state)


def test_binning_long_tiles_and_radix_composition_agree():
    """Tiles with more instances than the shared-memory sort holds (8192) take the in-place global path;
    the tile-segmented binning and the literal upstream composition (duplicateWithKeys + radix sort +
    identifyTileRanges) must both reproduce the oracle's keys / permutation / ranges bit for bit."""
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    W, H, M = 64, 48, 70000
    cam, means, colors, opac, scales, rots = scene(M, W, H, 77, sigma_px=(0.5, 3.0))
    fw = oracle_forward(cam, means, colors, opac, scales, rots, [1.0, 1.0, 1.0])
    bn = fw["bn"]
    counts = bn.ranges[:, 1] - bn.ranges[:, 0]
    assert counts.max() > 8192, f"scene does not exercise the long-tile path (max {counts.max()})"
    color, radii, state, _ = gpu_forward(cam, means, colors, opac, scales, rots, [1.0, 1.0, 1.0])
    g = assert_forward_parity(fw, color, radii, state)
    # the radix composition on fresh workspaces
    L = _lib.lib()
    binning = torch.zeros_like(state.binning)
    image = torch.zeros_like(state.image)
    check(L.splatco_binning_radix(state.P, state.R, H, W, ptr(state.radii_full), ptr(state.geom), ptr(binning), ptr(image),
                                  torch.cuda.current_stream().cuda_stream), "splatco_binning_radix")
    torch.cuda.synchronize()
    bo = layout("binning", state.R)
    sidx = L.splatco_sorted_buffer_index(H, W)
    keys = chunk(binning, bo[0 + sidx], torch.int64, state.R).cpu().numpy().view(np.uint64)
    plist = chunk(binning, bo[2 + sidx], torch.int32, state.R).cpu().numpy().view(np.uint32)
    io = layout("image", H, W)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = chunk(image, io[0], torch.int32, 2 * T).view(T, 2).cpu().numpy()
    assert np.array_equal(keys, g["keys"]) and np.array_equal(plist, g["point_list"]) and np.array_equal(ranges, g["ranges"])


def test_forward_parity_scale_modifier_and_debug():
    cam, means, colors, opac, scales, rots = scene(3000, 160, 120, 21)
    fw = oracle_forward(cam, means, colors, opac, scales, rots, [0, 0, 0], scale_mod=1.7)
    color, radii, state, _ = gpu_forward(cam, means, colors, opac, scales, rots, [0, 0, 0], scale_mod=1.7, debug=True)
    assert_forward_parity(fw, color, radii, state)


def test_visible_filter_bit_exact_strided():
    from oracle import raster as R
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizer
    cam, means, colors, opac, scales, rots = scene(200000, 980, 545, 31, sigma_px=(0.5, 6.0))
    s6 = torch.cat([scales, torch.rand(scales.shape[0], 3)], dim=1)
    tx, ty = tans(cam)
    want = R.visible_filter(means.numpy(), s6.numpy()[:, :3], rots.numpy(), 1.0, cam.world_view_transform.numpy(),
                            cam.full_proj_transform.numpy(), tx, ty, 545, 980)
    rast = GaussianRasterizer(settings_for(cam, [1, 1, 1]))
    s6d = s6.cuda()
    got = rast.visible_filter(means3D=means.cuda(), scales=s6d[:, :3], rotations=rots.cuda(), cov3D_precomp=None)
    assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want)
    assert 0 < (want > 0).sum() < want.size
    # the compacting variant prefilter_voxel uses: same decisions, plus the ascending index list
    from splatco_b200.diff_gaussian_rasterization import take_compaction, visible_mask_compact
    mask = visible_mask_compact(means.cuda(), s6d[:, :3], rots.cuda(), rast.raster_settings)
    assert mask.dtype == torch.bool and np.array_equal(mask.cpu().numpy(), want > 0)
    idx, V = take_compaction(mask)
    assert V == int((want > 0).sum()) and np.array_equal(idx.cpu().numpy(), np.nonzero(want > 0)[0].astype(np.int32))
    assert take_compaction(mask) is None                      # single use
    mask2 = visible_mask_compact(means.cuda(), s6d[:, :3], rots.cuda(), rast.raster_settings)
    mask2[0] = ~mask2[0]                                      # a modified mask must not be trusted
    assert take_compaction(mask2) is None


@pytest.mark.parametrize("seed,W,H,M,sig", [(41, 64, 48, 300, (1.0, 6.0)), (42, 250, 130, 5000, (0.5, 4.0)),
                                           (43, 97, 33, 2000, (4.0, 40.0)), (44, 256, 256, 100000, (0.5, 4.0))])
def test_backward_parity(seed, W, H, M, sig):
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizer
    cam, means, colors, opac, scales, rots = scene(M, W, H, seed, sigma_px=sig)
    bg = [1.0, 1.0, 1.0]
    fw = oracle_forward(cam, means, colors, opac, scales, rots, bg)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(seed))
    d = "cuda"
    m, c, o, s, q = [t.to(d).requires_grad_() for t in (means, colors, opac, scales, rots)]
    m2d = torch.zeros_like(m, requires_grad=True)
    rast = GaussianRasterizer(settings_for(cam, bg))
    img, radii = rast(means3D=m, means2D=m2d, shs=None, colors_precomp=c, opacities=o, scales=s, rotations=q,
                      cov3D_precomp=None)
    # L1 loss against a random target, gradient evaluated at the ORACLE image so both sides see the same dL/dpix
    dL = (torch.sign(torch.from_numpy(fw["image"]) - gt) / (3 * H * W)).float()
    img.backward(dL.to(d))
    bw = oracle_backward(fw, cam, means, colors, scales, rots, bg, dL.numpy())
    tol = 1e-3
    assert rel_err(m2d.grad.cpu().numpy(), bw["means2D"]) < tol
    assert rel_err(c.grad.cpu().numpy(), bw["colors"]) < tol
    assert rel_err(o.grad.cpu().numpy(), bw["opacities"]) < tol
    assert rel_err(m.grad.cpu().numpy(), bw["means3D"]) < tol
    assert rel_err(s.grad.cpu().numpy(), bw["scales"]) < tol
    assert rel_err(q.grad.cpu().numpy(), bw["rotations"]) < tol
    assert radii.dtype == torch.int32 and not radii.requires_grad


def test_empty_and_all_culled():
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizer
    cam, means, colors, opac, scales, rots = scene(64, 50, 34, 5)
    bg = [0.2, 0.4, 0.6]
    rast = GaussianRasterizer(settings_for(cam, bg))
    z = lambda *s: torch.zeros(*s, device="cuda")
    img, radii = rast(means3D=z(0, 3), means2D=z(0, 3), shs=None, colors_precomp=z(0, 3), opacities=z(0, 1),
                      scales=z(0, 3), rotations=z(0, 4), cov3D_precomp=None)
    assert radii.numel() == 0 and torch.allclose(img, torch.tensor(bg, device="cuda")[:, None, None].expand(3, 34, 50))
    behind = (cam.camera_center * 3.0)[None].repeat(64, 1)      # further out than the camera, behind it
    m = behind.cuda().requires_grad_()
    img, radii = rast(means3D=m, means2D=torch.zeros_like(m), shs=None, colors_precomp=colors.cuda(),
                      opacities=opac.cuda(), scales=scales.cuda(), rotations=rots.cuda(), cov3D_precomp=None)
    assert (radii == 0).all() and torch.allclose(img, torch.tensor(bg, device="cuda")[:, None, None].expand(3, 34, 50))
    img.sum().backward()
    assert torch.count_nonzero(m.grad) == 0


def test_full_size_properties_c2():
    """BASELINE configs[1] shape: ~1M Gaussians at 980x545.  Size-independent properties instead of the
    oracle: key sortedness, stable ties, ranges partition the list, tile counts sum to R, the blend is a
    convex combination (image within [0,1] for colours in [0,1] and bg in [0,1]), and linearity of the
    backward in dL/dpix."""
    from splatco_b200.diff_gaussian_rasterization import rasterize_backward_state
    W, H, M = 980, 545, 1_000_000
    cam, means, colors, opac, scales, rots = scene(M, W, H, 77, sigma_px=(0.5, 4.0))
    color, radii, state, st = gpu_forward(cam, means, colors, opac, scales, rots, [1, 1, 1])
    g = unpack_state(state)
    R = state.R
    assert R == int(g["tiles"].astype(np.int64).sum()) and R > M
    keys, pl = g["keys"], g["point_list"]
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert np.all(pl[1:][same] > pl[:-1][same])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    counts = np.bincount(tiles, minlength=g["ranges"].shape[0])
    assert np.array_equal(g["ranges"][:, 1] - g["ranges"][:, 0], counts)
    nz = counts > 0
    assert np.array_equal(g["ranges"][nz, 0], (np.cumsum(counts) - counts)[nz])
    # depth bits in the key equal the Gaussian's depth
    assert np.array_equal((keys & np.uint64(0xffffffff)).astype(np.uint32), g["depths"].view(np.uint32)[pl])
    img = color.cpu().numpy()
    assert img.min() >= -1e-6 and img.max() <= 1.0 + 1e-5
    assert np.all(g["final_T"] <= 1.0) and np.all(g["final_T"] >= 0.0)
    # linearity of backward in dL/dpix: bwd(a*g1 + g2) == a*bwd(g1) + bwd(g2)
    gen = torch.Generator(device="cuda").manual_seed(5)
    g1 = torch.randn(3, H, W, device="cuda", generator=gen) * 1e-3
    g2 = torch.randn(3, H, W, device="cuda", generator=gen) * 1e-3
    d = "cuda"
    m, s, q = means.to(d), scales.to(d), rots.to(d)
    b1 = rasterize_backward_state(state, g1, m, s, q, st)
    b2 = rasterize_backward_state(state, g2, m, s, q, st)
    b3 = rasterize_backward_state(state, 2.0 * g1 + g2, m, s, q, st)
    for k in ("means3D", "colors", "opacities", "scales", "rotations", "means2D"):
        want = (2.0 * b1[k] + b2[k]).cpu().numpy()
        assert rel_err(b3[k].cpu().numpy(), want) < 2e-3, k


def test_large_image_binning_paths_agree():
    """3840x2160 (32,400 tiles: the render-only sweep shape of BASELINE configs[4]) with 1.5 M Gaussians: the
    tile-segmented binning and the literal radix composition must produce identical sorted keys, permutation and
    ranges; the list must be sorted with stable ties and partition by tile."""
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    W, H, M = 3840, 2160, 1_500_000
    cam, means, colors, opac, scales, rots = scene(M, W, H, 91, sigma_px=(0.5, 4.0))
    color, radii, state, _ = gpu_forward(cam, means, colors, opac, scales, rots, [0.0, 0.0, 0.0])
    g = unpack_state(state)
    R = state.R
    assert R == int(g["tiles"].astype(np.int64).sum()) and R > M
    keys, pl = g["keys"], g["point_list"]
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert np.all(pl[1:][same] > pl[:-1][same])
    counts = np.bincount((keys >> np.uint64(32)).astype(np.int64), minlength=g["ranges"].shape[0])
    assert np.array_equal(g["ranges"][:, 1] - g["ranges"][:, 0], counts)
    L = _lib.lib()
    binning = torch.zeros_like(state.binning)
    image = torch.zeros_like(state.image)
    check(L.splatco_binning_radix(state.P, R, H, W, ptr(state.radii_full), ptr(state.geom), ptr(binning), ptr(image),
                                  torch.cuda.current_stream().cuda_stream), "splatco_binning_radix")
    torch.cuda.synchronize()
    bo = layout("binning", R)
    sidx = L.splatco_sorted_buffer_index(H, W)
    assert np.array_equal(chunk(binning, bo[0 + sidx], torch.int64, R).cpu().numpy().view(np.uint64), keys)
    assert np.array_equal(chunk(binning, bo[2 + sidx], torch.int32, R).cpu().numpy().view(np.uint32), pl)
    io = layout("image", H, W)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    assert np.array_equal(chunk(image, io[0], torch.int32, 2 * T).view(T, 2).cpu().numpy(), g["ranges"])
    img = color.cpu().numpy()
    assert np.isfinite(img).all() and img.min() >= -1e-6 and img.max() <= 1.0 + 1e-5


def test_upstream_structure_comparator_matches_product_and_oracle():
    """bench.py's same-GPU comparator (csrc/blend_upstream.cu: upstream's blend structure restated -- 256-instance batches,
    per-pixel atomics) must compute what the product kernels compute: same image / n_contrib as the oracle, same
    per-Gaussian blend gradients as splatco_blend_bwd."""
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    W, H, M = 300, 200, 30_000
    cam, means, colors, opac, scales, rots = scene(M, W, H, 19)
    bg = [0.3, 0.5, 0.7]
    fw = oracle_forward(cam, means, colors, opac, scales, rots, bg)
    color, radii, state, st = gpu_forward(cam, means, colors, opac, scales, rots, bg)
    L = _lib.lib()
    stream = _lib.raw_stream(color.device)
    bgd = torch.tensor(bg, device="cuda")
    n_prod = unpack_state(state)["n_contrib"].copy()
    img_up = torch.empty_like(color)
    check(L.splatco_blend_fwd_upstream(state.RL, H, W, ptr(bgd), ptr(state.geom), ptr(state.binning), ptr(state.image), ptr(img_up), stream),
          "splatco_blend_fwd_upstream")
    torch.cuda.synchronize()
    err = np.abs(img_up.cpu().numpy() - fw["image"])
    assert err[:, ~fw["fragile"]].max() <= 1e-4
    g_up = unpack_state(state)
    assert np.array_equal(g_up["n_contrib"][~fw["fragile"]], n_prod[~fw["fragile"]])
    dL = torch.randn(3, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) * 1e-3
    P = state.P

    def run(fn):
        z = lambda c: torch.zeros(P, c, device="cuda")
        g2d, gcon, gop, gcol = z(3), z(3), z(1), z(3)
        check(fn(P, state.RL, H, W, ptr(bgd), ptr(state.geom), ptr(state.binning), ptr(state.image), ptr(dL), ptr(g2d), ptr(gcon),
                 ptr(gop), ptr(gcol), stream), "blend_bwd")
        torch.cuda.synchronize()
        return [t.cpu().numpy() for t in (g2d, gcon, gop, gcol)]

    up = run(L.splatco_blend_bwd_upstream)
    # (the forward above re-wrote final_T / n_contrib with the comparator's values: both backward kernels replay from them)
    prod = run(L.splatco_blend_bwd)
    for name, a, b in zip(("mean2D", "conic", "opacity", "colour"), up, prod):
        assert rel_err(a, b) < 2e-3, name


@pytest.mark.parametrize("W,H,M,sigma", [(250, 131, 40_000, (0.3, 12.0)), (64, 48, 3_000, (4.0, 60.0)), (333, 200, 150_000, (0.2, 1.5))])
def test_blend_implementations_agree_on_hard_splat_mixes(W, H, M, sigma):
    """The round-2 blend kernels against the round-1 ones on the SAME binned state: the per-pixel candidate lists of
    blend_fwd2 are only a filter in front of an unchanged evaluation, so image, final_T and n_contrib must be identical
    bit for bit -- on image sizes that are not multiples of the tile, splats from sub-pixel to screen-filling, needle-shaped
    conics and opacities straddling 1/255; the matrix-form reduction of blend_bwd2 must reproduce the nine sums of the
    transposition-buffer kernel to rounding."""
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    cam, means, colors, opac, scales, rots = scene(M, W, H, 17, sigma_px=sigma)
    g = torch.Generator().manual_seed(5)
    scales = scales * torch.exp(1.5 * torch.randn(scales.shape, generator=g))          # strongly anisotropic (needles)
    opac = opac.clone()
    opac[::7] = 1.0 / 255.0 * (0.5 + torch.rand(opac[::7].shape, generator=g))           # around the alpha cut
    opac[3::11] = 0.999
    color, radii, state, st = gpu_forward(cam, means, colors, opac, scales, rots, [0.1, 0.3, 0.9])
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    bg = torch.tensor([0.1, 0.3, 0.9], device="cuda")
    P, R = state.P, state.RL
    assert R > 0
    try:
        outs = []
        for impl in (1, 2):
            check(L.splatco_blend_set_impl(impl, 0), "set_impl")
            img = torch.empty(3, H, W, device="cuda")
            check(L.splatco_blend_fwd(R, H, W, ptr(bg), ptr(state.geom), ptr(state.binning), ptr(state.image), ptr(img), stream), "blend_fwd")
            torch.cuda.synchronize()
            u = unpack_state(state)
            outs.append((img.clone(), u["final_T"].copy(), u["n_contrib"].copy()))
        assert torch.equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32)) and np.array_equal(outs[0][2], outs[1][2])
        dL = torch.randn(3, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) / (3 * H * W)
        grads = []
        for impl in (1, 2, 3):
            check(L.splatco_blend_set_impl(0, impl), "set_impl")
            gs = [torch.zeros(P, c, device="cuda") for c in (3, 3, 1, 3)]
            check(L.splatco_blend_bwd(P, R, H, W, ptr(bg), ptr(state.geom), ptr(state.binning), ptr(state.image), ptr(dL),
                                      *[ptr(t) for t in gs], stream), "blend_bwd")
            torch.cuda.synchronize()
            grads.append([t.double().cpu().numpy() for t in gs])
        for impl_g in grads[1:]:
            for name, a, b in zip(("mean2D", "conic", "opacity", "color"), impl_g, grads[0]):
                assert rel_err(a, b) < 5e-4, (name, rel_err(a, b))      # fp32 sums in different orders (atomics)
    finally:
        check(L.splatco_blend_set_impl(2, 2), "set_impl")


@pytest.mark.parametrize("quantum", [0.25, 2.0 ** -18])
def test_binning_with_massive_depth_collisions(quantum):
    """Depths quantised in view space: with quantum 0.25 thousands of Gaussians of a tile share the SAME depth (ties
    resolved by id), with 2^-18 they agree in the top 24 key bits and differ in the low byte -- the per-tile sort leaves
    that byte to its run-fixing step and must fall back to all four passes when the runs are long.  Keys, permutation and
    ranges must still equal the oracle's and the literal radix composition's bit for bit."""
    from splatco_b200 import _lib
    from splatco_b200._lib import check, ptr
    W, H, M = 320, 200, 60_000
    cam, means, colors, opac, scales, rots = scene(M, W, H, 23, sigma_px=(0.5, 3.0))
    Wv = cam.world_view_transform.double()
    pv = torch.cat([means.double(), torch.ones(M, 1, dtype=torch.float64)], dim=1) @ Wv
    g = torch.Generator().manual_seed(2)
    zq = torch.round(pv[:, 2] / 0.25) * 0.25
    if quantum < 0.25:
        zq = zq + quantum * torch.randint(0, 40, (M,), generator=g).double()      # 40 values inside one 24-bit bucket
    pv[:, 2] = zq
    means = (pv @ torch.linalg.inv(Wv))[:, :3].float().contiguous()
    fw = oracle_forward(cam, means, colors, opac, scales, rots, [0.0, 0.0, 0.0])
    color, radii, state, _ = gpu_forward(cam, means, colors, opac, scales, rots, [0.0, 0.0, 0.0])
    g_ = unpack_state(state)
    bn = fw["bn"]
    assert state.R == bn.R
    assert np.array_equal(g_["keys"], bn.keys), "sorted tile|depth keys differ"
    assert np.array_equal(g_["point_list"], bn.point_list), "sort permutation differs"
    assert np.array_equal(g_["ranges"], bn.ranges)
    keys = g_["keys"]
    same = keys[1:] == keys[:-1]
    if quantum == 0.25:
        assert same.mean() > 0.5          # the case really is dominated by ties
    else:
        top = (keys >> np.uint64(8))
        assert (top[1:] == top[:-1]).mean() > 0.5 and same.mean() < 0.5       # long runs that are NOT plain ties
