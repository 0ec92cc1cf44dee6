"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/splatco_b200.h declares; sizing/layout helpers (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from splatco_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "splatco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(splatco_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/splatco_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.splatco_abi_version() == 1
    assert isinstance(L.splatco_last_error(), bytes)


def test_workspace_sizing_monotone_and_aligned():
    L = _lib.lib()
    prev = 0
    for P in (0, 1, 255, 256, 257, 100000):
        b = L.splatco_geom_bytes(P)
        assert b % 256 == 0 and b >= prev
        prev = b
    assert L.splatco_geom_bytes(1000) >= 1000 * (48 + 4 + 4)
    assert L.splatco_binning_bytes(1000) >= 1000 * 24
    assert L.splatco_image_bytes(545, 980) >= 545 * 980 * 8 + 62 * 35 * 8
    offs = (C.c_size_t * 16)()
    assert L.splatco_geom_layout(1000, offs, 8) == 6 and list(offs)[:6] == sorted(list(offs)[:6])
    assert L.splatco_binning_layout(5000, offs, 8) == 6
    assert L.splatco_image_layout(545, 980, offs, 16) == 8


def test_sort_pass_parity_matches_key_width():
    L = _lib.lib()
    # 256x256 -> 256 tiles -> 8 tile bits -> 40 key bits -> 5 passes -> sorted data in buffer 1
    assert L.splatco_sorted_buffer_index(256, 256) == 1
    # 980x545 -> 62*35=2170 tiles -> 12 bits -> 44 bits -> 6 passes -> buffer 0
    assert L.splatco_sorted_buffer_index(545, 980) == 0


def test_argument_validation_without_gpu():
    L = _lib.lib()
    rc = L.splatco_visible_filter(-1, None, None, 3, None, 1.0, None, None, 1.0, 1.0, 16, 16, None, None)
    assert rc < 0 and b"bad sizes" in L.splatco_last_error()
    rc = L.splatco_visible_filter(4, None, None, 3, None, 1.0, None, None, 1.0, 1.0, 16, 16, None, None)
    assert rc < 0 and b"null pointer" in L.splatco_last_error()


def test_rasterizer_refuses_cpu_tensors():
    from splatco_b200.diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 1,
                                      torch.zeros(3), False, False)
    r = GaussianRasterizer(s)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r(means3D=torch.zeros(4, 3), means2D=torch.zeros(4, 3), shs=None, colors_precomp=torch.zeros(4, 3),
          opacities=torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4), cov3D_precomp=None)
    with pytest.raises(Exception, match="excatly one"):
        r(means3D=torch.zeros(4, 3), means2D=torch.zeros(4, 3), shs=None, colors_precomp=None,
          opacities=torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4), cov3D_precomp=None)


def test_argument_validation_of_the_training_loop_entry_points():
    """Bad sizes / null pointers are refused with a message before any device work (so this runs without a GPU)."""
    L = _lib.lib()
    err = lambda: L.splatco_last_error().decode()
    assert L.splatco_tv_add_grad(0, 4, 4, None, None, 1.0, None) < 0 and "tv_add_grad" in err()
    assert L.splatco_tv_add_grad(5, 4, 4, None, None, 1.0, None) < 0 and "null pointer" in err()
    assert L.splatco_scaling_reg_fwd(0, None, None, None, None) < 0 and "at least one row" in err()
    assert L.splatco_mv_consistency_fwd(1, 3, None, None, None, None, None, 0.6, None, None, None) < 0 and "views supported" in err()
    assert L.splatco_mv_consistency_fwd(9, 3, None, None, None, None, None, 0.6, None, None, None) < 0 and "views supported" in err()
    assert L.splatco_grow_count(-1, None, None, 0.0, None, None, 0.0, None, None) < 0 and "bad slot count" in err()
    assert L.splatco_grow_count(10, None, None, 0.0, None, None, 0.0, None, None) < 0 and "cand_mask" in err()
    assert L.splatco_grow_emit(10, 32, 0.1, 5, 6, None, None, None, None, None) < 0 and "bad sizes" in err()
    assert L.splatco_cvpm_mask(-1, None, None, None, None, 0.6, 0.01, 3.0, 0.5, None, None, None, None) < 0 and "bad N" in err()
    assert L.splatco_adam_step(-1, None, 0.9, 0.999, 1e-15, None) < 0 and "bad tensor list" in err()
    assert L.splatco_adam_step(0, None, 0.9, 0.999, 1e-15, None) == 0            # nothing to update is not an error
    assert L.splatco_mv_consistency_ws_bytes(4) >= 6 * 8 + 6 * 4 and L.splatco_grow_ws_bytes(1000) >= 1000 * (4 + 12 + 4 + 4 + 1 + 24)


def test_round2_switches_and_count_pointer():
    """The A/B switch of the blend kernels validates its arguments; the device-count pointer of the prefilter's
    compaction (handed to splatco_decode_desc::V_dev) sits in the last chunk of its workspace."""
    L = _lib.lib()
    err = lambda: L.splatco_last_error().decode()
    assert L.splatco_blend_set_impl(0, 0) == 0
    assert L.splatco_blend_set_impl(3, 0) < 0 and "blend_set_impl" in err()
    assert L.splatco_blend_set_impl(0, 9) < 0 and "blend_set_impl" in err()
    assert L.splatco_blend_set_impl(2, 2) == 0                      # the defaults
    for N in (1, 256, 257, 100_000):
        ws_bytes = L.splatco_visible_compact_ws_bytes(N)
        base = 1 << 20
        p = L.splatco_visible_compact_count_ptr(C.c_void_p(base), N)
        assert p is not None and base < p <= base + ws_bytes - 4 and (p - base) % 256 == 0, (N, p, ws_bytes)
    assert not L.splatco_visible_compact_count_ptr(None, 10)


def test_decode_descriptor_mirror_matches_the_header():
    """ctypes mirror of splatco_decode_desc: the round-2 fields (V_dev, V_layout) exist and the struct size is what the
    C compiler lays out for the header (checked by compiling a one-line probe with gcc)."""
    import subprocess
    import tempfile
    from splatco_b200.decode import DecodeDesc, DecodeGrads
    names = [f[0] for f in DecodeDesc._fields_]
    assert names[-2:] == ["V_dev", "V_layout"] and names[1] == "V"
    src = '#include <stdio.h>\n#include "splatco_b200.h"\nint main(void){printf("%zu %zu\\n", sizeof(splatco_decode_desc), sizeof(splatco_decode_grads));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert a == C.sizeof(DecodeDesc) and b == C.sizeof(DecodeGrads), (a, C.sizeof(DecodeDesc), b, C.sizeof(DecodeGrads))


def test_tile_segmented_binning_reaches_the_4k_sweep_sizes():
    """splatco_binning_accepts_capacity (pure host arithmetic): since round 2 the tile-segmented path grows its chunk size
    until the [chunks][tiles] matrix fits, so the 4K render sweep (BASELINE configs[4]) no longer falls back to the
    six-pass radix composition; images whose tile histogram exceeds shared memory still do."""
    L = _lib.lib()
    assert L.splatco_binning_accepts_capacity(680_000, 2_660_000, 545, 980) == 1          # C2
    assert L.splatco_binning_accepts_capacity(6_600_000, 24_000_000, 1080, 1920) == 1      # C4
    assert L.splatco_binning_accepts_capacity(15_000_000, 54_000_000, 2160, 3840) == 1     # C5, 20 M Gaussians
    assert L.splatco_binning_accepts_capacity(750_000, 2_700_000, 2160, 3840) == 1         # C5, 1 M
    assert L.splatco_binning_accepts_capacity(1_000_000, 4_000_000, 4320, 7680) == 0       # 8K: 129,600 tiles > 200 KB histogram
    assert L.splatco_binning_accepts_capacity(0, 10, 545, 980) == 0 and L.splatco_binning_accepts_capacity(10, 0, 545, 980) == 0
