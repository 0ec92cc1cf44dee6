"""Import-time stand-ins that let the reference's train.py / render.py import in an image without its four CUDA
extensions and three pip packages (reference import sites: scene/gaussian_model.py:15,20,22,
utils/grid_utils.py:5-6, utils/camera_utils.py:16, gaussian_renderer/__init__.py:15).  OUT of the hot path: plain torch /
numpy implementations of init-time and densification helpers, import-only stubs for code the reference never executes.

    import splatco_shims
    splatco_shims.install()            # before `import train` / `import scene` / `import gaussian_renderer`

install(hot_path=...) also aliases the two packages the hot path enters native code through:
    "rasterizer": diff_gaussian_rasterization -> splatco_b200.diff_gaussian_rasterization   (INTEGRATION.md option A)
    "full"      : + gaussian_renderer -> splatco_b200.gaussian_renderer                      (option B, the default)
    None        : neither (the caller aliases them itself)
A module that is really installed is never replaced.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types


def _have(name: str) -> bool:
    if name in sys.modules:
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError, AttributeError):
        return False


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__splatco_shim__ = True
    sys.modules[name] = m
    return m


def install(hot_path: str | None = "full") -> list[str]:
    """Returns the names of the modules that were shimmed."""
    from . import _impl
    done = []
    if not _have("simple_knn"):
        pkg = _module("simple_knn")
        pkg.__path__ = []
        pkg._C = _module("simple_knn._C", distCUDA2=_impl.distCUDA2)
        done += ["simple_knn", "simple_knn._C"]
    if not _have("torch_scatter"):
        _module("torch_scatter", scatter_max=_impl.scatter_max)
        done.append("torch_scatter")
    for name in ("_gridcreater", "_gridencoder"):
        if not _have(name):
            _module(name, __getattr__=_impl.never_executed(name))
            done.append(name)
    if not _have("plyfile"):
        _module("plyfile", PlyData=_impl.PlyData, PlyElement=_impl.PlyElement)
        done.append("plyfile")
    if not _have("kornia"):
        _module("kornia", create_meshgrid=_impl.create_meshgrid)
        done.append("kornia")
    if hot_path in ("rasterizer", "full"):
        sys.modules["diff_gaussian_rasterization"] = importlib.import_module("splatco_b200.diff_gaussian_rasterization")
        done.append("diff_gaussian_rasterization")
    if hot_path == "full":
        sys.modules["gaussian_renderer"] = importlib.import_module("splatco_b200.gaussian_renderer")
        done.append("gaussian_renderer")
    return done
