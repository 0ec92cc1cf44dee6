"""CPU tests of splatco_shims (import-time stand-ins for the packages the reference imports but this image lacks)."""
import inspect
import sys

import numpy as np
import torch


def test_install_and_module_introspection():
    import splatco_shims
    done = splatco_shims.install(hot_path=None)
    for name in ("simple_knn._C", "torch_scatter", "_gridcreater", "_gridencoder", "plyfile", "kornia"):
        assert name in sys.modules, (name, done)
    # inspect.getmodule walks every sys.modules entry and reads __file__: the stubs must not answer that with a callable
    assert inspect.getmodule(torch.nn.Linear) is not None
    import _gridcreater
    try:
        _gridcreater.anything(1)
        raise AssertionError("stub must raise")
    except NotImplementedError:
        pass


def test_scatter_max_and_knn_and_ply(tmp_path):
    from splatco_shims import _impl
    g = torch.Generator().manual_seed(0)
    src = torch.randn(200, 5, generator=g)
    idx = torch.randint(0, 9, (200,), generator=g)
    val, arg = _impl.scatter_max(src, idx.unsqueeze(1).expand(-1, 5), dim=0)
    for grp in range(9):
        m = idx == grp
        if m.any():
            assert torch.equal(val[grp], src[m].max(0).values)
            assert torch.equal(src[arg[grp], torch.arange(5)], val[grp])
    P = torch.randn(300, 3, generator=g)
    d = _impl.distCUDA2(P)
    dd = torch.cdist(P, P) ** 2
    dd.fill_diagonal_(float("inf"))
    assert torch.allclose(d, dd.topk(3, largest=False).values.mean(1), atol=1e-5)
    arr = np.zeros(7, dtype=[("x", "f4"), ("y", "f4"), ("n", "u1")])
    arr["x"] = np.arange(7)
    path = str(tmp_path / "t.ply")
    _impl.PlyData([_impl.PlyElement.describe(arr, "vertex")]).write(path)
    back = _impl.PlyData.read(path)
    assert np.array_equal(back["vertex"]["x"], arr["x"]) and back["vertex"].data.dtype.names == ("x", "y", "n")
    grid = _impl.create_meshgrid(4, 6, normalized_coordinates=False)
    assert grid.shape == (1, 4, 6, 2) and float(grid[0, 3, 5, 0]) == 5.0 and float(grid[0, 3, 5, 1]) == 3.0
