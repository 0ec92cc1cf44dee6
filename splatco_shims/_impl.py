"""Torch / numpy bodies of the shims (init-time and densification helpers; none is on the render hot path)."""
from __future__ import annotations

import numpy as np
import torch


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """simple_knn._C.distCUDA2 (used by create_from_pcd, scene/gaussian_model.py:478,494): mean squared distance of
    every point to its 3 nearest neighbours.  Chunked brute force on the points' device."""
    p = points.detach().float()
    n = p.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=p.device)
    if n <= 1:
        return out.zero_()
    k = min(3, n - 1)
    step = max(1, min(n, (1 << 26) // max(n, 1)))
    for s in range(0, n, step):
        d2 = torch.cdist(p[s:s + step], p).square_()
        d2[torch.arange(d2.shape[0], device=p.device), torch.arange(s, s + d2.shape[0], device=p.device)] = float("inf")
        out[s:s + step] = d2.topk(k, dim=1, largest=False).values.mean(dim=1)
    return out


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None, dim_size=None):
    """torch_scatter.scatter_max for the call in anchor_growing (scene/gaussian_model.py:897): values and arg-max."""
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    index = index.expand_as(src)
    size = list(src.shape)
    size[dim] = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    res = torch.full(size, torch.finfo(src.dtype).min if src.dtype.is_floating_point else torch.iinfo(src.dtype).min,
                     dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, index, src, reduce="amax", include_self=True)
    pos = torch.arange(src.shape[dim], device=src.device).view([-1 if d == dim else 1 for d in range(src.dim())]).expand_as(src)
    hit = src == res.gather(dim, index)
    arg = torch.full(size, src.shape[dim], dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index, torch.where(hit, pos, torch.full_like(pos, src.shape[dim])), reduce="amin", include_self=True)
    empty = torch.ones(size, dtype=torch.bool, device=src.device).scatter(dim, index, torch.zeros_like(hit))
    res = torch.where(empty, torch.zeros_like(res), res)
    if out is not None:
        out.copy_(res)
        res = out
    return res, arg


def never_executed(modname: str):
    def _getattr(name):
        if name.startswith("__"):          # module introspection (inspect.getmodule reads __file__ of every sys.modules entry)
            raise AttributeError(name)

        def _raise(*a, **k):
            raise NotImplementedError(
                f"{modname}.{name}: the reference imports this extension but never runs it (Spatial_CTX is constructed and "
                "never called, scene/gaussian_model.py:47-62,149-169); splatco_shims provides the import only")
        return _raise
    return _getattr


def create_meshgrid(height: int, width: int, normalized_coordinates: bool = True, device=None, dtype=torch.float32):
    """kornia.create_meshgrid: [1, H, W, 2] grid of (x, y)."""
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / max(width - 1, 1) - 0.5) * 2
        ys = (ys / max(height - 1, 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)


class PlyElement:
    """Minimal plyfile.PlyElement: holds a structured numpy array (GaussianModel.save_ply, scene/gaussian_model.py:653-670)."""

    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    def __getitem__(self, key):
        return self.data[key]

    @property
    def properties(self):
        return [type("P", (), {"name": n})() for n in (self.data.dtype.names or ())]


class PlyData:
    """Minimal plyfile.PlyData: binary little-endian PLY with one or more scalar-property elements."""

    def __init__(self, elements):
        self.elements = list(elements)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    _TYPES = {"f4": "float", "f8": "double", "i4": "int", "u1": "uchar", "i2": "short", "u2": "ushort", "u4": "uint", "i1": "char"}

    def write(self, path):
        with open(path, "wb") as fh:
            head = ["ply", "format binary_little_endian 1.0"]
            for e in self.elements:
                head.append(f"element {e.name} {len(e.data)}")
                for n in e.data.dtype.names:
                    head.append(f"property {self._TYPES[e.data.dtype[n].str[1:]]} {n}")
            head.append("end_header")
            fh.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                fh.write(e.data.astype(e.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        rev = {v: k for k, v in PlyData._TYPES.items()}
        with open(path, "rb") as fh:
            if fh.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt = fh.readline().split()
            if fmt[1] != b"binary_little_endian":
                raise NotImplementedError("splatco_shims.plyfile reads binary little-endian PLY only")
            elems, cur = [], None
            while True:
                line = fh.readline().split()
                if not line:
                    raise ValueError("truncated PLY header")
                if line[0] == b"element":
                    cur = [line[1].decode(), int(line[2]), []]
                    elems.append(cur)
                elif line[0] == b"property":
                    if line[1] == b"list":
                        raise NotImplementedError("list properties are not supported by the shim")
                    cur[2].append((line[2].decode(), "<" + rev[line[1].decode()]))
                elif line[0] == b"end_header":
                    break
            out = []
            for name, count, props in elems:
                dt = np.dtype(props)
                out.append(PlyElement(name, np.frombuffer(fh.read(dt.itemsize * count), dtype=dt, count=count)))
        return PlyData(out)
