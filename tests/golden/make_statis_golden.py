"""Generates tests/golden/statis.npz by running the REFERENCE's own GaussianModel.training_statis
(scene/gaussian_model.py:761-782) on CPU, unmodified, on seeded inputs of the shapes the render path hands it
(train.py:264-266): neural_opacity [V*K,1], visibility_filter [M], selection mask [V*K], voxel_visible_mask [N],
viewspace_points.grad [M,3].  Two consecutive calls (the accumulators carry over).  Run in the authoring
container:  python tests/golden/make_statis_golden.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_decode_golden import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    _, gm = import_reference()
    g = torch.Generator().manual_seed(4242)
    N, K = 3000, 10
    me = SimpleNamespace(n_offsets=K,
                         opacity_accum=torch.zeros(N, 1), anchor_demon=torch.zeros(N, 1),
                         offset_gradient_accum=torch.zeros(N * K, 1), offset_denom=torch.zeros(N * K, 1))
    out = {"N": N, "K": K}
    for call in range(2):
        vis = torch.rand(N, generator=g) < (0.7 if call == 0 else 0.4)
        V = int(vis.sum())
        nopac = torch.tanh(torch.randn(V * K, 1, generator=g))
        sel = (nopac > 0).view(-1)
        M = int(sel.sum())
        upd = torch.rand(M, generator=g) < 0.8
        vp = SimpleNamespace(grad=torch.randn(M, 3, generator=g) * 1e-3)
        gm.GaussianModel.training_statis(me, vp, nopac, upd, sel, vis)
        out.update({f"c{call}.vis": vis.numpy(), f"c{call}.nopac": nopac.numpy(), f"c{call}.sel": sel.numpy(),
                    f"c{call}.upd": upd.numpy(), f"c{call}.grad": vp.grad.numpy(),
                    f"c{call}.opacity_accum": me.opacity_accum.numpy().copy(), f"c{call}.anchor_demon": me.anchor_demon.numpy().copy(),
                    f"c{call}.offset_gradient_accum": me.offset_gradient_accum.numpy().copy(),
                    f"c{call}.offset_denom": me.offset_denom.numpy().copy()})
    np.savez_compressed(os.path.join(OUT, "statis.npz"), **out)
    print("wrote statis.npz", {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("c1.o")})


if __name__ == "__main__":
    main()
