"""CPU tests: the decode oracle (oracle/decode_oracle.py) against golden vectors produced by the
reference's own Python decode (tests/golden/make_decode_golden.py).  Outputs 1e-5 abs, autograd
gradients 1e-3 relative (floor 1e-3*max|g|; fp32 reassociation); the opacity mask must match exactly except where
|neural_opacity| < 1e-6 (none in the fixtures)."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from tests.util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["xyz", "color", "opacity", "scaling", "rot"]


def load(tag):
    d = np.load(os.path.join(GOLD, f"decode_{tag}.npz"))
    p = {k[len("param."):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("param.")}
    return d, p


@pytest.mark.parametrize("tag", ["base", "variants"])
@pytest.mark.parametrize("level", [0, 1, 2])
def test_decode_oracle_matches_reference_fixture(tag, level):
    d, p = load(tag)
    leaves = {}
    for k, v in p.items():
        if v.dtype.is_floating_point and "running" not in k and "xyz_m" not in k:
            p[k] = v.clone().requires_grad_()
            leaves[k] = p[k]
    ins = {k: torch.from_numpy(d[f"in.{k}"]).requires_grad_() for k in ("_anchor", "_offset", "_anchor_feat", "_scaling")}
    outs = D.decode(p, ins["_anchor_feat"], ins["_anchor"], ins["_offset"], torch.exp(ins["_scaling"]),
                    torch.from_numpy(d["vis"]), torch.from_numpy(d["cam_center"]), level, int(d["K"]),
                    appearance_dim=int(d["appearance_dim"]), uid=int(d["uid"]), dists=bool(d["dists"]))
    assert np.array_equal(outs[6].numpy(), d[f"L{level}.out.mask"])
    assert np.abs(outs[5].detach().numpy() - d[f"L{level}.out.neural_opacity"]).max() < 1e-5
    loss = 0
    for nm, t in zip(NAMES, outs[:5]):
        want = d[f"L{level}.out.{nm}"]
        assert t.shape == want.shape
        assert np.abs(t.detach().numpy() - want).max() < 1e-5, nm
        loss = loss + (t * torch.from_numpy(d[f"L{level}.w.{nm}"])).sum()
    loss.backward()
    checked = 0
    for k in d.files:
        if not k.startswith(f"L{level}.grad."):
            continue
        name = k[len(f"L{level}.grad."):]
        got = ins[name].grad if name in ins else leaves[name].grad
        assert got is not None, name
        assert rel_err(got.numpy(), d[k]) < 1e-3, name
        checked += 1
    assert checked >= 20


def test_bn_running_stats_restated():
    d, p = load("base")
    anchor = torch.from_numpy(d["in._anchor"])
    vis = torch.from_numpy(d["vis"])
    x = D.plane_features(p, 1, anchor[vis])
    rm, rv = p["feat.models.1.0.running_mean"], p["feat.models.1.0.running_var"]
    # level-1 BN ran twice in the generator (activate_level 1 and 2)
    for _ in range(2):
        rm, rv = D.bn_running_update(x, rm, rv)
    assert np.abs(rm.numpy() - d["after.feat.models.1.0.running_mean"]).max() < 1e-6
    assert np.abs(rv.numpy() - d["after.feat.models.1.0.running_var"]).max() < 1e-6
    assert int(d["after.feat.models.1.0.num_batches_tracked"]) == 2
