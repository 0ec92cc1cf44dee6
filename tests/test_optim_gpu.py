"""FusedAdam against torch.optim.Adam — the optimizer the reference builds
(scene/gaussian_model.py:519-572: eps 1e-15, per-group lr) — run on CPU in float64-free plain fp32: parameters after
several steps agree to 2e-6 relative (floor 1e-7), moments to 1e-6 of their scale; state keys / param_groups stay torch-compatible."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _params(seed, device):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1000, 3), (1000, 10, 3), (1000, 32), (1, 5, 61, 61), (32, 99), (32,), (7,), (3, 1, 1, 1), (70000,)]
    return [nn.Parameter((torch.randn(*s, generator=g) * 0.3).to(device)) for s in shapes]


def test_fused_adam_matches_torch_adam():
    from splatco_b200.optim import FusedAdam
    ref_p, our_p = _params(1, "cpu"), _params(1, "cuda")
    lrs = [1.6e-4, 0.01, 0.0075, 0.005, 0.002, 0.002, 0.0, 0.008, 0.05]
    mk = lambda ps: [{"params": [p], "lr": lr, "name": f"g{i}"} for i, (p, lr) in enumerate(zip(ps, lrs))]
    ref = torch.optim.Adam(mk(ref_p), lr=0.0, eps=1e-15)
    our = FusedAdam(mk(our_p), lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(2)
    for it in range(6):
        for n, (a, b) in enumerate(zip(ref_p, our_p)):
            if it == 2 and n == 1:                      # a parameter without a gradient is skipped (its step does not advance)
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g)))
            if n == 2:
                gr[::3] = 0.0                           # zero gradients with eps 1e-15: 0 / (0 + eps) stays finite
            a.grad, b.grad = gr.clone(), gr.cuda()
        if it == 3:
            for opt in (ref, our):
                opt.param_groups[0]["lr"] = 3e-5        # the reference rewrites group lrs every iteration
        ref.step()
        our.step()
    for n, (a, b) in enumerate(zip(ref_p, our_p)):
        want, got = a.detach().numpy(), b.detach().cpu().numpy()
        assert np.isfinite(got).all()
        assert (np.abs(got - want) <= 2e-6 * np.abs(want) + 1e-7).all(), (n, np.abs(got - want).max())
        sr, so = ref.state[a], our.state[b]
        assert set(so.keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(so["step"]) == float(sr["step"])
        for key in ("exp_avg", "exp_avg_sq"):            # moments: 1e-6 of the tensor's scale (lerp of opposite-sign terms cancels)
            w = sr[key].numpy()
            assert np.abs(so[key].cpu().numpy() - w).max() <= 1e-6 * np.abs(w).max(), (n, key)
    sd = our.state_dict()                                # torch-compatible checkpoint layout
    assert len(sd["param_groups"]) == len(lrs) and sd["param_groups"][0]["name"] == "g0"


def test_fused_adam_many_tensors_and_errors():
    from splatco_b200.optim import FusedAdam
    ps = [nn.Parameter(torch.full((17 + i,), 1.0, device="cuda")) for i in range(150)]        # > 64 tensors: several launches
    opt = FusedAdam(ps, lr=0.1, eps=1e-15)
    for p in ps:
        p.grad = torch.ones_like(p)
    opt.step()
    for p in ps:                                         # first Adam step moves every element by exactly lr (bias-corrected m/sqrt(v) = 1)
        assert torch.allclose(p.detach(), torch.full_like(p, 0.9), rtol=1e-6)
    with pytest.raises(RuntimeError):
        q = nn.Parameter(torch.zeros(4))
        o = FusedAdam([q], lr=0.1)
        q.grad = torch.ones(4)
        o.step()
    with pytest.raises(NotImplementedError):
        FusedAdam(ps, lr=0.1, weight_decay=0.1)
