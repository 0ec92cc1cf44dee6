"""Fused L1 + SSIM loss against tests/golden/loss.npz, produced by the reference's own
utils/loss_utils.py (tests/golden/make_loss_golden.py): value to 1e-6, gradient to 1e-3 relative (floor 1e-3 max|g|)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz")


@pytest.mark.parametrize("n", [0, 1])
def test_l1_ssim_matches_reference(n):
    from splatco_b200.loss import l1_loss, l1_ssim_loss, ssim
    d = np.load(GOLD)
    img = torch.from_numpy(d[f"p{n}.img"]).cuda().requires_grad_()
    gt = torch.from_numpy(d[f"p{n}.gt"]).cuda()
    loss, parts = l1_ssim_loss(img, gt, 0.2, return_parts=True)
    (loss * 1.0).backward()
    assert abs(loss.item() - float(d[f"p{n}.loss"])) < 1e-6
    assert abs(parts[1].item() - float(d[f"p{n}.l1"])) < 1e-6 and abs(parts[2].item() - float(d[f"p{n}.ssim"])) < 1e-6
    want = d[f"p{n}.grad"]
    err = np.abs(img.grad.cpu().numpy() - want)
    assert (err <= 1e-3 * np.maximum(np.abs(want), 1e-3 * np.abs(want).max())).all(), err.max()
    with torch.no_grad():
        assert abs(l1_loss(img, gt).item() - float(d[f"p{n}.l1"])) < 1e-6
        assert abs(ssim(img, gt).item() - float(d[f"p{n}.ssim"])) < 1e-6


def test_l1_ssim_scales_with_upstream_gradient_and_large_image():
    from splatco_b200.loss import l1_ssim_loss
    g = torch.Generator(device="cuda").manual_seed(3)
    gt = torch.rand(3, 545, 980, device="cuda", generator=g)
    img = (gt + 0.1 * torch.randn(3, 545, 980, device="cuda", generator=g)).requires_grad_()
    l1_ssim_loss(img, gt, 0.2).backward()
    g1 = img.grad.clone()
    img.grad = None
    (3.0 * l1_ssim_loss(img, gt, 0.2)).backward()
    assert torch.allclose(img.grad, 3.0 * g1, rtol=1e-5, atol=1e-12)
    # finite-difference check of the summed loss along a random direction (fp32: loose)
    v = torch.randn_like(img)
    eps = 1e-2
    with torch.no_grad():
        lp = l1_ssim_loss(img + eps * v, gt, 1.0).double().item()
        lm = l1_ssim_loss(img - eps * v, gt, 1.0).double().item()
    img.grad = None
    l1_ssim_loss(img, gt, 1.0).backward()
    ana = (img.grad.double() * v.double()).sum().item()
    assert abs((lp - lm) / (2 * eps) - ana) <= 5e-2 * abs(ana) + 1e-7


def test_scaling_reg_matches_torch_prod_mean():
    """train.py:195 `scaling.prod(dim=1).mean()`: value and gradient, including rows with zeros (where torch's backward takes
    its slow exact path) and a non-unit upstream gradient."""
    from splatco_b200.loss import scaling_reg
    g = torch.Generator(device="cuda").manual_seed(11)
    s = torch.rand(700_001, 3, device="cuda", generator=g) * 0.05
    s[::1000, 1] = 0.0
    s[5::2000] = 0.0
    a = s.clone().requires_grad_()
    b = s.clone().requires_grad_()
    la = 0.01 * scaling_reg(a)
    lb = 0.01 * b.prod(dim=1).mean()
    la.backward()
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-6 * abs(lb.item())
    assert torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-20)
    with pytest.raises(RuntimeError):
        scaling_reg(torch.rand(4, 3))
