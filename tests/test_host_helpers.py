"""CPU: small host-side helpers that need no GPU — allocation buckets, FusedAdam argument checks, and that the
training-loop ops refuse CPU tensors instead of falling back."""
import pytest
import torch


def test_bucket_is_monotone_tight_and_coarse():
    from splatco_b200._lib import bucket
    for n in list(range(0, 5000, 7)) + [10 ** k + d for k in range(4, 10) for d in (-1, 0, 1, 12345)]:
        b = bucket(n)
        assert b >= n and b <= n + max(n // 4, 0) + 1, (n, b)      # at most 25 % slack
        if n <= 4096:
            assert b == n
    vals = [bucket(n) for n in range(1 << 20, 1 << 21, 4099)]
    assert vals == sorted(vals)                                       # monotone
    assert len({bucket(n) for n in range(1 << 20, 1 << 21, 257)}) <= 5


def test_fused_adam_rejects_what_it_does_not_implement():
    from splatco_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(NotImplementedError):
        FusedAdam([p], lr=0.1, weight_decay=0.01)
    with pytest.raises(NotImplementedError):
        FusedAdam([p], lr=0.1, amsgrad=True)
    opt = FusedAdam([{"params": [p], "lr": 0.1, "name": "anchor"}], lr=0.0, eps=1e-15)
    assert opt.param_groups[0]["name"] == "anchor" and opt.param_groups[0]["eps"] == 1e-15
    opt.step()                                    # no gradients anywhere: nothing to do, no library call, no error
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):             # CPU parameters: loud failure, never a fallback
        opt.step()


def test_new_ops_fail_loudly_on_cpu_tensors():
    from splatco_b200.cvpm import cvpm_mask
    from splatco_b200.densify import grow_pass
    from splatco_b200.loss import multiview_consistency_loss, scaling_reg
    with pytest.raises(RuntimeError):
        scaling_reg(torch.rand(5, 3))
    with pytest.raises(RuntimeError):
        multiview_consistency_loss([torch.rand(3, 4, 4), torch.rand(3, 4, 4)], [torch.rand(3, 4, 4), torch.rand(3, 4, 4)],
                                   pair_ssim_values=torch.ones(1))
    with pytest.raises(RuntimeError):
        cvpm_mask(torch.rand(10, 3), torch.zeros(3), torch.ones(3))
    with pytest.raises(RuntimeError):
        grow_pass(torch.rand(4, 3), torch.rand(4, 10, 3), torch.rand(4, 6), torch.rand(4, 32), 0.1,
                  candidate_mask=torch.ones(40, dtype=torch.bool))
