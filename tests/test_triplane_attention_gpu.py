"""TriPlaneAttention kernels (csrc/triplane_attention.cu) against the module maths of the reference
(scene/grids.py:22-64, mirrored by splatco_b200.model.TriPlaneAttention, evaluated by torch in fp64)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(rc, E, seed):
    from splatco_b200.decode import _TriPlaneAttention
    from splatco_b200.model import TriPlaneAttention
    g = torch.Generator().manual_seed(seed)
    C = 3 * rc
    planes = [(torch.randn(1, rc, E, E, generator=g) * 0.5) for _ in range(3)]
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        ta = TriPlaneAttention(C)
    gout = [torch.randn(1, rc, E, E, generator=g) for _ in range(3)]
    # reference maths in fp64 on the CPU
    ta64 = ta.double()
    p64 = [p.double().requires_grad_() for p in planes]
    out64 = torch.chunk(ta64(torch.cat(p64, dim=1)), 3, dim=1)
    sum((o * go.double()).sum() for o, go in zip(out64, gout)).backward()
    # kernels
    d = "cuda:0"
    pg = [p.to(d).requires_grad_() for p in planes]
    ws = [w.detach().float().to(d).requires_grad_() for w in (ta.ca.sharedMLP[0].weight, ta.ca.sharedMLP[2].weight, ta.sa.conv.weight)]
    outs = _TriPlaneAttention.apply(*pg, *ws)
    sum((o * go.to(d)).sum() for o, go in zip(outs, gout)).backward()
    for o, r in zip(outs, out64):
        assert (o.detach().cpu().double() - r.detach()).abs().max() < 2e-6
    refs = [p.grad for p in p64] + [ta64.ca.sharedMLP[0].weight.grad, ta64.ca.sharedMLP[2].weight.grad, ta64.sa.conv.weight.grad]
    for name, got, ref in zip(("xy", "xz", "yz", "w_ca1", "w_ca2", "w_sa"), pg + ws, refs):
        err = (got.grad.cpu().double() - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 1e-3 * scale + 1e-7, f"{name}: err {err} vs scale {scale}"      # 1e-3 relative (north star), floor 1e-3 * max|g|


@pytest.mark.parametrize("rc,E", [(5, 70), (5, 33), (2, 96), (5, 161), (2, 20)])
def test_triplane_attention_matches_reference_maths(rc, E):
    _run(rc, E, 11 + E)
