"""Host logic of the shared gradient buffers (splatco_b200/_gradacc.py), on CPU tensors: a toy
Function with the same protocol as the decode node (accumulate in place, the first node of a
backward pass returns the buffer, later ones return None) must give autograd's own sums."""
import torch

from splatco_b200 import _gradacc


class _Scale(torch.autograd.Function):
    """y = a * w + b * s   (w is shared between calls, s is a per-call tensor)"""

    @staticmethod
    def forward(ctx, w, s, a, b):
        ctx.origs = (w, s)
        ctx.a, ctx.b = a, b
        return a * w.detach() + b * s.detach()

    @staticmethod
    def backward(ctx, g):
        w, s = ctx.origs
        need = ctx.needs_input_grad
        got = _gradacc.acquire(w.device, [(id(w) if need[0] else None, w.shape), (id(s) if need[1] else None, s.shape)], want_views=True)
        (_, ret_w, buf_w), (_, ret_s, buf_s) = got
        buf_w += ctx.a * g
        buf_s += ctx.b * g
        del got, buf_w, buf_s
        return ret_w if need[0] else None, ret_s if need[1] else None, None, None


def _run(share_s):
    torch.manual_seed(0)
    w = torch.randn(5, 3, requires_grad=True)
    p = torch.randn(5, 3, requires_grad=True)
    outs = []
    s_shared = torch.exp(p)
    for v in range(4):
        s = s_shared if share_s else torch.exp(p)           # get_scaling-style per-view temporary
        outs.append((_Scale.apply(w, s, float(v + 1), 0.5) * (v + 2)).sum())
    return w, p, outs


def test_single_backward_over_summed_loss_matches_autograd():
    for share_s in (False, True):
        w, p, outs = _run(share_s)
        sum(outs).backward()
        gw = sum((v + 1.0) * (v + 2.0) for v in range(4))
        assert torch.allclose(w.grad, torch.full_like(w, gw))
        gp = sum(0.5 * (v + 2.0) for v in range(4))
        assert torch.allclose(p.grad, gp * torch.exp(p.detach()))
        assert not _gradacc._passes, "table must be dropped at the end of the backward pass"


def test_separate_backward_calls_accumulate_into_grad():
    w, p, outs = _run(False)
    for o in outs:
        o.backward()
    gw = sum((v + 1.0) * (v + 2.0) for v in range(4))
    assert torch.allclose(w.grad, torch.full_like(w, gw))
    assert not _gradacc._passes


def test_grad_buffer_is_adopted_without_copy_and_no_grad_inputs_are_private():
    w = torch.randn(4, requires_grad=True)
    s = torch.randn(4)                                       # does not require grad
    y = _Scale.apply(w, s, 2.0, 1.0) + _Scale.apply(w, s, 3.0, 1.0)
    y.sum().backward()
    assert torch.allclose(w.grad, torch.full_like(w, 5.0))
    assert w.grad.untyped_storage().size() > w.numel() * 4, "expected .grad to be a view of the flat buffer"
    (g,) = torch.autograd.grad((_Scale.apply(w, s, 2.0, 1.0) + _Scale.apply(w, s, 3.0, 1.0)).sum(), [w])
    assert torch.allclose(g, torch.full_like(w, 5.0))
