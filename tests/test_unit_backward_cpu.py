"""Host logic of criteria._fused_apply: `loss.backward()` on the criterion's own result is recognised on the host (the root
gradient is autograd's implicit ones) and takes the stashed gradient without the scaling launch; every other route scales."""
import torch

from mono_depth_estimation_b200 import criteria


def _run(monkeypatch, how):
    calls = []

    def fake_scale(grad, grad_output):
        calls.append(float(grad_output))
        return grad * grad_output

    monkeypatch.setattr(criteria, "_scale_grad", fake_scale)
    x = torch.arange(6, dtype=torch.float32).requires_grad_(True)

    def launch(p, need_grad):
        assert need_grad
        return (p.detach() ** 2).sum(), 2 * p.detach()

    loss = criteria._fused_apply(x, launch)
    how(loss)
    return x.grad, calls


def test_plain_backward_skips_the_scale_launch(monkeypatch):
    g, calls = _run(monkeypatch, lambda l: l.backward())
    assert calls == [] and torch.equal(g, 2 * torch.arange(6.))


def test_every_other_route_scales(monkeypatch):
    g, calls = _run(monkeypatch, lambda l: (3 * l).backward())
    assert calls == [3.0] and torch.equal(g, 6 * torch.arange(6.))
    g, calls = _run(monkeypatch, lambda l: l.backward(torch.tensor(0.5)))
    assert calls == [0.5] and torch.equal(g, torch.arange(6.))
    g, calls = _run(monkeypatch, lambda l: torch.autograd.backward([l]))
    assert calls == [1.0] and torch.equal(g, 2 * torch.arange(6.))
    g, calls = _run(monkeypatch, lambda l: (l + l.detach()).backward())
    assert calls == [1.0]


def test_flag_is_lowered_after_the_call_and_no_grad_result_is_plain(monkeypatch):
    x = torch.ones(3, requires_grad=True)
    loss = criteria._fused_apply(x, lambda p, need: (p.detach().sum(), torch.ones(3)))
    loss.backward()
    y = torch.ones(3)
    plain = criteria._fused_apply(y, lambda p, need: (p.sum(), None))
    assert not plain.requires_grad and type(plain) is torch.Tensor
    # the temporary of `criterion(p, t).backward()` stays alive through the call, results of arithmetic are plain tensors
    z = torch.ones(3, requires_grad=True)
    criteria._fused_apply(z, lambda p, need: (p.detach().sum(), 3 * torch.ones(3))).backward()
    assert torch.equal(z.grad, 3 * torch.ones(3))
    l2 = criteria._fused_apply(z, lambda p, need: (p.detach().sum(), torch.ones(3)))
    assert type(l2 * 2) is torch.Tensor and type(l2.detach()) is torch.Tensor and isinstance(l2, torch.Tensor)
