"""ORACLE (test infrastructure, not product code): CPU restatement of the masked depth losses.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product package (mono_depth_estimation_b200) never does and fails
loudly when its CUDA library is missing.

Each function restates one reference loss as a plain function of CPU tensors, in the SAME
floating-point operation order as the reference so that an fp32 run reproduces the reference
to rounding, and an fp64 run (pass double tensors) gives the "true" value that separates the
kernel's error from the reference's own fp32 summation error. Gradients come from autograd
over the restated forward, exactly as in the reference.

Parity is pinned: tests/test_oracle_vs_golden.py checks every function here against golden
vectors produced by the reference's own code (oracle/gen_golden.py imports
/root/reference/criteria.py by path), and tests/test_oracle_vs_reference.py re-runs that
comparison live whenever /root/reference is present.
"""
from __future__ import annotations

import torch


def _check_dims(pred, target):
    # reference: criteria.py:22,72,85,116
    assert pred.dim() == target.dim(), "inconsistent dimensions"


def masked_l1(pred, target):
    """reference criteria.py:80-90 (MaskedL1Loss.forward)."""
    _check_dims(pred, target)
    valid = (target > 0).detach()
    return (target - pred)[valid].abs().mean()


def masked_mse(pred, target):
    """reference criteria.py:67-77 (MaskedMSELoss.forward)."""
    _check_dims(pred, target)
    valid = (target > 0).detach()
    return ((target - pred)[valid] ** 2).mean()


def berhu(pred, target):
    """reference criteria.py:111-133 (berHuLoss.forward).

    Quirks kept: the threshold is 0.2*max(pred-target) over ALL pixels, signed (:118-119);
    the loss is the mean of the concatenation [|d| over valid ; |d|^2 over valid & |d|>c] (:131).
    """
    _check_dims(pred, target)
    c = 0.2 * torch.max(pred - target)
    valid = (target > 0).detach()
    a = (target - pred)[valid].abs()
    hub = (a > c).detach()
    return torch.cat((a, a[hub] ** 2)).mean()


def laina_berhu(inp, target, mask=None, size_average=True, use_logs=True, clamp_val=1e-9):
    """reference criteria.py:476-506 (LainaBerHuLoss.forward); c=0.2*max is differentiable."""
    if mask is None:
        mask = target > 0
    if use_logs:
        n = torch.log(inp.clamp(min=clamp_val)) - torch.log(target.clamp(min=clamp_val))
    else:
        n = inp - target
    n = torch.abs(n) * mask
    n = n.squeeze(1)
    c = 0.2 * n.max()
    loss = torch.where(n < c, n, (n ** 2 + c ** 2) / (2 * c + 1e-9)).sum()
    if size_average:
        return loss / mask.sum()
    return loss


def silog(depth_est, depth_gt, variance_focus=0.85):
    """reference criteria.py:724-732 (silog_loss.forward); mask is gt > 1e-2, not > 0."""
    mask = depth_gt > 1e-2
    d = torch.log(depth_est[mask]) - torch.log(depth_gt[mask])
    return torch.sqrt((d ** 2).mean() - variance_focus * (d.mean() ** 2)) * 10.0


def eigen_masked_depth(pred, target):
    """reference criteria.py:17-64 (MaskedDepthLoss.forward): Eigen scale-invariant + gradient term."""
    _check_dims(pred, target)
    B = target.shape[0]
    mask = (target > 0).detach().to(pred.dtype)
    t = target.reshape(B, -1)
    m = mask.reshape(B, -1)
    p = pred.reshape(B, -1)
    d = p * m - t * m
    n_b = m.sum(dim=1)
    depth_cost = ((n_b * (d ** 2).sum(dim=1)).sum() - 0.5 * (d.sum(dim=1) ** 2).sum()) / (n_b ** 2).sum()
    if pred.ndim == 4:
        pred = pred[:, 0]
    if target.ndim == 4:
        target = target[:, 0]
    if mask.ndim == 4:
        mask = mask[:, 0]
    p_di = pred[:, 1:, :] - pred[:, :-1, :]
    p_dj = pred[:, :, 1:] - pred[:, :, :-1]
    t_di = target[:, 1:, :] - target[:, :-1, :]
    t_dj = target[:, :, 1:] - target[:, :, :-1]
    m_di = torch.logical_and(mask[:, 1:, :], mask[:, :-1, :])
    m_dj = torch.logical_and(mask[:, :, 1:], mask[:, :, :-1])
    grad_cost = (m_di * (p_di - t_di) ** 2).sum() / m_di.sum() \
        + (m_dj * (p_dj - t_dj) ** 2).sum() / m_dj.sum()
    return depth_cost + grad_cost


LOSSES = {
    "l1": masked_l1,
    "mse": masked_mse,
    "berhu": berhu,
    "laina_berhu": laina_berhu,
    "silog": silog,
    "eigen": eigen_masked_depth,
}


def loss_and_grad(fn, pred, *args, **kwargs):
    """Run `fn(pred, *args)` with autograd and return (loss, dloss/dpred) as detached tensors."""
    p = pred.detach().clone().requires_grad_(True)
    loss = fn(p, *args, **kwargs)
    (g,) = torch.autograd.grad(loss, p, allow_unused=True)
    if g is None:
        g = torch.zeros_like(p)
    return loss.detach(), g.detach()
