"""ORACLE (test infrastructure, not product code): CPU restatement of the MiDaS alignment step.

  compute_scale_and_shift   reference criteria.py:154-176
  scale_shift               reference modules/midas.py:56-62 (MidasModule.scale_shift)

Pinned by tests/golden/midas_small.npz, produced by the reference's own criteria.compute_scale_and_shift.
"""
from __future__ import annotations

import torch


def compute_scale_and_shift(prediction, target, mask=None):
    """prediction, target [B,H,W]; mask float [B,H,W] (default target > 0). Returns (scale [B], shift [B])."""
    if mask is None:
        mask = (target > 0).type(prediction.dtype)
    a_00 = torch.sum(mask * prediction * prediction, (1, 2))
    a_01 = torch.sum(mask * prediction, (1, 2))
    a_11 = torch.sum(mask, (1, 2))
    b_0 = torch.sum(mask * prediction * target, (1, 2))
    b_1 = torch.sum(mask * target, (1, 2))
    x_0 = torch.zeros_like(b_0)
    x_1 = torch.zeros_like(b_1)
    det = a_00 * a_11 - a_01 * a_01
    valid = torch.nonzero(det, as_tuple=True)
    x_0[valid] = (a_11[valid] * b_0[valid] - a_01[valid] * b_1[valid]) / det[valid]
    x_1[valid] = (-a_01[valid] * b_0[valid] + a_00[valid] * b_1[valid]) / det[valid]
    return x_0, x_1


def scale_shift(pred, target):
    """modules/midas.py:56-62: [B,1,H,W] (or [B,H,W]) in, aligned pred and target as [B,1,H,W] out."""
    if pred.ndim == 4:
        pred = pred.squeeze(1)
    if target.ndim == 4:
        target = target.squeeze(1)
    scale, shift = compute_scale_and_shift(pred, target)
    pred = scale.view(-1, 1, 1) * pred + shift.view(-1, 1, 1)
    return pred.unsqueeze(1), target.unsqueeze(1)
