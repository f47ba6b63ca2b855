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


# ---- MidasLoss (criteria.py:306-332) and its parts, restated op for op --------------------------------------
def reduction_batch_based(image_loss, M):
    """criteria.py:179-188."""
    divisor = torch.sum(M)
    if divisor == 0:
        return 0
    return torch.sum(image_loss) / divisor


def mse_loss(prediction, target, mask):
    """criteria.py:219-223 (batch-based)."""
    M = torch.sum(mask, (1, 2))
    res = prediction - target
    return reduction_batch_based(mask * res * res, 2 * M)


def l1_loss(prediction, target, mask):
    """criteria.py:201-206 (batch-based)."""
    M = torch.sum(mask, (1, 2))
    diff = (target - prediction)[mask.bool()]
    return reduction_batch_based(diff.abs(), 2 * M)


def trimmed_mae_loss(prediction, target, mask, trim=0.2):
    """criteria.py:208-217 AS WRITTEN: `torch.sort(...)[: k]` slices the (values, indices) tuple, so nothing is
    trimmed and the value equals the l1 data term (needs k >= 2)."""
    M = torch.sum(mask, (1, 2))
    res = (prediction - target)[mask.bool()].abs()
    trimmed, _ = torch.sort(res.view(-1), descending=False)[: int(len(res) * (1.0 - trim))]
    return reduction_batch_based(trimmed, 2 * M)


def gradient_loss(prediction, target, mask):
    """criteria.py:226-244 (batch-based)."""
    M = torch.sum(mask, (1, 2))
    diff = torch.mul(mask, prediction - target)
    grad_x = torch.abs(diff[:, :, 1:] - diff[:, :, :-1])
    grad_x = torch.mul(torch.mul(mask[:, :, 1:], mask[:, :, :-1]), grad_x)
    grad_y = torch.abs(diff[:, 1:, :] - diff[:, :-1, :])
    grad_y = torch.mul(torch.mul(mask[:, 1:, :], mask[:, :-1, :]), grad_y)
    image_loss = torch.sum(grad_x, (1, 2)) + torch.sum(grad_y, (1, 2))
    return reduction_batch_based(image_loss, M)


def midas_loss(prediction, target, alpha=0.5, scales=4, loss="mse"):
    """MidasLoss.forward (criteria.py:319-332), batch-based reduction; 'ssi' in `loss` adds the alignment."""
    if prediction.ndim == 4:
        prediction = prediction.squeeze(1)
    if target.ndim == 4:
        target = target.squeeze(1)
    mask = (target > 0).type(prediction.dtype)
    if "ssi" in loss:
        scale, shift = compute_scale_and_shift(prediction, target, mask)
        prediction = scale.view(-1, 1, 1) * prediction + shift.view(-1, 1, 1)
    data = trimmed_mae_loss if "trim" in loss else (mse_loss if "mse" in loss else l1_loss)
    total = data(prediction, target, mask)
    if alpha > 0:
        reg = 0
        for s in range(scales):
            step = 2 ** s
            reg = reg + gradient_loss(prediction[:, ::step, ::step], target[:, ::step, ::step], mask[:, ::step, ::step])
        total = total + alpha * reg
    return total


# ---- TrimmedProcrustesLoss (criteria.py:335-363) ------------------------------------------------------------
def normalize_prediction_robust(target, mask=None):
    """criteria.py:135-152, with the mask and the statistics in the dtype of `target` (the reference hard-codes
    float32 at :137 and therefore only runs in fp32; in fp32 this is the same op sequence)."""
    if mask is None:
        mask = (target > 0).type(target.dtype)
    ssum = torch.sum(mask, (1, 2))
    valid = ssum > 0
    m = torch.zeros_like(ssum)
    s = torch.ones_like(ssum)
    m[valid] = torch.median((mask[valid] * target[valid]).view(int(valid.sum()), -1), dim=1).values
    target = target - m.view(-1, 1, 1)
    sq = torch.sum(mask * target.abs(), (1, 2))
    s[valid] = torch.clamp(sq[valid] / ssum[valid], min=1e-6)
    return target / s.view(-1, 1, 1)


def trimmed_procrustes_loss(prediction, target, alpha=0.5, scales=4):
    """TrimmedProcrustesLoss.forward (criteria.py:345-358), batch-based. Returns (loss, normalised prediction)."""
    if prediction.ndim == 4:
        prediction = prediction.squeeze(1)
    if target.ndim == 4:
        target = target.squeeze(1)
    mask = (target > 0).type(prediction.dtype)
    p_ssi = normalize_prediction_robust(prediction, mask)
    t_ssi = normalize_prediction_robust(target, mask)
    total = trimmed_mae_loss(p_ssi, t_ssi, mask)
    if alpha > 0:
        reg = 0
        for s in range(scales):
            step = 2 ** s
            reg = reg + gradient_loss(p_ssi[:, ::step, ::step], t_ssi[:, ::step, ::step], mask[:, ::step, ::step])
        total = total + alpha * reg
    return total, p_ssi
