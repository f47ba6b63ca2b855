"""ORACLE (test infrastructure, not product code): CPU restatement of DORN's ordinal head.

  ordinal_layer           reference network/Dorn.py:292-321  (OrdinalRegressionLayer.forward)
  label_to_depth/...      reference modules/dorn.py:95-107   (DORNModule.label_to_depth / depth_to_label)
  sid_table_*             reference modules/dorn.py:10-71    (get_depth_sid / get_labels_sid)
  ord_loss                reference criteria.py:744-787      (ordLoss.forward)
  ordinal_regression_loss reference criteria.py:789-836      (OrdinalRegressionLoss.__call__)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ordinal_layer(x):
    """x [N,2K,H,W] -> (decode int64 [N,1,H,W], P fp [N,K,H,W]).

    Pairs are interleaved: A = even channels, B = odd channels (Dorn.py:305-306); both are
    clamped to [1e-8, 1e4] (:312) BEFORE the 2-way softmax (:314); decode counts P > 0.5 (:319).
    """
    N, C, H, W = x.shape
    K = C // 2
    a = x[:, 0::2].reshape(N, 1, K * H * W)
    b = x[:, 1::2].reshape(N, 1, K * H * W)
    pair = torch.clamp(torch.cat((a, b), dim=1), min=1e-8, max=1e4)
    sm = F.softmax(pair, dim=1)
    P = sm[:, 1, :].reshape(N, K, H, W).clone()
    decode = torch.sum(P > 0.5, dim=1).view(-1, 1, H, W)
    return decode, P


def label_to_depth(label, alpha, beta, ord_num, discretization="SID"):
    """modules/dorn.py:95-100. alpha/beta are fp32 0-dim tensors, ord_num an int32 0-dim tensor
    (modules/dorn.py:76-78)."""
    alpha = torch.as_tensor(alpha).float()
    beta = torch.as_tensor(beta).float()
    K = torch.as_tensor(ord_num).int()
    if discretization == "SID":
        return torch.exp(torch.log(alpha) + torch.log(beta / alpha) * label / K)
    return alpha + (beta - alpha) * label / K


def depth_to_label(depth, alpha, beta, ord_num, discretization="SID"):
    """modules/dorn.py:102-107: float label, NOT floored; depth 0 -> -inf."""
    alpha = torch.as_tensor(alpha).float()
    beta = torch.as_tensor(beta).float()
    K = torch.as_tensor(ord_num).int()
    if discretization == "SID":
        return K * torch.log(depth / alpha) / torch.log(beta / alpha)
    return K * (depth - alpha) / (beta - alpha)


SID_TABLE = {  # modules/dorn.py:11-26 / :45-59
    "kitti": (0.001, 80.0, 71.0),
    "nyu": (0.02, 10.0, 68.0),
    "floorplan3d": (0.0552, 10.0, 68.0),
    "stdepth": (1e-3, 1.0, 68.0),
}


def sid_table_depth(dataset, labels):
    """modules/dorn.py:10-41 (get_depth_sid)."""
    a, b, k = SID_TABLE[dataset]
    a, b, k = torch.tensor(a).float(), torch.tensor(b).float(), torch.tensor(k).float()
    return torch.exp(torch.log(a) + torch.log(b / a) * labels / k).float()


def sid_table_labels(dataset, depth):
    """modules/dorn.py:43-71 (get_labels_sid); returns int32 (truncation). In the reference the
    'stdepth' branch never defines alpha (NameError, :56-63) - the table value is used here."""
    a, b, k = SID_TABLE[dataset]
    a, b, k = torch.tensor(a).float(), torch.tensor(b).float(), torch.tensor(k).float()
    return (k * torch.log(depth / a) / torch.log(b / a)).int()


def ord_loss(P, target):
    """criteria.py:744-787. P [N,K,H,W], target [N,1,H,W] float SID label (broadcast).
    loss = -( sum_{k<=y} ln clamp(P,1e-8,1e8) + sum_{k>y} ln clamp(1-P,1e-8,1e8) ) / (N*H*W)."""
    N, K, H, W = P.shape
    kidx = torch.arange(K, dtype=torch.int32).view(1, K, 1, 1).expand(N, K, H, W)
    m0 = (kidx <= target).detach()
    m1 = (kidx > target).detach()
    s = torch.sum(torch.log(torch.clamp(P[m0], min=1e-8, max=1e8))) \
        + torch.sum(torch.log(torch.clamp(1.0 - P[m1], min=1e-8, max=1e8)))
    return s / (-(N * H * W))


def ordinal_regression_loss(prob, gt, ord_num, alpha, beta, discretization="SID"):
    """criteria.py:789-836. prob [N,2K,H,W] = log-probabilities laid out [K '<=' planes | K '>' planes]
    (concatenated, :817); label = trunc toward zero (:805); mean over gt > 0 pixels (:829-836)."""
    if prob.shape != gt.shape:
        prob = F.interpolate(prob, size=gt.shape[-2:], mode="bilinear", align_corners=True)
    alpha = torch.as_tensor(alpha, dtype=gt.dtype)
    beta = torch.as_tensor(beta, dtype=gt.dtype)
    N, _, H, W = gt.shape
    if discretization == "SID":
        label = ord_num * torch.log(gt / alpha) / torch.log(beta / alpha)
    else:
        label = ord_num * (gt - alpha) / (beta - alpha)
    label = label.long()
    kidx = torch.arange(ord_num).view(1, ord_num, 1, 1).expand(N, ord_num, H, W)
    gtmask = kidx > label
    c0 = torch.ones(N, ord_num, H, W, dtype=prob.dtype)
    c0[gtmask] = 0
    ord_label = torch.cat((c0, 1 - c0), dim=1)
    valid = (gt > 0.).squeeze(1)
    ent = -prob * ord_label
    return torch.sum(ent, dim=1)[valid].mean()
