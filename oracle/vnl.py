"""ORACLE (test infrastructure, not product code): CPU restatement of the virtual-normal loss.

Follows reference criteria.py:866-1045 (VNL_Loss) in its per-triplet form,
with the SUPPLIED triplet tensor in place of the reference's np.random sampling
(select_index, criteria.py:912-932 - the same triplets are used for every image of the batch,
:948-950). Operation order mirrors the reference so an fp32 run reproduces it to rounding.
"""
from __future__ import annotations

import torch


def back_project(depth, fx, fy):
    """criteria.py:890-910: u0 = W//2, v0 = H//2 (integer halves); x=(u-u0)|d|/fx, y=(v-v0)|d|/fy, z=d.
    depth [B,1,H,W] -> [B,H,W,3]."""
    B, _, H, W = depth.shape
    dt = depth.dtype
    u = torch.arange(W, dtype=torch.float32).to(dt).view(1, 1, 1, W) - float(W // 2)
    v = torch.arange(H, dtype=torch.float32).to(dt).view(1, 1, H, 1) - float(H // 2)
    fx_t = torch.tensor([fx], dtype=torch.float32).to(dt)
    fy_t = torch.tensor([fy], dtype=torch.float32).to(dt)
    x = u * torch.abs(depth) / fx_t
    y = v * torch.abs(depth) / fy_t
    return torch.cat([x, y, depth], 1).permute(0, 2, 3, 1)


def _groups(pw, trip, W):
    """criteria.py:934-953: [B,N,3(xyz),3(p1,p2,p3)] gathered at the three flat indices."""
    pts = []
    for m in range(3):
        yy = torch.div(trip[m], W, rounding_mode="floor")
        xx = trip[m] % W
        pts.append(pw[:, yy, xx, :].unsqueeze(3))
    return torch.cat(pts, 3)


def triplet_mask(gt_groups, delta_cos=0.867, dx=0.005, dy=0.005, dz=0.005, delta_z=0.0001):
    """criteria.py:955-988 (filter_mask). gt_groups [B,N,3,3] -> bool [B,N]."""
    pw = gt_groups
    d12 = pw[..., 1] - pw[..., 0]
    d13 = pw[..., 2] - pw[..., 0]
    d23 = pw[..., 2] - pw[..., 1]
    diff = torch.stack([d12, d13, d23], 3)               # [B,N,3(xyz),3(diff)]
    B, N = diff.shape[:2]
    key = diff.reshape(B * N, 3, 3)                      # [bn, xyz, diff]
    query = key.permute(0, 2, 1)                         # [bn, diff, xyz]
    qn = query.norm(2, dim=2)
    nm = torch.bmm(qn.view(B * N, 3, 1), qn.view(B * N, 1, 3))
    energy = torch.bmm(query, key)
    ne = (energy / (nm + 1e-8)).view(B * N, -1)
    mask_cos = (torch.sum((ne > delta_cos) | (ne < -delta_cos), 1) > 3).view(B, N)
    mask_pad = torch.sum(pw[:, :, 2, :] > delta_z, 2) == 3
    mx = torch.sum(torch.abs(diff[:, :, 0, :]) < dx, 2) > 0
    my = torch.sum(torch.abs(diff[:, :, 1, :]) < dy, 2) > 0
    mz = torch.sum(torch.abs(diff[:, :, 2, :]) < dz, 2) > 0
    ignore = (mx & my & mz) | mask_cos
    return mask_pad & ~ignore


def per_triplet_loss(gt_groups, pred_groups):
    """criteria.py:1019-1040 on [M,3,3] valid groups -> [M] per-triplet L1 between unit normals."""
    g12 = gt_groups[:, :, 1] - gt_groups[:, :, 0]
    g13 = gt_groups[:, :, 2] - gt_groups[:, :, 0]
    q12 = pred_groups[:, :, 1] - pred_groups[:, :, 0]
    q13 = pred_groups[:, :, 2] - pred_groups[:, :, 0]
    gn = torch.linalg.cross(g12, g13, dim=1)
    qn = torch.linalg.cross(q12, q13, dim=1)
    qnorm = torch.norm(qn, 2, dim=1, keepdim=True)
    gnorm = torch.norm(gn, 2, dim=1, keepdim=True)
    qnorm = qnorm + (qnorm == 0.0).to(qnorm.dtype) * 0.01
    gnorm = gnorm + (gnorm == 0.0).to(gnorm.dtype) * 0.01
    return torch.abs(gn / gnorm - qn / qnorm).sum(dim=1)


def vnl_loss(gt_depth, pred_depth, trip, fx, fy, select=True, return_parts=False):
    """criteria.py:990-1045 with supplied triplets `trip` int64 [3,N] (flat indices y*W+x)."""
    B, _, H, W = gt_depth.shape
    pw_gt = back_project(gt_depth, fx, fy)
    pw_pr = back_project(pred_depth, fx, fy)
    gt_groups = _groups(pw_gt, trip, W)
    mask = triplet_mask(gt_groups)
    pr_groups = _groups(pw_pr, trip, W)
    # quirk criteria.py:1004: the [B,N,3(point)] boolean indexes dims [B,N,3(xyz)] -> if point j
    # has z == 0, coordinate ROW j of all three points becomes 1e-4 (gradient is cut there).
    zmask = pr_groups[:, :, 2, :] == 0
    pr_groups = pr_groups.clone()
    pr_groups[zmask] = 0.0001
    gt_valid = gt_groups[mask]       # [M,3,3] in (b, n) row-major order, as the reference pools them
    pr_valid = pr_groups[mask]
    per = per_triplet_loss(gt_valid, pr_valid)
    M = per.shape[0]
    if select:
        srt, _ = torch.sort(per, dim=0, descending=False)
        kept = srt[int(M * 0.25):]
    else:
        kept = per
    loss = torch.mean(kept)
    if return_parts:
        return loss, per.detach(), mask
    return loss
