"""ORACLE (test infrastructure, not product code): CPU restatement of the layered-depth base criterion.

  BaseModule.setup_criterion -> _loss      reference modules/base_module.py:124-208

restricted to the masked-reduction terms ('silma', 'silms', 'mse', 'mae', 'fbdivergence'); the SSIM and
compositing terms belong to stdepth_utils.py (out of scope). Pinned by tests/golden/stdepth_small.npz, produced by
the reference's own closure (compiled from modules/base_module.py by `ast`, because the module needs
pytorch_lightning) with the reference's own criteria.silog_loss.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import losses as olosses


def stdepth_loss(pred, targ, rgba, loss_name, variance_focus=0.85, depth_w=1.0, fbdiv_w=1.0, single_layer=True):
    """Returns (loss, loss_dict). pred/targ [B,C,H,W], rgba [B,4+,H,W]."""
    def silog(p, t):                                                     # base_module.py:125-127
        return torch.nan_to_num(olosses.silog(p, t, variance_focus))
    mask1 = rgba[:, [3]] > 0.0                                           # :133
    mask8 = mask1.expand(-1, 8, -1, -1)
    maskN = mask1.expand(-1, targ.size(1), -1, -1)
    depth_idx = (slice(None), slice(8, 10)) if single_layer else (slice(None), slice(16, 20))
    maskD = targ[depth_idx] > 0.0
    d = {}
    if "silma" in loss_name:                                             # :156-158
        d["depth_silog"] = depth_w * torch.nan_to_num(silog(pred[depth_idx][maskD], targ[depth_idx][maskD]))
        d["color_mae"] = F.l1_loss(pred[:, :8][mask8], targ[:, :8][mask8])
    if "silms" in loss_name:                                             # :159-161
        d["depth_silog"] = depth_w * torch.nan_to_num(silog(pred[depth_idx][maskD], targ[depth_idx][maskD]))
        d["color_mse"] = F.mse_loss(pred[:, :8][mask8], targ[:, :8][mask8])
    if "mse" in loss_name:                                               # :162-164
        d["all_mse"] = F.mse_loss(pred[maskN], targ[maskN])
        d["all_mse"] = d["all_mse"] + depth_w * F.mse_loss(pred[depth_idx][maskD], targ[depth_idx][maskD])
    if "mae" in loss_name:                                               # :165-167
        d["all_mae"] = F.l1_loss(pred[maskN], targ[maskN])
        d["all_mae"] = d["all_mae"] + depth_w * F.l1_loss(pred[depth_idx][maskD], targ[depth_idx][maskD])
    if "fbdivergence" in loss_name:                                      # :184-194
        fpbg = (torch.linalg.vector_norm(pred[:, :3], dim=1, keepdim=True) *
                torch.linalg.vector_norm(targ[:, 4:7], dim=1, keepdim=True)) + 1e-3
        fgbp = (torch.linalg.vector_norm(pred[:, 4:7], dim=1, keepdim=True) *
                torch.linalg.vector_norm(targ[:, :3], dim=1, keepdim=True)) + 1e-3
        fb = ((pred[:, :3] * targ[:, 4:7] / fpbg).sum(dim=1) + (pred[:, 4:7] * targ[:, :3] / fgbp).sum(dim=1))[mask1.squeeze(1)]
        d["fb_divergence"] = fbdiv_w * fb.mean()
    loss = torch.stack(list(d.values())).sum()                           # :196
    return loss, d
