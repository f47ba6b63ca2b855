"""ORACLE (test infrastructure, not product code): CPU restatement of the other half of VNL's ModelLoss.

  wcel_loss        reference criteria.py:839-863   (WCEL_Loss.__init__ / forward)
  vnl_params       reference modules/vnl.py:160-163 (depth_min_log, depth_bin_interval, wce_loss_weight,
                                                     depth_bin_border)
  depth_to_bins    reference modules/vnl.py:202-217 (VNLModule.depth_to_bins; mutates `depth` in place)
  bins_to_depth    reference modules/vnl.py:219-230 (VNLModule.bins_to_depth)
  model_loss       reference criteria.py:1047-1062  (ModelLoss.forward)

Pinned by tests/golden/wcel_small.npz, produced by the reference's own code: WCEL_Loss/ModelLoss are
imported from criteria.py; depth_to_bins / bins_to_depth live in a LightningModule that cannot be
imported here (pytorch_lightning is absent), so oracle/gen_golden.py compiles the two method bodies
straight from the reference source file (ast) and runs them against a stand-in `self`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def vnl_params(depth_min=0.01, depth_max=1.1, dec_out_c=150):
    """modules/vnl.py:160-163 with the defaults of :341-346."""
    depth_min_log = np.log10(depth_min)
    interval = (np.log10(depth_max) - np.log10(depth_min)) / dec_out_c
    weight = [[np.exp(-0.2 * (i - j) ** 2) for i in range(dec_out_c)] for j in np.arange(dec_out_c)]
    border = np.array([np.log10(depth_min) + interval * (i + 0.5) for i in range(dec_out_c)])
    return {"depth_min": depth_min, "depth_max": depth_max, "dec_out_c": dec_out_c, "depth_min_log": depth_min_log,
            "depth_bin_interval": interval, "wce_loss_weight": weight, "depth_bin_border": border}


def normalised_weight(wce_loss_weight):
    """criteria.py:846-848: rows divided by their sum (numpy float64), then a torch tensor cast to fp32 in forward (:851)."""
    w = np.array(wce_loss_weight, dtype=np.float64)
    w = w / np.sum(w, 1, keepdims=True)
    return torch.from_numpy(w)


def wcel_loss(pred_logit, gt_bins, gt, wce_loss_weight, dec_out_c):
    """criteria.py:850-863. pred_logit [B,C,H,W]; gt_bins integer [B,1,H,W] (C+1 marks padding: its one-hot
    row is all zero); gt [B,1,H,W] only supplies the count of valid pixels (gt > 0)."""
    weight = normalised_weight(wce_loss_weight).to(dtype=pred_logit.dtype)
    classes = torch.arange(dec_out_c, dtype=gt_bins.dtype)
    log_pred = F.log_softmax(pred_logit, 1)
    log_pred = torch.t(torch.transpose(log_pred, 0, 1).reshape(log_pred.size(1), -1))
    one_hot = (gt_bins.reshape(-1, 1) == classes).to(dtype=pred_logit.dtype)
    w = torch.matmul(one_hot, weight)
    valid = torch.sum(gt > 0.).to(dtype=pred_logit.dtype)
    return -1 * torch.sum(w * log_pred) / valid


def depth_to_bins(depth, p):
    """modules/vnl.py:202-217. Returns int32 bins; `depth` is modified in place exactly as the reference does
    (clamped to [depth_min, depth_max], padding restored to -1)."""
    invalid = depth < 0.
    depth[depth < p["depth_min"]] = p["depth_min"]
    depth[depth > p["depth_max"]] = p["depth_max"]
    bins = ((torch.log10(depth) - p["depth_min_log"]) / p["depth_bin_interval"]).to(torch.int)
    bins[invalid] = p["dec_out_c"] + 1
    bins[bins == p["dec_out_c"]] = p["dec_out_c"] - 1
    depth[invalid] = -1.0
    return bins


def bins_to_depth(depth_bin, p):
    """modules/vnl.py:219-230: 10 ** sum_c softmax_c * border_c, [b,c,h,w] -> [b,1,h,w]."""
    x = depth_bin.permute(0, 2, 3, 1)
    border = torch.tensor(p["depth_bin_border"], dtype=torch.float32).to(depth_bin.dtype)
    d = torch.sum(x * border, dim=3, dtype=depth_bin.dtype, keepdim=True)
    return (10 ** d).permute(0, 3, 1, 2)


def model_loss(pred_depth, pred_logit, depth_bins, depth_gt, p, vnl_fn, diff_loss_weight):
    """criteria.py:1054-1062: WCEL + diff_loss_weight * VNL(gt, pred)."""
    return wcel_loss(pred_logit, depth_bins, depth_gt, p["wce_loss_weight"], p["dec_out_c"]) + \
        diff_loss_weight * vnl_fn(depth_gt, pred_depth)
