"""The classification half of VNL's ModelLoss on B200, same signatures as the reference:

  WCEL_Loss(args).forward(pred_logit, gt_bins, gt)     reference criteria.py:839-863
  depth_to_bins(depth, ...)                            reference modules/vnl.py:202-217 (VNLModule.depth_to_bins)
  bins_to_depth(depth_bin, ...)                        reference modules/vnl.py:219-230 (VNLModule.bins_to_depth)
  vnl_params(...)                                      reference modules/vnl.py:160-163 (the derived constants)

`args` needs `.wce_loss_weight` ([C][C], un-normalised) and `.dec_out_c`, as in the reference. Kernels: csrc/wcel.cu
through the C ABI (mde_wcel_loss, mde_depth_to_bins, mde_bins_to_depth, mde_bins_to_depth_bwd). No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .criteria import _fused_apply, _compute_copy

__all__ = ["WCEL_Loss", "depth_to_bins", "bins_to_depth", "vnl_params", "VNLBins"]


def vnl_params(depth_min=0.01, depth_max=1.1, dec_out_c=150):
    """The constants VNLModule derives in its constructor (modules/vnl.py:160-163; defaults :341-346)."""
    depth_min_log = np.log10(depth_min)
    interval = (np.log10(depth_max) - np.log10(depth_min)) / dec_out_c
    weight = [[np.exp(-0.2 * (i - j) ** 2) for i in range(dec_out_c)] for j in np.arange(dec_out_c)]
    border = np.array([np.log10(depth_min) + interval * (i + 0.5) for i in range(dec_out_c)])
    return {"depth_min": depth_min, "depth_max": depth_max, "dec_out_c": dec_out_c, "depth_min_log": depth_min_log,
            "depth_bin_interval": interval, "wce_loss_weight": weight, "depth_bin_border": border}


class WCEL_Loss(nn.Module):
    """Weighted cross-entropy (reference criteria.py:839-863), forward and backward in one pass over the logits."""

    def __init__(self, args):
        super(WCEL_Loss, self).__init__()
        self.args = args
        w = np.array(self.args.wce_loss_weight, dtype=np.float64)
        w = w / np.sum(w, 1, keepdims=True)                       # criteria.py:846-847
        self.weight = torch.from_numpy(w)                         # plain attribute, not in state_dict (as the reference)
        self._dev = {}

    def _tables(self, dev):
        t = self._dev.get(dev)
        if t is None:
            w32 = self.weight.to(device=dev, dtype=torch.float32).contiguous()   # criteria.py:851
            t = (w32, w32.double().sum(1).float().contiguous())
            self._dev[dev] = t
        return t

    def forward(self, pred_logit, gt_bins, gt):
        lib = _lib.load()
        dev = _lib.require_cuda(pred_logit, gt_bins, gt)
        C = int(self.args.dec_out_c)
        assert pred_logit.dim() == 4 and int(pred_logit.shape[1]) == C, "pred_logit must be [B, dec_out_c, H, W]"
        B, _, H, W = pred_logit.shape
        assert gt_bins.numel() == B * H * W and gt.numel() == B * H * W, "gt_bins / gt must hold one value per pixel"
        weight, rowsum = self._tables(dev)
        bins = gt_bins.detach().to(torch.int32).contiguous()
        gtf = gt.detach().to(torch.float32).contiguous()

        def launch(p, need_grad):
            pc = _compute_copy(p)     # half-precision logits are widened: the stashed gradient must not underflow (criteria._compute_copy)
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, 1)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                _lib.check(lib.mde_wcel_loss(_lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(bins), _lib.ptr(gtf),
                                             _lib.ptr(weight), _lib.ptr(rowsum), B, C, H * W, 1.0, _lib.ptr(ws),
                                             _lib.ptr(loss), _lib.ptr(grad), _lib.stream_ptr(dev)))
            return loss, grad

        return _fused_apply(pred_logit, launch)


def depth_to_bins(depth, depth_min=0.01, depth_max=1.1, dec_out_c=150, depth_min_log=None, depth_bin_interval=None):
    """VNLModule.depth_to_bins (modules/vnl.py:202-217): int32 bins, padding (depth < 0) marked dec_out_c + 1;
    `depth` is modified IN PLACE exactly as the reference does (clamped to [depth_min, depth_max], padding = -1)."""
    lib = _lib.load()
    dev = _lib.require_cuda(depth)
    if depth.dtype != torch.float32 or not depth.is_contiguous():
        raise TypeError("depth_to_bins works in place on a contiguous fp32 tensor (as the reference mutates its argument)")
    if depth_min_log is None:
        depth_min_log = np.log10(depth_min)
    if depth_bin_interval is None:
        depth_bin_interval = (np.log10(depth_max) - np.log10(depth_min)) / dec_out_c
    with torch.cuda.device(dev):
        bins = torch.empty(depth.shape, dtype=torch.int32, device=dev)
        _lib.check(lib.mde_depth_to_bins(_lib.ptr(depth), depth.numel(), float(depth_min), float(depth_max),
                                         float(depth_min_log), float(depth_bin_interval), int(dec_out_c), _lib.ptr(bins),
                                         _lib.stream_ptr(dev)))
    return bins


class _BinsToDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, border):
        lib = _lib.load()
        dev = _lib.require_cuda(prob, border)
        B, C, H, W = prob.shape
        pc = prob.detach().contiguous()
        with torch.cuda.device(dev):
            depth = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
            _lib.check(lib.mde_bins_to_depth(_lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(border), B, C, H * W,
                                             _lib.ptr(depth), _lib.stream_ptr(dev)))
        ctx.save_for_backward(depth, border)
        ctx.meta = (B, C, H, W, pc.dtype)
        return depth

    @staticmethod
    def backward(ctx, grad_depth):
        lib = _lib.load()
        depth, border = ctx.saved_tensors
        B, C, H, W, dt = ctx.meta
        dev = depth.device
        g = grad_depth.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            gp = torch.empty((B, C, H, W), dtype=dt, device=dev)
            _lib.check(lib.mde_bins_to_depth_bwd(_lib.ptr(depth), _lib.ptr(g), _lib.ptr(border), B, C, H * W,
                                                 _lib.dtype_code(gp), _lib.ptr(gp), _lib.stream_ptr(dev)))
        return gp, None


def bins_to_depth(depth_bin, depth_bin_border):
    """VNLModule.bins_to_depth (modules/vnl.py:219-230): [b,c,h,w] class probabilities -> [b,1,h,w] fp32 depth
    10 ** sum_c p_c * border_c, differentiable w.r.t. depth_bin."""
    dev = _lib.require_cuda(depth_bin)
    border = torch.as_tensor(np.asarray(depth_bin_border), dtype=torch.float32).to(dev).contiguous()
    assert depth_bin.dim() == 4 and int(depth_bin.shape[1]) == border.numel(), "depth_bin must be [B, C, H, W]"
    return _BinsToDepthFn.apply(depth_bin, border)


class VNLBins:
    """The two bin maps bound to one set of VNL constants, callable like the reference's methods."""

    def __init__(self, depth_min=0.01, depth_max=1.1, dec_out_c=150):
        self.p = vnl_params(depth_min, depth_max, dec_out_c)

    def depth_to_bins(self, depth):
        p = self.p
        return depth_to_bins(depth, p["depth_min"], p["depth_max"], p["dec_out_c"], p["depth_min_log"], p["depth_bin_interval"])

    def bins_to_depth(self, depth_bin):
        return bins_to_depth(depth_bin, self.p["depth_bin_border"])
