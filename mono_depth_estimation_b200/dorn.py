"""DORN ordinal head on B200: drop-ins for the reference's OrdinalRegressionLayer
(network/Dorn.py:288-321) and DORNModule.label_to_depth / depth_to_label (modules/dorn.py:95-107),
plus the fused supervision step `DornOrdinalHead` (logits -> decode, depth, loss, grad in one pass).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

__all__ = ["OrdinalRegressionLayer", "label_to_depth", "depth_to_label", "get_depth_sid", "get_labels_sid",
           "DornOrdinalHead", "dorn_fused"]

_DISC = {"SID": _lib.DISC_SID, "UD": _lib.DISC_UD}


class _OrdinalLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        dev = _lib.require_cuda(x)
        xc = x.detach().contiguous()
        N, C, H, W = xc.shape
        K = C // 2
        with torch.cuda.device(dev):
            prob = torch.empty((N, K, H, W), dtype=torch.float32, device=dev)
            decode = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev)
            _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(xc), _lib.dtype_code(xc), N, K, H * W, _lib.ptr(prob),
                                                 _lib.ptr(decode), _lib.stream_ptr(dev)))
        ctx.save_for_backward(xc)
        ctx.mark_non_differentiable(decode)
        if prob.dtype != x.dtype:
            prob = prob.to(x.dtype)
        return decode, prob

    @staticmethod
    def backward(ctx, _gdecode, gprob):
        lib = _lib.load()
        (xc,) = ctx.saved_tensors
        dev = xc.device
        N, C, H, W = xc.shape
        gp = gprob.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            gx = torch.empty_like(xc)
            _lib.check(lib.mde_ordinal_layer_bwd(_lib.ptr(xc), _lib.dtype_code(xc), _lib.ptr(gp), N, C // 2, H * W,
                                                 _lib.ptr(gx), _lib.stream_ptr(dev)))
        return gx


class OrdinalRegressionLayer(nn.Module):
    """reference network/Dorn.py:288-321: x [N,2K,H,W] -> (decode int64 [N,1,H,W], P [N,K,H,W])."""

    def __init__(self):
        super(OrdinalRegressionLayer, self).__init__()

    def forward(self, x):
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        return _OrdinalLayerFn.apply(x)


def _f(v):
    return float(v.item()) if torch.is_tensor(v) else float(v)


def label_to_depth(label, alpha, beta, ord_num, discretization="SID"):
    """reference modules/dorn.py:95-100; label int64 (decode) or float."""
    lib = _lib.load()
    dev = _lib.require_cuda(label)
    K = int(ord_num.item()) if torch.is_tensor(ord_num) else int(ord_num)
    lc = label.detach().contiguous()
    with torch.cuda.device(dev):
        out = torch.empty(lc.shape, dtype=torch.float32, device=dev)
        if lc.dtype == torch.int64:
            fn = lib.mde_label_to_depth_i64
        else:
            lc = lc.to(torch.float32)
            fn = lib.mde_label_to_depth_f32
        _lib.check(fn(_lib.ptr(lc), lc.numel(), _f(alpha), _f(beta), K, _DISC[discretization], _lib.ptr(out),
                      _lib.stream_ptr(dev)))
    return out


def depth_to_label(depth, alpha, beta, ord_num, discretization="SID"):
    """reference modules/dorn.py:102-107: float label (not floored); depth 0 -> -inf."""
    lib = _lib.load()
    dev = _lib.require_cuda(depth)
    K = int(ord_num.item()) if torch.is_tensor(ord_num) else int(ord_num)
    dc = depth.detach().to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        out = torch.empty_like(dc)
        _lib.check(lib.mde_depth_to_label(_lib.ptr(dc), dc.numel(), _f(alpha), _f(beta), K, _DISC[discretization],
                                          _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


_SID_TABLE = {"kitti": (0.001, 80.0, 71), "nyu": (0.02, 10.0, 68), "floorplan3d": (0.0552, 10.0, 68),
              "stdepth": (1e-3, 1.0, 68)}  # reference modules/dorn.py:11-26


def get_depth_sid(dataset, labels):
    """reference modules/dorn.py:10-41."""
    a, b, k = _SID_TABLE[dataset]
    return label_to_depth(labels, a, b, k, "SID")


def get_labels_sid(dataset, depth):
    """reference modules/dorn.py:43-71 (returns int32, truncated toward zero)."""
    a, b, k = _SID_TABLE[dataset]
    return depth_to_label(depth, a, b, k, "SID").int()


class _DornFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gt, alpha, beta, K, disc, want_prob):
        lib = _lib.load()
        dev = _lib.require_cuda(x, gt)
        from .criteria import _compute_copy
        xc = _compute_copy(x)     # half-precision logits are widened: the stashed gradient must not underflow
        N, C, H, W = xc.shape
        assert C == 2 * K, "logits must have 2*ord_num channels"
        gtc = gt.detach().to(torch.float32).reshape(N, H * W).contiguous()
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(dev):
            ws = _lib.workspace(dev, 1)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            decode = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev)
            depth = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
            prob = torch.empty((N, K, H, W), dtype=torch.float32, device=dev) if want_prob else None
            gx = torch.empty_like(xc) if need_grad else None
            _lib.check(lib.mde_dorn_fused(_lib.ptr(xc), _lib.dtype_code(xc), _lib.ptr(gtc), N, K, H * W, alpha, beta,
                                          disc, 1.0, _lib.ptr(ws), _lib.ptr(loss), _lib.ptr(prob), _lib.ptr(decode),
                                          _lib.ptr(depth), _lib.ptr(gx), _lib.stream_ptr(dev)))
        ctx.gx = gx
        ctx.x_dtype = x.dtype
        ctx.used = False
        if prob is None:
            prob = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(decode, depth, prob)   # ONE call: a second call replaces the first set
        return loss, decode, depth, prob

    @staticmethod
    def backward(ctx, gloss, *_):
        if ctx.gx is None:
            return (None,) * 7
        if ctx.used:
            raise RuntimeError("the fused DORN gradient was already consumed; run the forward again")
        ctx.used = True
        from .criteria import _scale_grad
        g = _scale_grad(ctx.gx, gloss)
        ctx.gx = None
        if g.dtype != ctx.x_dtype:
            g = g.to(ctx.x_dtype)
        return (g,) + (None,) * 6


def dorn_fused(logits, gt_depth, ord_num, alpha, beta, discretization="SID", want_prob=False):
    """(loss, decode int64 [N,1,H,W], depth fp32 [N,1,H,W], P or empty) in one pass over the logits."""
    return _DornFusedFn.apply(logits, gt_depth, _f(alpha), _f(beta), int(ord_num), _DISC[discretization], want_prob)


class DornOrdinalHead(nn.Module):
    """Fused replacement for the DORN training-step tail (reference modules/dorn.py:160-163):

        pred_d, pred_ord = OrdinalRegressionLayer()(logits); y_hat = label_to_depth(pred_d)
        y_sid = depth_to_label(y); loss = ordLoss()(pred_ord, y_sid)

    forward(logits, gt_depth) -> (loss, y_hat depth, decode)."""

    def __init__(self, ord_num=68, alpha=0.001, beta=1.0, discretization="SID"):
        super().__init__()
        self.ord_num, self.alpha, self.beta, self.discretization = int(ord_num), float(alpha), float(beta), discretization

    def forward(self, logits, gt_depth):
        loss, decode, depth, _ = dorn_fused(logits, gt_depth, self.ord_num, self.alpha, self.beta, self.discretization)
        return loss, depth, decode
