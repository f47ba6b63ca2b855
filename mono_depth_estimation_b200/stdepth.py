"""The layered-depth ("stdepth") base criterion: host-side mirror of BaseModule.setup_criterion
(reference modules/base_module.py:124-208), the criterion the registered methods `bts` and `laina` inherit
(call sites modules/bts.py:106,114,130 and modules/laina.py:25,33,42).

    criterion = setup_criterion(method, single_layer=True)           # method.{loss, variance_focus, depth_loss_weight,
    loss, = criterion(pred, targ, rgba)                              #         comp_loss_weight, fbdiv_loss_weight, ssim_loss_weight}
    loss, pred_full, loss_dict = criterion(pred, targ, rgba, return_composited=True, return_loss_dict=True)

All masked-reduction terms ('silma', 'silms', 'mse', 'mae', 'fbdivergence' in `method.loss`) run in ONE cooperative
launch, forward and backward (C ABI mde_stdepth_loss). The SSIM and compositing TERMS ('ssim', 'composite' in
`method.loss`) live in the reference's stdepth_utils.py, which is out of scope: they raise
NotImplementedError. `return_composited=True` (visualisation, base_module.py:142-155) calls the compositing functions
the caller hands in (`composite_layers`, `depth_sort` of the reference's stdepth_utils) - this package does not ship them.
"""
from __future__ import annotations

import torch

from . import _lib
from .criteria import _fused_apply

__all__ = ["setup_criterion", "TERM_FLAGS"]

TERM_FLAGS = {"depth_silog": 1, "color_mae": 2, "color_mse": 4, "all_mse": 8, "all_mae": 16, "fb_divergence": 32}


def _flags_of(loss_name: str) -> int:
    """Substring tests in the reference's order (base_module.py:156-194)."""
    if "ssim" in loss_name or "composite" in loss_name:
        raise NotImplementedError("loss %r: the SSIM / compositing terms (base_module.py:168-183, stdepth_utils.py) are "
                                  "out of scope of this package" % (loss_name,))
    f = 0
    if "silma" in loss_name:
        f |= TERM_FLAGS["depth_silog"] | TERM_FLAGS["color_mae"]
    if "silms" in loss_name:
        f |= TERM_FLAGS["depth_silog"] | TERM_FLAGS["color_mse"]
    if "silma" in loss_name and "silms" in loss_name:
        raise NotImplementedError("'silma' and 'silms' together: the second overwrites 'depth_silog' in the reference's dict")
    if "mse" in loss_name:
        f |= TERM_FLAGS["all_mse"]
    if "mae" in loss_name:
        f |= TERM_FLAGS["all_mae"]
    if "fbdivergence" in loss_name:
        f |= TERM_FLAGS["fb_divergence"]
    if f == 0:
        raise RuntimeError("loss %r selects no term (the reference's torch.stack of an empty list raises too)" % (loss_name,))
    return f


def setup_criterion(method, single_layer=True, composite_layers=None, depth_sort=None):
    """Returns `_loss(pred, targ, rgba, return_composited=False, return_loss_dict=False) -> tuple`, as
    BaseModule.setup_criterion does (base_module.py:124-208)."""
    flags = _flags_of(method.loss)
    depth_w = float(method.depth_loss_weight)
    fbdiv_w = float(getattr(method, "fbdiv_loss_weight", 1.0))
    lam = float(method.variance_focus)
    names = [k for k in ("depth_silog", "color_mae", "color_mse", "all_mse", "all_mae", "fb_divergence") if flags & TERM_FLAGS[k]]
    slot = {"depth_silog": 1, "color_mae": 2, "color_mse": 2, "all_mse": 3, "all_mae": 4, "fb_divergence": 5}

    def _loss(pred, targ, rgba, return_composited=False, return_loss_dict=False):
        lib = _lib.load()
        dev = _lib.require_cuda(pred, targ, rgba)
        assert pred.dim() == 4 and pred.shape == targ.shape, "inconsistent dimensions"
        B, C, H, W = (int(v) for v in pred.shape)
        if C != (10 if single_layer else 20):
            raise ValueError("pred/targ need %d channels (single_layer=%s, base_module.py:137)" % (10 if single_layer else 20, single_layer))
        assert rgba.dim() == 4 and rgba.shape[0] == B and rgba.shape[1] >= 4 and tuple(rgba.shape[2:]) == (H, W), "rgba must be [B, >=4, H, W]"
        t = targ.detach().to(torch.float32).contiguous()
        x = rgba.detach().to(torch.float32).contiguous()
        box = {}

        def launch(p, need_grad):
            pc = p.detach()
            if pc.dtype != torch.float32:
                pc = pc.float()       # also fp16 / bf16: the stashed gradient must not underflow (criteria._compute_copy)
            pc = pc.contiguous()
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, B)
                out8 = torch.empty(8, dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                _lib.check(lib.mde_stdepth_loss(_lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(t), _lib.ptr(x), int(x.shape[1]), B, C,
                                                H * W, flags, depth_w, fbdiv_w, lam, 1.0, _lib.ptr(ws), _lib.ptr(out8),
                                                _lib.ptr(grad), _lib.stream_ptr(dev)))
            box["out8"] = out8
            return out8[0], grad

        loss = _fused_apply(pred, launch)
        ret = [loss]
        if return_composited:                                         # base_module.py:142-155
            if composite_layers is None or (not single_layer and depth_sort is None):
                raise NotImplementedError("return_composited=True needs the reference's stdepth_utils.composite_layers "
                                          "(and depth_sort for three layers) passed to setup_criterion")
            with torch.no_grad():
                if single_layer:
                    pred_full = composite_layers(torch.stack([pred[:, :4], pred[:, 4:8]], dim=1))
                else:
                    l1 = torch.cat([pred[:, :4], pred[:, [16]]], dim=1)
                    l2 = torch.cat([pred[:, 4:8], pred[:, [17]]], dim=1)
                    l3 = torch.cat([pred[:, 8:12], pred[:, [18]]], dim=1)
                    back = pred[:, 12:16].unsqueeze(1)
                    sorted_layers = depth_sort(torch.stack([l1, l2, l3], dim=1))[:, :, :4]
                    pred_full = composite_layers(torch.cat([sorted_layers, back], dim=1))
            ret.append(pred_full)
        if return_loss_dict:
            o = box["out8"]
            ret.append({k: o[slot[k]] for k in names})
        return tuple(ret)

    return _loss
