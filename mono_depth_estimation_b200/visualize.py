"""Host-side mirror of the one per-pixel function of the reference's visualize.py that sits next to the hot path:

    colored_depthmap(depth, d_min=None, d_max=None, do_mapping=True)        reference visualize.py:8-17

(the panels that merge_into_row / save_visualization assemble from it stay matplotlib / cv2 code of the caller).
The map is normalised, quantised and coloured on the device (C ABI mde_colored_depthmap); a numpy input gives a numpy
result of shape [H, W, 3] (BGR, as cv2.applyColorMap) or [H, W], a CUDA tensor gives a CUDA uint8 tensor.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["colored_depthmap"]


def colored_depthmap(depth, d_min=None, d_max=None, do_mapping=True):
    lib = _lib.load()
    was_numpy = isinstance(depth, np.ndarray)
    if was_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("colored_depthmap needs a CUDA device; there is no CPU fallback")
        depth = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).cuda()
    dev = _lib.require_cuda(depth)
    d = depth.detach().to(torch.float32).contiguous()
    n = d.numel()
    # visualize.py:9-12 takes each bound from the data when it is None
    auto = d_min is None or d_max is None
    with torch.cuda.device(dev):
        if auto and not (d_min is None and d_max is None):
            lo, hi = (float(d.min()) if d_min is None else float(d_min)), (float(d.max()) if d_max is None else float(d_max))
            auto, d_min, d_max = False, lo, hi
        out = torch.empty(tuple(d.shape) + ((3,) if do_mapping else ()), dtype=torch.uint8, device=dev)
        scratch = torch.empty(4, dtype=torch.int32, device=dev) if auto else None
        _lib.check(lib.mde_colored_depthmap(_lib.ptr(d), n, 0.0 if auto else float(np.float32(d_min)),
                                            0.0 if auto else float(np.float32(d_max)), 1 if auto else 0, 1 if do_mapping else 0,
                                            _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr(dev)))
    return out.cpu().numpy() if was_numpy else out
