#!/usr/bin/env python
"""Per-CTA phase timeline of the fused SILog(+metrics) kernel at config C2 (uses mde_debug_set_trace)."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth
lib = _lib.load(); dev = torch.device("cuda", 0)
shape = (16, 1, 480, 640)
ring = [synth.depth_pair(shape, 500 + i, device=dev) for i in range(8)]
ws = _lib.workspace(dev, 16)
loss_t = torch.empty((), device=dev); grad_t = torch.empty(shape, device=dev)
o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
lp = _lib.LossParams(0.85, 1e-9, 1, 1); mflags = _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_RSQ
trace = torch.zeros(296 * 8, dtype=torch.int64, device=dev)
def run(i, fused):
    pr, g = ring[i % 8]
    if fused:
        _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, 16, 480, 640, C.byref(lp), 1.0, mflags,
                   _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grad_t), _lib.ptr(o64), _lib.ptr(o32), _lib.stream_ptr(dev)))
    else:
        _lib.check(lib.mde_masked_loss(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, 16, 480, 640, C.byref(lp), 1.0,
                   _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grad_t), _lib.stream_ptr(dev)))
for fused in (True, False):
    for i in range(5): run(i, fused)
    torch.cuda.synchronize()
    _lib.check(lib.mde_debug_set_trace(_lib.ptr(trace)))
    stats = []
    for rep in range(5):
        trace.zero_(); torch.cuda.synchronize()
        run(5 + rep, fused); torch.cuda.synchronize()
        t = trace.view(296, 8).cpu().double()
        t0 = t[:, 0].min()
        rel = (t[:, :6] - t0) / 1e3   # us
        stats.append({"start_max": float(rel[:, 0].max()), "loopdone_min": float(rel[:, 1].min()), "loopdone_mean": float(rel[:, 1].mean()),
                      "loopdone_max": float(rel[:, 1].max()), "published_max": float(rel[:, 2].max()), "barrier_exit_min": float(rel[:, 3].min()),
                      "barrier_exit_max": float(rel[:, 3].max()), "totals_max": float(rel[:, 4].max()), "end_min": float(rel[:, 5].min()),
                      "end_mean": float(rel[:, 5].mean()), "end_max": float(rel[:, 5].max())})
    _lib.check(lib.mde_debug_set_trace(None))
    print(json.dumps({"kernel": "silog+metrics fused" if fused else "silog", "us_since_first_cta_start": stats[-3:]}))
