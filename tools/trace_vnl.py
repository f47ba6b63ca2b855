#!/usr/bin/env python
"""Per-CTA phase timeline of the VNL kernel at config C4 (mde_debug_set_trace): us since the first CTA started."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth
lib = _lib.load(); dev = torch.device("cuda", 0)
gt, pred, trip = synth.vnl_inputs((8, 1, 385, 385), 104, device=dev)
B, n_trip = 8, trip.shape[1]
ws = _lib.workspace(dev, B)
scratch = torch.empty(int(lib.mde_vnl_scratch_bytes(B, n_trip, 385, 385)), dtype=torch.uint8, device=dev)
loss_t = torch.empty((), device=dev); grad = torch.empty_like(pred)
trace = torch.zeros(296 * 8, dtype=torch.int64, device=dev)
def run(with_grad=True):
    _lib.check(lib.mde_vnl_loss(_lib.ptr(gt), _lib.ptr(pred), 0, _lib.ptr(trip), B, 385, 385, n_trip, 519.0, 519.0, 1, 1.0, _lib.ptr(ws), _lib.ptr(scratch),
                                _lib.ptr(loss_t), None, _lib.ptr(grad) if with_grad else None, _lib.stream_ptr(dev)))
names = ["start", "zeroed+sync", "phase1 loop done", "phase1 sync", "radix select done", "final loop done", "exit"]
for wg in (True, False):
    for _ in range(20): run(wg)
    torch.cuda.synchronize()
    _lib.check(lib.mde_debug_set_trace(_lib.ptr(trace)))
    rows = []
    for rep in range(4):
        trace.zero_(); torch.cuda.synchronize()
        run(wg); torch.cuda.synchronize()
        t = trace.view(296, 8).cpu().double()
        t0 = t[:, 0].min()
        rel = (t[:, :7] - t0) / 1e3
        rows.append({n: [round(float(rel[:, k].min()), 1), round(float(rel[:, k].median()), 1), round(float(rel[:, k].max()), 1)] for k, n in enumerate(names)})
    _lib.check(lib.mde_debug_set_trace(None))
    print(json.dumps({"grad": wg, "min_median_max_us": rows[-1]}))
