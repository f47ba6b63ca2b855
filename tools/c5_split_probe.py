#!/usr/bin/env python
"""How much of the C5 metrics launch is per-image work (flushes + the last CTA's finaliser)? Same 200.9 M pixels presented as
654 images (C5) and as 6 / 109 images: python tools/c5_split_probe.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth
lib = _lib.load(); dev = torch.device("cuda", 0)
pr, gt = synth.depth_pair((654, 1, 480, 640), 105, device=dev)
npx = 654 * 480 * 640
ws = _lib.workspace(dev, 654)
o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
out = {}
for keys, mflags in (("7keys", _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_RSQ), ("10keys", 0)):
    for n_img in (654, 109, 6, 82):
        hw = npx // n_img if n_img != 82 else 480 * 640
        f = lambda: _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(gt), n_img, hw, mflags, _lib.ptr(ws), _lib.ptr(o64), _lib.ptr(o32), None, None, _lib.stream_ptr(dev)))
        for _ in range(3): f()
        torch.cuda.synchronize()
        ts = []
        for r in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); f(); b.record(); torch.cuda.synchronize()
            ts.append(1e3 * a.elapsed_time(b) / 2)
        ts.sort()
        out["%s_%dimg_us" % (keys, n_img)] = round(ts[len(ts) // 2], 2)
print(json.dumps(out))
