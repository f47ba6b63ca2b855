#!/usr/bin/env python
"""Run ONE hot-path call a few times (for ncu / compute-sanitizer).  python tools/run_one.py <name> [reps]
names: c5_metrics, c5_metrics10, c2_metrics, c2_fused, dorn_fused, dorn_decode, ord_loss, vnl, c1_berhu, c1_eigen, pointcloud, wcel"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth, dorn
lib = _lib.load(); dev = torch.device("cuda", 0)
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sp = lambda: _lib.stream_ptr(dev)
loss_t = torch.empty((), device=dev)
lp = _lib.LossParams(0.85, 1e-9, 1, 1)
mflags = _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_RSQ
if name in ("c5_metrics", "c5_metrics10", "c2_metrics"):
    B = 16 if name == "c2_metrics" else 654
    if name == "c5_metrics10":
        mflags = 0     # all groups: the 10 keys of the evaluation list
    pr, gt = synth.depth_pair((B, 1, 480, 640), 105, device=dev)
    ws = _lib.workspace(dev, B)
    o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
    f = lambda: _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(gt), B, 480 * 640, mflags, _lib.ptr(ws), _lib.ptr(o64), _lib.ptr(o32), None, None, sp()))
elif name in ("c2_fused", "c1_berhu"):
    shape = (16, 1, 480, 640) if name == "c2_fused" else (8, 1, 228, 304)
    kind = _lib.LOSS_SILOG if name == "c2_fused" else _lib.LOSS_BERHU
    pr, gt = synth.depth_pair(shape, 102, device=dev)
    ws = _lib.workspace(dev, shape[0]); grad = torch.empty(shape, device=dev)
    o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
    f = lambda: _lib.check(lib.mde_masked_loss_metrics(kind, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0, mflags,
                                                       _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grad), _lib.ptr(o64), _lib.ptr(o32), sp()))
elif name == "c1_eigen":
    shape = (8, 1, 228, 304)
    pr, gt = synth.depth_pair(shape, 102, device=dev)
    ws = _lib.workspace(dev, shape[0]); grad = torch.empty(shape, device=dev)
    f = lambda: _lib.check(lib.mde_masked_loss(_lib.LOSS_EIGEN, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0,
                                               _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grad), sp()))
elif name in ("dorn_fused", "dorn_decode", "ord_loss"):
    shape = (8, 136, 257, 353); N, C2, H, W = shape
    x, gt = synth.dorn_inputs(shape, 103, device=dev)
    ws = _lib.workspace(dev, 1)
    dec = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev); dep = torch.empty((N, 1, H, W), device=dev)
    if name == "dorn_fused":
        gx = torch.empty(shape, device=dev)
        f = lambda: _lib.check(lib.mde_dorn_fused(_lib.ptr(x), 0, _lib.ptr(gt), N, C2 // 2, H * W, 0.001, 1.0, 0, 1.0, _lib.ptr(ws), _lib.ptr(loss_t), None,
                                                  _lib.ptr(dec), _lib.ptr(dep), _lib.ptr(gx), sp()))
    elif name == "dorn_decode":
        f = lambda: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, C2 // 2, H * W, None, _lib.ptr(dec), sp()))
    else:
        prob = torch.rand((N, C2 // 2, H, W), device=dev); y = dorn.depth_to_label(gt, 0.001, 1.0, C2 // 2); gp = torch.empty_like(prob)
        f = lambda: _lib.check(lib.mde_ord_loss(_lib.ptr(prob), _lib.ptr(y), N, C2 // 2, H * W, 1.0, _lib.ptr(ws), _lib.ptr(loss_t), _lib.ptr(gp), sp()))
elif name == "vnl":
    gt, pred, trip = synth.vnl_inputs((8, 1, 385, 385), 104, device=dev)
    ws = _lib.workspace(dev, 8); scratch = torch.empty(int(lib.mde_vnl_scratch_bytes(8, 100000, 385, 385)), dtype=torch.uint8, device=dev)
    grad = torch.empty_like(pred)
    f = lambda: _lib.check(lib.mde_vnl_loss(_lib.ptr(gt), _lib.ptr(pred), 0, _lib.ptr(trip), 8, 385, 385, 100000, 519.0, 519.0, 1, 1.0, _lib.ptr(ws), _lib.ptr(scratch),
                                            _lib.ptr(loss_t), None, _lib.ptr(grad), sp()))
elif name == "wcel":
    from mono_depth_estimation_b200 import wcel
    B, Cc, H, W = 2, 150, 97, 131
    pw = wcel.vnl_params(0.01, 1.1, Cc)
    w = torch.tensor(pw["wce_loss_weight"], dtype=torch.float64)
    w32 = (w / w.sum(1, keepdim=True)).float().to(dev).contiguous(); rowsum = w32.double().sum(1).float().contiguous()
    gt = torch.rand((B, 1, H, W), device=dev) * 1.2 + 0.005; gt[1, :, :9, :] = -1.0
    bins = wcel.depth_to_bins(gt, 0.01, 1.1, Cc)
    logits = torch.randn((B, Cc, H, W), device=dev) * 3.0; gl = torch.empty_like(logits)
    ws = _lib.workspace(dev, B)
    f = lambda: _lib.check(lib.mde_wcel_loss(_lib.ptr(logits), 0, _lib.ptr(bins), _lib.ptr(gt), _lib.ptr(w32), _lib.ptr(rowsum), B, Cc, H * W,
                                             1.0, _lib.ptr(ws), _lib.ptr(loss_t), _lib.ptr(gl), sp()))
elif name in ("midas_mse", "robust"):
    Bm, Hm, Wm = 64, 384, 384
    tgt = torch.rand((Bm, Hm, Wm), device=dev) * 9.5 + 0.5
    tgt[torch.rand((Bm, Hm, Wm), device=dev) < 0.2] = 0.0
    prd = 0.7 / tgt.clamp_min(0.3) + 0.2 + torch.rand((Bm, Hm, Wm), device=dev) * 1e-3
    ws = _lib.workspace(dev, Bm); lg = torch.empty_like(prd)
    if name == "midas_mse":
        f = lambda: _lib.check(lib.mde_midas_loss(_lib.ptr(prd), 0, _lib.ptr(tgt), None, None, Bm, Hm, Wm, 0, 0.5, 4, 1.0, _lib.ptr(ws), _lib.ptr(loss_t), _lib.ptr(lg), sp()))
    else:
        stp = torch.empty((Bm, 8), device=dev); stt = torch.empty((Bm, 8), device=dev); tn = torch.empty_like(tgt)
        scr = torch.empty(int(lib.mde_robust_scratch_bytes(Bm)) // 8, dtype=torch.float64, device=dev)
        f = lambda: _lib.check(lib.mde_robust_normalize(_lib.ptr(prd), _lib.ptr(tgt), Bm, Hm * Wm, _lib.ptr(scr), _lib.ptr(stp), _lib.ptr(stt), _lib.ptr(lg), _lib.ptr(tn), sp()))
elif name == "stdepth":
    from mono_depth_estimation_b200 import stdepth
    from mono_depth_estimation_b200.synth import stdepth_inputs
    B, Cc, H, W = 8, 10, 512, 512
    pred, targ, rgba = (v.to(dev) for v in stdepth_inputs(900, B, Cc, H, W))
    ws = _lib.workspace(dev, B); out8 = torch.empty(8, device=dev); gr = torch.empty_like(pred)
    f = lambda: _lib.check(lib.mde_stdepth_loss(_lib.ptr(pred), 0, _lib.ptr(targ), _lib.ptr(rgba), 4, B, Cc, H * W, 3, 1.0, 1.0, 0.85, 1.0, _lib.ptr(ws), _lib.ptr(out8), _lib.ptr(gr), sp()))
elif name == "pointcloud":
    d = torch.rand((64, 480, 640), device=dev) * 12; out = torch.empty((64, 480, 640, 3), device=dev)
    f = lambda: _lib.check(lib.mde_point_cloud(_lib.ptr(d), 64, 480, 640, 0.8575, 0.1, 100.0, None, 0, _lib.ptr(out), sp()))
else:
    raise SystemExit("unknown name")
for _ in range(reps):
    f()
torch.cuda.synchronize()
print("ok", name, float(loss_t) if name not in ("c5_metrics", "c5_metrics10", "c2_metrics", "dorn_decode", "pointcloud", "robust", "stdepth") else "")
