#!/usr/bin/env python
"""Kernel-level numbers for every BASELINE.json config on one B200 (fills the tables of DESIGN.md).

For each config the hot-path call is replayed from a CUDA graph over a ring of distinct inputs larger
than L2 (or re-run eagerly where a ring does not fit), timed with CUDA events on the launching stream.
Prints one JSON object per line: config, kernel, px, us per call, Mpix/s, algorithmic bytes, achieved
GB/s and the fraction of the measured HBM peak.  python tools/bench_all.py [--quick]
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from mono_depth_estimation_b200 import _lib, criteria, dorn, metrics, pointcloud, synth  # noqa: E402

PEAK = 6454.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])
dev = torch.device("cuda", 0)
QUICK = "--quick" in sys.argv


def timed(fns, reps, use_graph=True):
    """fns: list of zero-arg callables (one per ring slot). Returns us per call."""
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        for f in fns:
            f()
        side.synchronize()
        g = None
        if use_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for f in fns:
                        f()
            except Exception as e:
                sys.stderr.write("graph capture failed: %s\n" % str(e).splitlines()[0])
                g = None
                torch.cuda.synchronize()
        for _ in range(3):
            if g is not None:
                g.replay()
            else:
                for f in fns:
                    f()
        side.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        for _ in range(reps):
            if g is not None:
                g.replay()
            else:
                for f in fns:
                    f()
        b.record(side)
        side.synchronize()
    return 1e3 * a.elapsed_time(b) / (reps * len(fns)), g is not None


ONLY = os.environ.get("BENCH_ONLY", "")


def report(config, kernel, px, us, bytes_per_px, graph, extra=None):
    gbs = bytes_per_px * px / (us * 1e-6) / 1e9
    d = {"config": config, "kernel": kernel, "px": px, "us": round(us, 2), "mpix_s": round(px / us, 1),
         "alg_bytes_per_px": bytes_per_px, "achieved_gbs": round(gbs, 1), "frac_measured_peak": round(gbs / PEAK, 4),
         "frac_nominal_8000": round(gbs / 8000.0, 4), "graph": graph}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def ring_of(make, nbytes_each, min_total=400e6, max_slots=16):
    n = int(min(max_slots, max(2, -(-min_total // nbytes_each))))
    return [make(i) for i in range(n)]


def main():
    lib = _lib.load()
    reps = 5 if QUICK else 30
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731

    # ---------------- C1: berHu fwd+bwd (+ fused metrics), L1, MSE, Laina, Eigen on 8x1x228x304 ------------
    shape = synth.SHAPES["C1"]
    px = shape[0] * shape[2] * shape[3]
    ring = ring_of(lambda i: synth.depth_pair(shape, 101 + i, device=dev), 8 * px, min_total=300e6, max_slots=48)
    names = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
    for kname, kind in (("berhu", _lib.LOSS_BERHU), ("l1", _lib.LOSS_L1), ("mse", _lib.LOSS_MSE),
                        ("laina_berhu", _lib.LOSS_LAINA_BERHU), ("silog", _lib.LOSS_SILOG), ("eigen", _lib.LOSS_EIGEN)):
        fns = []
        for pr, gt in ring:
            def f(pr=pr, gt=gt, kind=kind):
                criteria.masked_loss(kind, pr.detach().requires_grad_(True), gt).backward()
            fns.append(f)
        us, g = timed(fns, reps)
        report("C1", kname + "_fwd_bwd(module+autograd)", px, us, 12.0, g)
    mc = metrics.MetricComputation(names, strict=False)
    crit = criteria.berHuLoss().fuse_metrics(mc)
    fns = []
    for pr, gt in ring:
        def f(pr=pr, gt=gt):
            p_ = pr.detach().requires_grad_(True)
            crit(p_, gt).backward()
            mc.compute(p_.detach(), gt)
        fns.append(f)
    us, g = timed(fns, reps)
    report("C1", "berhu_fwd_bwd+metrics fused (config C1 step)", px, us, 12.0, g)
    fns = [lambda pr=pr, gt=gt: metrics.fused_metrics(pr, gt, names=names) for pr, gt in ring]
    us, g = timed(fns, reps)
    report("C1", "metrics", px, us, 8.0, g)

    # ---------------- C2 at several batch sizes: where the fixed costs stop mattering ------------------------
    for B in ((16, 64, 128) if not QUICK else (16,)):
        shp = (B, 1, 480, 640)
        px = B * 480 * 640
        ring = ring_of(lambda i: synth.depth_pair(shp, 300 + i, device=dev), 8 * px, min_total=400e6, max_slots=8)
        mc = metrics.MetricComputation(names, strict=False)
        crit = criteria.silog_loss(0.85).fuse_metrics(mc)
        ws = _lib.workspace(dev, B)
        loss_t = torch.empty((), device=dev)
        grad_t = torch.empty(shp, device=dev)
        o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
        o32 = torch.empty(24, device=dev)
        lp = _lib.LossParams(0.85, 1e-9, 1, 1)
        mflags = _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_REL
        fns = [lambda pr=pr, gt=gt: _lib.check(lib.mde_masked_loss_metrics(
            _lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(gt), None, B, 480, 640, C.byref(lp), 1.0, mflags, _lib.ptr(ws),
            _lib.ptr(loss_t), None, _lib.ptr(grad_t), _lib.ptr(o64), _lib.ptr(o32), sp())) for pr, gt in ring]
        us, g = timed(fns, reps)
        report("C2 B=%d" % B, "silog fwd+bwd + metrics, fused kernel", px, us, 12.0, g)
        fns = [lambda pr=pr, gt=gt: _lib.check(lib.mde_masked_loss(
            _lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(gt), None, B, 480, 640, C.byref(lp), 1.0, _lib.ptr(ws),
            _lib.ptr(loss_t), None, _lib.ptr(grad_t), sp())) for pr, gt in ring]
        us, g = timed(fns, reps)
        report("C2 B=%d" % B, "silog fwd+bwd kernel", px, us, 12.0, g)
        fns = [lambda pr=pr, gt=gt: _lib.check(lib.mde_metrics(
            _lib.ptr(pr), 0, _lib.ptr(gt), B, 480 * 640, mflags, _lib.ptr(ws), _lib.ptr(o64), _lib.ptr(o32), None, None,
            sp())) for pr, gt in ring]
        us, g = timed(fns, reps)
        report("C2 B=%d" % B, "metrics kernel (7 default metrics)", px, us, 8.0, g)
        del ring, grad_t

    # ---------------- C3: DORN fused logits -> decode, depth, loss, grad (K = 68) ------------------------------
    shape = synth.SHAPES["C3"] if not QUICK else (2, 136, 257, 353)
    N, C2, H, W = shape
    px = N * H * W
    ring = [synth.dorn_inputs(shape, 103 + i, device=dev) for i in range(2)]   # 395 MB of logits each
    ws = _lib.workspace(dev, 1)
    loss_t = torch.empty((), device=dev)
    dec = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev)
    dep = torch.empty((N, 1, H, W), device=dev)
    gx = torch.empty(shape, device=dev)
    prob = torch.empty((N, C2 // 2, H, W), device=dev)
    fns = [lambda x=x, gt=gt: _lib.check(lib.mde_dorn_fused(_lib.ptr(x), 0, _lib.ptr(gt), N, C2 // 2, H * W, 0.001, 1.0, 0, 1.0,
                                                          _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(dec), _lib.ptr(dep),
                                                          _lib.ptr(gx), sp())) for x, gt in ring]
    us, g = timed(fns, reps)
    report("C3", "dorn fused: decode+depth+loss+grad", px, us, 8.0 * C2 + 16.0, g, {"logit_gbs": round(8.0 * C2 * px / us / 1e3, 1)})
    fns = [lambda x=x, gt=gt: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, C2 // 2, H * W, _lib.ptr(prob), _lib.ptr(dec), sp()))
           for x, gt in ring]
    us, g = timed(fns, reps)
    report("C3", "ordinal layer fwd (decode + P)", px, us, 4.0 * C2 + 2.0 * C2 + 8.0, g)
    fns = [lambda x=x, gt=gt: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, C2 // 2, H * W, None, _lib.ptr(dec), sp()))
           for x, gt in ring]
    us, g = timed(fns, reps)
    report("C3", "decode only (inference)", px, us, 4.0 * C2 + 8.0, g)
    y = dorn.depth_to_label(ring[0][1], 0.001, 1.0, C2 // 2)
    gp = torch.empty_like(prob)
    fns = [lambda: _lib.check(lib.mde_ord_loss(_lib.ptr(prob), _lib.ptr(y), N, C2 // 2, H * W, 1.0, _lib.ptr(ws), _lib.ptr(loss_t),
                                               _lib.ptr(gp), sp()))]
    us, g = timed(fns, reps)
    report("C3", "ordLoss(P, y) fwd+bwd", px, us, 4.0 * (C2 // 2) * 2 + 4.0, g)
    del ring, gx, prob, gp

    # ---------------- C4: VNL, 100k triplets x 8 images at 385x385 ------------------------------------------------
    shape = synth.SHAPES["C4"]
    gt, pred, trip = synth.vnl_inputs(shape, 104, device=dev)
    B, _, H, W = shape
    px = B * H * W
    n_trip = trip.shape[1]
    ws = _lib.workspace(dev, B)
    scratch = torch.empty(int(lib.mde_vnl_scratch_bytes(B, n_trip, H, W)), dtype=torch.uint8, device=dev)
    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.empty_like(pred)
    fns = [lambda: _lib.check(lib.mde_vnl_loss(_lib.ptr(gt), _lib.ptr(pred), 0, _lib.ptr(trip), B, H, W, n_trip, 519.0, 519.0, 1, 1.0,
                                               _lib.ptr(ws), _lib.ptr(scratch), _lib.ptr(loss_t), _lib.ptr(stats), _lib.ptr(grad), sp()))]
    us, g = timed(fns, reps)
    report("C4", "vnl fwd+bwd (100k triplets x 8)", px, us, 12.0 + 24.0 * n_trip / px, g,
           {"mtriplets_s": round(B * n_trip / us, 1), "valid_triplets": float(stats[0].item()), "loss": float(loss_t.item())})

    # ---------------- C4 companion: WCEL fwd+bwd, bins -> depth, depth -> bins at 8x150x385x385 ---
    from mono_depth_estimation_b200 import wcel
    Cc = 150
    pw = wcel.vnl_params(0.01, 1.1, Cc)
    w = torch.tensor(pw["wce_loss_weight"], dtype=torch.float64)
    w32 = (w / w.sum(1, keepdim=True)).float().to(dev).contiguous()
    rowsum = w32.double().sum(1).float().contiguous()
    border = torch.tensor(pw["depth_bin_border"], dtype=torch.float32, device=dev)
    logits = torch.randn((B, Cc, H, W), device=dev) * 3.0
    gtw = gt.clone()
    bins = wcel.depth_to_bins(gtw, 0.01, 1.1, Cc)
    gl = torch.empty_like(logits)
    ws = _lib.workspace(dev, B)
    fns = [lambda: _lib.check(lib.mde_wcel_loss(_lib.ptr(logits), 0, _lib.ptr(bins), _lib.ptr(gtw), _lib.ptr(w32), _lib.ptr(rowsum), B, Cc, H * W,
                                                1.0, _lib.ptr(ws), _lib.ptr(loss_t), _lib.ptr(gl), sp()))]
    us, g = timed(fns, reps)
    report("C4 WCEL", "wcel fwd+bwd (150 bins), count + fused pass", px, us, 8.0 * Cc + 8.0, g, {"loss": float(loss_t.item())})
    fns = [lambda: _lib.check(lib.mde_wcel_loss(_lib.ptr(logits), 0, _lib.ptr(bins), _lib.ptr(gtw), _lib.ptr(w32), _lib.ptr(rowsum), B, Cc, H * W,
                                                1.0, _lib.ptr(ws), _lib.ptr(loss_t), None, sp()))]
    us, g = timed(fns, reps)
    report("C4 WCEL", "wcel forward only", px, us, 4.0 * Cc + 8.0, g)
    sm = torch.softmax(logits, 1)
    dep = torch.empty((B, 1, H, W), device=dev)
    fns = [lambda: _lib.check(lib.mde_bins_to_depth(_lib.ptr(sm), 0, _lib.ptr(border), B, Cc, H * W, _lib.ptr(dep), sp()))]
    us, g = timed(fns, reps)
    report("C4 WCEL", "bins_to_depth (150 bins)", px, us, 4.0 * Cc + 4.0, g)
    fns = [lambda: _lib.check(lib.mde_bins_to_depth_bwd(_lib.ptr(dep), _lib.ptr(dep), _lib.ptr(border), B, Cc, H * W, 0, _lib.ptr(gl), sp()))]
    us, g = timed(fns, reps)
    report("C4 WCEL", "bins_to_depth backward", px, us, 4.0 * Cc + 8.0, g)
    del logits, gl, sm

    # ---------------- the 'next' rows: MiDaS family, TrimmedProcrustes, layered-depth criterion -----------
    import bench_next
    bench_next.run(report, timed, reps)

    # ---------------- C5: NYU-test-shaped eval, 654 x 480 x 640, 10 metrics ------------------------------------
    B = 654 if not QUICK else 64
    shp = (B, 1, 480, 640)
    px = B * 480 * 640
    pr, gtt = synth.depth_pair(shp, 105, device=dev)
    ws = _lib.workspace(dev, B)
    o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
    o32 = torch.empty(24, device=dev)
    piv = torch.empty((B, 12), dtype=torch.float64, device=dev)
    pir = torch.empty((B, 12), dtype=torch.float64, device=dev)
    for label, flags in (("10 metrics (all groups)", 0), ("7 default metrics", _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_REL),
                         ("10 metrics, reference math", _lib.METRICS_REFERENCE_MATH)):
        fns = [lambda flags=flags: _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(gtt), B, 480 * 640, flags, _lib.ptr(ws), _lib.ptr(o64),
                                                              _lib.ptr(o32), _lib.ptr(piv), _lib.ptr(pir), sp()))]
        us, g = timed(fns, reps)
        report("C5 (%d imgs)" % B, "eval metrics, " + label, px, us, 8.0, g)
    del pr, gtt

    # ---------------- point cloud ------------------------------------------------------------------------------------
    d = torch.rand((64, 480, 640), device=dev) * 12
    cam = pointcloud.Camera(0.8575, 0.1, 100.0, [[1, 0, 0, 0.5], [0, 1, 0, 0.2], [0, 0, 1, 1.0], [0, 0, 0, 1]])
    px = d.numel()
    out32 = torch.empty(tuple(d.shape) + (3,), device=dev)
    out64 = torch.empty(tuple(d.shape) + (3,), dtype=torch.float64, device=dev)
    mat = (C.c_float * 16)(*[float(v) for row in cam.matrix_world for v in row])
    fns = [lambda: _lib.check(lib.mde_point_cloud(_lib.ptr(d), 64, 480, 640, 0.8575, 0.1, 100.0, None, 0, _lib.ptr(out32), sp()))]
    us, g = timed(fns, reps)
    report("pointcloud 64x480x640", "fp32 out", px, us, 16.0, g)
    fns = [lambda: _lib.check(lib.mde_point_cloud(_lib.ptr(d), 64, 480, 640, 0.8575, 0.1, 100.0, mat, 1, _lib.ptr(out64), sp()))]
    us, g = timed(fns, reps)
    report("pointcloud 64x480x640", "fp64 out + world transform", px, us, 28.0, g)


if __name__ == "__main__":
    main()
