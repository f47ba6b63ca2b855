#!/usr/bin/env python
"""Kernel-level numbers for the 'next' rows of the scope table: the MiDaS family, TrimmedProcrustesLoss
and the layered-depth base criterion. Same method as tools/bench_all.py (which calls run()); standalone:
    python tools/bench_next.py [--quick]
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402


def run(report, timed, reps):
    from mono_depth_estimation_b200 import _lib, criteria, stdepth
    from mono_depth_estimation_b200.synth import stdepth_inputs
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731
    loss_t = torch.empty((), device=dev)

    # ---------------- MiDaS alignment + MidasLoss: 64 x 384 x 384 (75 MB pred+target: L2-resident when replayed) ------
    Bm, Hm, Wm = 64, 384, 384
    tgt = torch.rand((Bm, Hm, Wm), device=dev) * 9.5 + 0.5
    tgt[torch.rand((Bm, Hm, Wm), device=dev) < 0.2] = 0.0
    prd = 0.7 / tgt.clamp_min(0.3) + 0.2 + torch.rand((Bm, Hm, Wm), device=dev) * 1e-3
    sc = torch.empty(Bm, device=dev); sh = torch.empty(Bm, device=dev); al = torch.empty_like(prd)
    ws = _lib.workspace(dev, Bm)
    pxm = Bm * Hm * Wm
    cfg = "MiDaS 64x384x384"
    fns = [lambda: _lib.check(lib.mde_scale_and_shift(_lib.ptr(prd), 0, _lib.ptr(tgt), None, Bm, Hm * Wm, _lib.ptr(ws), _lib.ptr(sc), _lib.ptr(sh), None, sp()))]
    us, g = timed(fns, reps)
    report(cfg, "compute_scale_and_shift (5 masked sums + 2x2 solve per image)", pxm, us, 8.0, g)
    fns = [lambda: _lib.check(lib.mde_apply_scale_shift(_lib.ptr(prd), 0, _lib.ptr(sc), _lib.ptr(sh), Bm, Hm * Wm, _lib.ptr(al), sp()))]
    us, g = timed(fns, reps)
    report(cfg, "apply scale/shift", pxm, us, 8.0, g)
    lg = torch.empty_like(prd)
    fns = [lambda: _lib.check(lib.mde_midas_loss(_lib.ptr(prd), 0, _lib.ptr(tgt), None, None, Bm, Hm, Wm, 0, 0.5, 4, 1.0, _lib.ptr(ws), _lib.ptr(loss_t), _lib.ptr(lg), sp()))]
    us, g = timed(fns, reps)
    report(cfg, "MidasLoss('mse', alpha 0.5, 4 scales) fwd+bwd, one launch", pxm, us, 12.0, g)
    for name, mod in (("MidasLoss('ssimse') fwd+bwd via module (4 launches)", criteria.MidasLoss(alpha=0.5, loss="ssimse")),
                      ("TrimmedProcrustesLoss fwd+bwd via module (5 launches)", criteria.TrimmedProcrustesLoss(alpha=0.5))):
        def f(mod=mod):
            mod(prd.detach().requires_grad_(True), tgt).backward()
        us, g = timed([f], reps)
        report(cfg, name, pxm, us, 12.0, g)
    stp = torch.empty((Bm, 8), device=dev); stt = torch.empty((Bm, 8), device=dev); tn = torch.empty_like(tgt)
    scr = torch.empty(int(lib.mde_robust_scratch_bytes(Bm)) // 8, dtype=torch.float64, device=dev)
    fns = [lambda: _lib.check(lib.mde_robust_normalize(_lib.ptr(prd), _lib.ptr(tgt), Bm, Hm * Wm, _lib.ptr(scr), _lib.ptr(stp), _lib.ptr(stt), _lib.ptr(al), _lib.ptr(tn), sp()))]
    us, g = timed(fns, reps)
    report(cfg, "normalize_prediction_robust of pred and target (exact medians + normalisation, 2 launches)", pxm, us, 16.0, g)
    # the same at the training batch of the `midas` method (8 images): every pair is cut into units that fill the GPU
    B8 = 8
    fns = [lambda: _lib.check(lib.mde_robust_normalize(_lib.ptr(prd), _lib.ptr(tgt), B8, Hm * Wm, _lib.ptr(scr), _lib.ptr(stp), _lib.ptr(stt), _lib.ptr(al), _lib.ptr(tn), sp()))]
    us, g = timed(fns, reps)
    report("MiDaS 8x384x384", "normalize_prediction_robust of pred and target (2 launches)", B8 * Hm * Wm, us, 16.0, g)
    mod = criteria.TrimmedProcrustesLoss(alpha=0.5)
    p8, t8 = prd[:B8].contiguous(), tgt[:B8].contiguous()
    def f8():
        mod(p8.detach().requires_grad_(True), t8).backward()
    us, g = timed([f8], reps)
    report("MiDaS 8x384x384", "TrimmedProcrustesLoss fwd+bwd via module (5 launches)", B8 * Hm * Wm, us, 12.0, g)
    del tgt, prd, al, lg, tn

    # ---------------- layered-depth base criterion (bts default 'silma'): 8 x 10 x 512 x 512 ---------------------------
    for (B, C, H, W), loss_name in (((8, 10, 512, 512), "silma"), ((8, 10, 512, 512), "mae+mse+fbdivergence"), ((4, 20, 512, 512), "silma")):
        ring = []
        for i in range(3):   # 3 x (pred + targ + grad) x 84 MB: beyond L2
            pred, targ, rgba = stdepth_inputs(900 + i, B, C, H, W)
            ring.append((pred.to(dev), targ.to(dev), rgba.to(dev), torch.empty((B, C, H, W), device=dev)))
        flags = stdepth._flags_of(loss_name)
        ws = _lib.workspace(dev, B)
        out8 = torch.empty(8, device=dev)
        fns = [lambda p=p, t=t, x=x, gr=gr: _lib.check(lib.mde_stdepth_loss(_lib.ptr(p), 0, _lib.ptr(t), _lib.ptr(x), 4, B, C, H * W, flags, 1.0, 1.0, 0.85,
                                                                           1.0, _lib.ptr(ws), _lib.ptr(out8), _lib.ptr(gr), sp()))
               for p, t, x, gr in ring]
        us, g = timed(fns, reps)
        report("stdepth %dx%dx%dx%d" % (B, C, H, W), "base criterion '%s' fwd+bwd, one launch" % loss_name, B * H * W, us, 12.0 * C + 4.0, g)
        del ring


def run_ord(report, timed, reps):
    """ordLoss(P, y) alone (reference criteria.py:744-787 as modules/dorn.py:160-163 calls it) at C3."""
    from mono_depth_estimation_b200 import _lib, dorn, synth
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    N, C2, H, W = 8, 136, 257, 353
    _, gt = synth.dorn_inputs((N, C2, H, W), 103, device=dev)
    prob = torch.rand((N, C2 // 2, H, W), device=dev)
    y = dorn.depth_to_label(gt, 0.001, 1.0, C2 // 2)
    gp = torch.empty_like(prob)
    ws = _lib.workspace(dev, 1)
    loss_t = torch.empty((), device=dev)
    fns = [lambda: _lib.check(lib.mde_ord_loss(_lib.ptr(prob), _lib.ptr(y), N, C2 // 2, H * W, 1.0, _lib.ptr(ws), _lib.ptr(loss_t),
                                               _lib.ptr(gp), _lib.stream_ptr(dev)))]
    us, g = timed(fns, reps)
    report("C3", "ordLoss(P, y) fwd+bwd", N * H * W, us, 4.0 * (C2 // 2) * 2 + 4.0, g)


if __name__ == "__main__":
    import bench_all
    bench_all.dev  # noqa: B018
    if "--ord" in sys.argv:
        run_ord(bench_all.report, bench_all.timed, 30)
    else:
        run(bench_all.report, bench_all.timed, 5 if "--quick" in sys.argv else 30)
