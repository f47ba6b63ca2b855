#!/usr/bin/env python
"""Parity + timing of the C2 kernels (fused SILog + metrics, plain SILog) through the C ABI.

    python tools/check_c2.py [--no-parity] [--batch 16]

Parity: loss, full gradient, every metric value and the exact counts against the CPU oracle on the same
seeded batch. Timing: one CUDA graph of `ring` launches over distinct input AND gradient buffers
(ring > L2), CUDA events, median of 20 replays."""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--no-parity", action="store_true")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--ring", type=int, default=8)
args = ap.parse_args()
lib = _lib.load(); dev = torch.device("cuda", 0)
shape = (args.batch, 1, 480, 640); npx = shape[0] * 480 * 640
NAMES = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
mflags = 0
for n in NAMES:
    mflags |= _lib.METRIC_GROUP.get(n, 0)
ring = [synth.depth_pair(shape, 700 + i, device=dev) for i in range(args.ring)]
grads = [torch.empty(shape, device=dev) for _ in range(args.ring)]
ws = _lib.workspace(dev, shape[0])
loss_t = torch.empty((), device=dev)
o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
lp = _lib.LossParams(0.85, 1e-9, 1, 1)
sp = lambda: _lib.stream_ptr(dev)


def fused(i, flags=mflags):
    pr, g = ring[i % args.ring]
    _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], 480, 640, C.byref(lp), 1.0, flags,
                                           _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i % args.ring]), _lib.ptr(o64), _lib.ptr(o32), sp()))


def plain(i):
    pr, g = ring[i % args.ring]
    _lib.check(lib.mde_masked_loss(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], 480, 640, C.byref(lp), 1.0,
                                   _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i % args.ring]), sp()))


def metrics_only(i):
    pr, g = ring[i % args.ring]
    _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(g), shape[0], 480 * 640, mflags, _lib.ptr(ws), _lib.ptr(o64), _lib.ptr(o32), None, None, sp()))


out = {}
if not args.no_parity:
    from oracle import losses as olosses, metrics as ometrics
    pr, g = ring[0]
    p64, g64 = pr.double().cpu(), g.double().cpu()
    l64, gr64 = olosses.loss_and_grad(olosses.silog, p64, g64, 0.85)
    all_names = ["delta1", "delta2", "delta3", "mae", "mse", "log10", "msle", "absrel", "sqrel", "rmse"]
    v64 = {n: float(v) for n, v in zip(all_names, ometrics.compute(p64, g64, all_names))}
    raw = ometrics.raw_sums(pr.cpu(), g.cpu()) if hasattr(ometrics, "raw_sums") else None
    for label, fn, fl in (("fused7", fused, mflags), ("fused_all", fused, 0), ("plain", plain, None)):
        grads[0].fill_(float("nan"))
        if fl is None:
            fn(0)
        else:
            fn(0, fl)
        torch.cuda.synchronize()
        gerr = float((grads[0].double().cpu() - gr64).abs().max() / gr64.abs().max())
        res = {"loss_rel": abs(float(loss_t) - float(l64)) / abs(float(l64)), "grad_rel_max": gerr}
        if fl is not None:
            vals = o64[:_lib.METRIC_NM].cpu()
            chk = NAMES if fl else all_names
            res["metric_rel_max"] = max(abs(float(vals[_lib.METRIC_INDEX[n]]) - v64[n]) / abs(v64[n]) for n in chk)
            rawg = o64[2 * _lib.METRIC_NM:2 * _lib.METRIC_NM + 4].cpu()
            valid = g.cpu() > 0
            pc = pr.cpu().clamp_min(1e-7)[valid]; tc = g.cpu()[valid]
            ratio = torch.maximum(pc / tc, tc / pc)
            exact = [int(valid.sum())] + [int((ratio < 1.25 ** k).sum()) for k in (1, 2, 3)]
            res["counts_gpu"] = [int(x) for x in rawg]
            res["counts_exact"] = exact
            res["counts_match"] = res["counts_gpu"] == exact
        out[label] = res


def timeit(fn, label, bytes_px):
    for i in range(3 * args.ring):
        fn(i)
    torch.cuda.synchronize()
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=st):
            for i in range(args.ring):
                fn(i)
        gph.replay(); st.synchronize()
        ts = []
        for r in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(4):
                gph.replay()
            b.record(st); st.synchronize()
            ts.append(1e3 * a.elapsed_time(b) / (4 * args.ring))
    ts.sort()
    us = ts[len(ts) // 2]
    out[label + "_us"] = us
    out[label + "_frac_of_6454.6"] = bytes_px * npx / (us * 1e-6) / 1e9 / 6454.6


# workspaces are per stream: take the timing stream's workspace inside timeit via the default stream one (same device)
timeit(fused, "fused7", 12.0)
timeit(lambda i: fused(i, 0), "fused_all", 12.0)
timeit(plain, "plain", 12.0)
timeit(metrics_only, "metrics7", 8.0)
print(json.dumps(out))
