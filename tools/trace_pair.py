#!/usr/bin/env python
"""Phase trace of TWO consecutive launches of the C2 SILog kernel (instrumented twin tools/libmde_dbg.so, see SS_TP in
csrc/silog_ss.cuh): per-phase times, the latest exit over all warps, the launch-to-launch period and the gap between one
grid's last exit and the next grid's first CTA.

    MDE_B200_LIB=$PWD/tools/libmde_dbg.so python tools/trace_pair.py [--plain] [--graph]
"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--plain", action="store_true")
ap.add_argument("--graph", action="store_true", help="replay one CUDA graph of 8 launches instead of 300 stream launches")
ap.add_argument("--tag", default="")
args = ap.parse_args()
lib = _lib.load(); dev = torch.device("cuda", 0)
shape = (16, 1, 480, 640)
ring = [synth.depth_pair(shape, 500 + i, device=dev) for i in range(8)]
grads = [torch.empty(shape, device=dev) for _ in range(8)]
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    ws = _lib.workspace(dev, 16)
loss_t = torch.empty((), device=dev)
lp = _lib.LossParams(0.85, 1e-9, 1, 1)
trace = torch.zeros(2 * 296 * 8, dtype=torch.int64, device=dev)
o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
mflags = _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_RSQ


def run(i):
    pr, g = ring[i % 8]
    if not args.plain:
        _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, 16, 480, 640, C.byref(lp), 1.0, mflags,
                   _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i % 8]), _lib.ptr(o64), _lib.ptr(o32), _lib.stream_ptr(dev)))
    else:
        _lib.check(lib.mde_masked_loss(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, 16, 480, 640, C.byref(lp), 1.0,
                   _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i % 8]), _lib.stream_ptr(dev)))


NAMES = ["start", "a_done", "published", "bar_exit", "b_start", "exit_all", "gathered"]
out = []
with torch.cuda.stream(st):
    for i in range(16): run(i)
    st.synchronize()
    _lib.check(lib.mde_debug_set_trace(_lib.ptr(trace)))
    gph = None
    if args.graph:
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=st):
            for i in range(8): run(i)
    for rep in range(6):
        if gph is not None:
            for _ in range(40): gph.replay()
        else:
            for i in range(304): run(i)
        st.synchronize()
        t = trace.view(2, 296, 8).cpu()
        halves = []
        for h in range(2):
            rows = t[h][t[h][:, 0] > 0]
            halves.append(rows)
        halves.sort(key=lambda r: int(r[:, 0].min()))          # older launch first
        t0 = int(halves[0][:, 0].min())
        rec = {}
        for lab, rows in zip(("prev", "last"), halves):
            rec[lab] = {n: [round((int(v) - t0) / 1e3, 3) for v in rows[:, k]] for k, n in enumerate(NAMES)}
            rec[lab]["smid"] = rows[:, 7].tolist()
        out.append(rec)
    _lib.check(lib.mde_debug_set_trace(None))
name = "trace_pair_%s%s%s.json" % ("plain" if args.plain else "fused", "_graph" if args.graph else "", args.tag)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"))


def stat(x):
    x = sorted(x)
    return "min %.2f med %.2f max %.2f" % (x[0], x[len(x) // 2], x[-1])


for rec in out[-3:]:
    p, l = rec["prev"], rec["last"]
    period = min(l["start"]) - min(p["start"])
    gap = min(l["start"]) - max(p["exit_all"])
    print("%s period %.2f us | prev: start spread %.2f, loop done [%s], published max %.2f, gathered [%s], coefficients max %.2f, "
          "exit(all warps) [%s] | gap last exit -> next first CTA %.2f us, next grid's CTAs start within %.2f us"
          % (name, period, max(p["start"]) - min(p["start"]), stat(p["a_done"]), max(p["published"]), stat(p["gathered"]),
             max(p["bar_exit"]), stat(p["exit_all"]), gap, max(l["start"]) - min(l["start"])))
