#!/usr/bin/env python
"""Does the launch time of the C2 fused kernel depend on how long the GPU has been busy (clock ramp)?
Replays a graph of 8 launches continuously for ~1.5 s and prints us/launch + nvidia-smi SM clock at intervals."""
import ctypes as C, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from mono_depth_estimation_b200 import _lib, synth
lib = _lib.load(); dev = torch.device("cuda", 0)
shape = (16, 1, 480, 640)
ring = [synth.depth_pair(shape, 700 + i, device=dev) for i in range(8)]
grads = [torch.empty(shape, device=dev) for _ in range(8)]
ws = _lib.workspace(dev, 16); loss_t = torch.empty((), device=dev)
o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
lp = _lib.LossParams(0.85, 1e-9, 1, 1); mflags = _lib.METRICS_NEED_LOG | _lib.METRICS_NEED_RSQ
def fused(i):
    pr, g = ring[i % 8]
    _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, 16, 480, 640, C.byref(lp), 1.0, mflags,
               _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i % 8]), _lib.ptr(o64), _lib.ptr(o32), _lib.stream_ptr(dev)))
def smi():
    try:
        return subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as e:
        return str(e)
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    for i in range(24): fused(i)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(8): fused(i)
    time.sleep(1.0)   # let the GPU go idle first
    print("idle:", smi())
    t_start = time.perf_counter()
    rows = []
    while time.perf_counter() - t_start < 1.5:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(25): g.replay()
        b.record(st); st.synchronize()
        rows.append((time.perf_counter() - t_start, 1e3 * a.elapsed_time(b) / 200))
    for k in (0, 1, 2, 5, 10, 20, 50, 100, 200, len(rows) - 1):
        if k < len(rows): print("t=%.3fs  %.2f us/launch" % rows[k])
    print("busy:", smi())
