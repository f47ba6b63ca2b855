"""Small-input losses (C1 size): parity against the CPU oracle and time per launch for every loss kind.

    python tools/c1_probe.py            # register-resident kernel (resident_loss.cu) where it applies
    MDE_NO_RESIDENT=1 python tools/c1_probe.py   # the generic persistent kernel for the same calls

Writes one JSON line per (kind, shape) to stdout. The oracle is used as the checker only."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mono_depth_estimation_b200 import _lib, criteria, metrics, synth  # noqa: E402
from oracle import losses as olosses, metrics as ometrics  # noqa: E402
import bench  # noqa: E402

NAMES = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
KINDS = [("l1", _lib.LOSS_L1, olosses.masked_l1), ("mse", _lib.LOSS_MSE, olosses.masked_mse), ("berhu", _lib.LOSS_BERHU, olosses.berhu),
         ("laina", _lib.LOSS_LAINA_BERHU, olosses.laina_berhu), ("silog", _lib.LOSS_SILOG, olosses.silog)]


def parity(shape, seed, with_metrics, noise=0.5):
    dev = torch.device("cuda:0")
    pred, gt = synth.depth_pair(shape, seed, border=2, noise=noise)
    rows = []
    for name, kind, ofn in KINDS:
        p = pred.to(dev).requires_grad_(True)
        mc = metrics.MetricComputation(NAMES) if with_metrics else None
        g = gt.to(dev)                                # kept alive: the metric hand-over holds weak references
        n0 = _lib.launch_count()
        loss = criteria.masked_loss(kind, p, g, metrics=mc)
        loss.backward()
        l64, g64 = olosses.loss_and_grad(ofn, pred.double(), gt.double())
        rel = abs(float(loss) - float(l64)) / max(abs(float(l64)), 1e-30)
        gerr = float((p.grad.double().cpu() - g64).abs().max() / g64.abs().max().clamp_min(1e-30))
        merr = 0.0
        cnt_ok = True
        if with_metrics:
            vals = mc.compute(p.detach(), g)
            assert _lib.launch_count() - n0 == 1, "metrics must come from the loss launch"
            v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), NAMES)]
            merr = max(abs(float(a) - b) / max(abs(b), 1e-30) for a, b in zip(vals, v64))
            # the three threshold counts are integers: exact against the fp32 oracle
            v32 = [float(v) for v in ometrics.compute(pred, gt, NAMES[:3])]
            n = int((gt > 0).sum())
            cnt_ok = all(round(float(a) * n) == round(b * n) for a, b in zip(vals[:3], v32))
        rows.append({"kind": name, "shape": list(shape), "metrics": with_metrics, "loss_rel": rel, "grad_rel": gerr, "metric_rel": merr,
                     "counts_exact": cnt_ok})
    return rows


def timing(shape, with_metrics):
    dev = torch.device("cuda:0")
    lib = _lib.load()
    px = shape[0] * shape[2] * shape[3]
    nring = max(4, int(300e6 // (px * 12)))
    ring = [synth.depth_pair(shape, 101 + i, device=dev) for i in range(nring)]
    grads = [torch.empty(shape, device=dev) for _ in range(nring)]
    ws = _lib.workspace(dev, shape[0])
    o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
    o32 = torch.empty(24, device=dev)
    loss_t = torch.empty((), device=dev)
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    mflags = 0
    for n in NAMES:
        mflags |= _lib.METRIC_GROUP.get(n, 0)
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731
    out = {}
    for name, kind, _ in KINDS:
        if with_metrics:
            fns = [lambda pr=pr, gt=gt, gr=gr, kind=kind: _lib.check(lib.mde_masked_loss_metrics(
                kind, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0, mflags, _lib.ptr(ws),
                _lib.ptr(loss_t), None, _lib.ptr(gr), _lib.ptr(o64), _lib.ptr(o32), sp())) for (pr, gt), gr in zip(ring, grads)]
        else:
            fns = [lambda pr=pr, gt=gt, gr=gr, kind=kind: _lib.check(lib.mde_masked_loss(
                kind, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0, _lib.ptr(ws),
                _lib.ptr(loss_t), None, _lib.ptr(gr), sp())) for (pr, gt), gr in zip(ring, grads)]
        us, graphed = bench.graph_timed(fns, dev, 20)
        out[name] = round(us, 2)
    return {"timing_us": out, "shape": list(shape), "metrics": with_metrics, "ring": nring,
            "resident": os.environ.get("MDE_NO_RESIDENT", "0") in ("", "0")}


if __name__ == "__main__":
    torch.cuda.set_device(0)
    bad = 0
    for shape, seed, noise in (((8, 1, 228, 304), 3, 0.5), ((8, 1, 228, 304), 13, 3.0), ((2, 1, 47, 63), 4, 3.0), ((8, 1, 300, 400), 5, 3.0),
                               ((1, 1, 9, 7), 6, 0.5)):
        for wm in (False, True):
            for r in parity(shape, seed, wm, noise):
                ok = r["loss_rel"] < 1e-5 and r["grad_rel"] < 1e-5 and r["metric_rel"] < 1e-5 and r["counts_exact"]
                r["ok"] = ok
                bad += 0 if ok else 1
                print(json.dumps(r))
    for shape in ((8, 1, 228, 304), (8, 1, 300, 400)):
        for wm in (True, False):
            print(json.dumps(timing(shape, wm)))
    print(json.dumps({"failures": bad}))
    sys.exit(1 if bad else 0)
