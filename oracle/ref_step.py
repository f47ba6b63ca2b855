"""CPU baseline process of bench.py: times the REFERENCE's own code for one C2 step on the host cores.

    CUDA_VISIBLE_DEVICES="" python -m oracle.ref_step --batch 16 --steps 12 --warmup 3 [--budget 25] [--port]

Measurement infrastructure (only bench.py starts it; never imported by the product). One step = what the BTS
training step does on the hot path (reference modules/bts.py:106-108): `silog_loss(0.85)(pred, gt)`, `.backward()`,
`MetricComputation(names).compute(pred, gt)` - the reference's criteria.py / metrics.py loaded by file path
(`oracle/_ref_loader.py`; `kind: "reference"`). When neither /root/reference nor baseline/_ref exists, or with --port,
the oracle's restatement is timed instead (`kind: "port"`). Runs with CUDA hidden: the reference moves work to
`cuda` whenever it is visible. Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--budget", type=float, default=25.0, help="stop the timed loop after this many seconds")
    ap.add_argument("--names", default="delta1,delta2,delta3,mse,mae,log10,rmse")
    ap.add_argument("--port", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    from mono_depth_estimation_b200 import synth
    from oracle import _ref_loader
    names = args.names.split(",")
    shape = (args.batch, 1, 480, 640)
    pred, gt = synth.depth_pair(shape, synth.SEEDS["C2"])
    kind = "port"
    if _ref_loader.available() and not args.port:
        crit = _ref_loader.load("criteria").silog_loss(0.85)
        MC = _ref_loader.load("metrics").MetricComputation
        kind = "reference"

        def step():
            p = pred.detach().clone().requires_grad_(True)
            loss = crit(p, gt)
            loss.backward()
            vals = MC(names).compute(p.detach(), gt)
            return float(loss.detach()) + float(vals[0])
    else:
        from oracle import losses as olosses, metrics as ometrics

        def step():
            loss, grad = olosses.loss_and_grad(olosses.silog, pred, gt, 0.85)
            vals = ometrics.compute(pred, gt, names)
            return float(loss.detach()) + float(vals[0])
    for _ in range(args.warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > args.budget:
            break
    ms = 1e3 * sum(times) / len(times)
    print(json.dumps({"ms_per_step": ms, "steps": len(times), "warmup": args.warmup, "cores": torch.get_num_threads(), "kind": kind,
                      "npx": pred.numel(), "ref_root": _ref_loader.REF_ROOT if kind == "reference" else None}))


if __name__ == "__main__":
    main()
