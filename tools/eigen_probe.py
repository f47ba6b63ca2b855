"""MaskedDepthLoss (criteria.py:17-64) at small sizes: us per launch, fwd+bwd, replayed from a CUDA graph over a ring of
inputs larger than L2.

    python tools/eigen_probe.py                    # register-resident kernel where it applies (csrc/eigen.cu)
    MDE_NO_RESIDENT=1 python tools/eigen_probe.py  # the cooperative kernel for the same calls
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mono_depth_estimation_b200 import _lib, synth  # noqa: E402
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    loss_t = torch.empty((), device=dev)
    for shape in ((8, 1, 228, 304), (4, 1, 228, 304), (2, 1, 480, 640), (1, 1, 240, 320)):
        px = shape[0] * shape[2] * shape[3]
        nring = max(4, int(300e6 // (px * 12)))
        ring = [synth.depth_pair(shape, 101 + i, device=dev) for i in range(nring)]
        grads = [torch.empty(shape, device=dev) for _ in range(nring)]
        ws = _lib.workspace(dev, shape[0])
        for with_grad in (True, False):
            fns = [lambda pr=pr, gt=gt, gr=gr: _lib.check(lib.mde_masked_loss(
                _lib.LOSS_EIGEN, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0,
                _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(gr) if with_grad else None, sp()))
                for (pr, gt), gr in zip(ring, grads)]
            us, graph = bench.graph_timed(fns, dev, 20)
            print(json.dumps({"kind": "eigen", "shape": list(shape), "grad": with_grad, "us_per_launch": round(us, 2), "graph": graph,
                              "resident": not os.environ.get("MDE_NO_RESIDENT")}), flush=True)
        del ring, grads


if __name__ == "__main__":
    main()
