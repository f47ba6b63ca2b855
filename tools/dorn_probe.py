"""DORN head at C3 (8x136x257x353, K = 68): us per launch of the fused supervision step and of the decode-only pass,
replayed from a CUDA graph over two input sets (2 x 395 MB of logits > L2). Select a library build with MDE_B200_LIB
(tools/build_variant-style A/B on ONE box: box-to-box spread is +-3 %)."""
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
    shape = synth.SHAPES["C3"]
    N, C2, H, W = shape
    px = N * H * W
    loss_t = torch.empty((), device=dev)
    ring = [synth.dorn_inputs(shape, 103 + i, device=dev) for i in range(2)]
    dec = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev)
    dep = torch.empty((N, 1, H, W), device=dev)
    gxs = [torch.empty(shape, device=dev) for _ in range(2)]
    ws1 = _lib.workspace(dev, 1)
    fns = [lambda x=x, gt=gt, gx=gx: _lib.check(lib.mde_dorn_fused(_lib.ptr(x), 0, _lib.ptr(gt), N, C2 // 2, H * W, 0.001, 1.0, 0, 1.0,
                                                               _lib.ptr(ws1), _lib.ptr(loss_t), None, _lib.ptr(dec), _lib.ptr(dep),
                                                               _lib.ptr(gx), sp())) for (x, gt), gx in zip(ring, gxs)]
    us, _ = bench.graph_timed(fns, dev, 20)
    out = {"lib": os.path.basename(os.environ.get("MDE_B200_LIB", "libmde_b200.so")), "fused_us": round(us, 2),
           "fused_gbs": round((8.0 * C2 + 16.0) * px / us / 1e3, 1), "loss": float(loss_t), "decode_sum": int(dec.sum()),
           "gx_abs_sum": float(gxs[0].double().abs().sum())}
    fns = [lambda x=x: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, C2 // 2, H * W, None, _lib.ptr(dec), sp())) for x, _ in ring]
    us, _ = bench.graph_timed(fns, dev, 20)
    out.update({"decode_us": round(us, 2), "decode_gbs": round((4.0 * C2 + 8.0) * px / us / 1e3, 1)})
    # the reference's own call sequence (module API): layer forward with P, ordLoss(P, label) fwd+bwd, layer backward
    x, gt = ring[0]
    del gxs
    K = C2 // 2
    P = torch.empty((N, K, H, W), device=dev)
    gP = torch.empty((N, K, H, W), device=dev)
    gx = torch.empty(shape, device=dev)
    lab = torch.rand((N, 1, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * K
    rows = (("layer_fwd_prob", [lambda: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, K, H * W, _lib.ptr(P), _lib.ptr(dec), sp()))],
             (8.0 * K + 4.0 * K + 8.0) * px),
            ("ord_loss", [lambda: _lib.check(lib.mde_ord_loss(_lib.ptr(P), _lib.ptr(lab), N, K, H * W, 1.0, _lib.ptr(ws1), _lib.ptr(loss_t),
                                                              _lib.ptr(gP), sp()))], (8.0 * K + 4.0) * px),
            ("layer_bwd", [lambda: _lib.check(lib.mde_ordinal_layer_bwd(_lib.ptr(x), 0, _lib.ptr(gP), N, K, H * W, _lib.ptr(gx), sp()))],
             (8.0 * K + 4.0 * K + 8.0 * K) * px))
    for name, fns, nbytes in rows:
        us, _ = bench.graph_timed(fns, dev, 20)
        out[name + "_us"] = round(us, 2)
        out[name + "_gbs"] = round(nbytes / us / 1e3, 1)
    out["unfused_check"] = [float(loss_t), float(gP.double().abs().sum()), float(gx.double().abs().sum()), float(P.double().sum())]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
