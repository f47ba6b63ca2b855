"""Host-side logic of the image-sharded evaluation, world_size 2 over gloo on CPU: the partition, the
packed all-reduce vector and the un-packing must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mono_depth_estimation_b200 import distributed as D
from mono_depth_estimation_b200 import synth
from oracle import metrics as ometrics

NAMES = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse", "absrel", "sqrel", "msle"]
ALL = ["delta1", "delta2", "delta3", "mae", "mse", "log10", "msle", "absrel", "sqrel", "rmse", "rmse_true", "rmse_log"]


def _per_image(pred, gt):
    vals, raws = [], []
    for b in range(pred.shape[0]):
        p, t = pred[b:b + 1].double(), gt[b:b + 1].double()
        if (t > 0).sum() == 0:
            vals.append(torch.full((12,), float("nan"), dtype=torch.float64))
            raws.append(torch.zeros(12, dtype=torch.float64))
            continue
        vals.append(torch.tensor([float(v) for v in ometrics.compute(p, t, ALL)], dtype=torch.float64))
        raws.append(ometrics.raw_sums(p, t))
    return torch.stack(vals), torch.stack(raws)


def test_shard_range_nyu():
    sizes = [D.shard_range(654, r, 8) for r in range(8)]
    assert [b - a for a, b in sizes] == [82] * 6 + [81] * 2
    assert sizes[0][0] == 0 and sizes[-1][1] == 654
    assert all(sizes[i][1] == sizes[i + 1][0] for i in range(7))
    assert [b - a for a, b in (D.shard_range(654, r, 4) for r in range(4))] == [164, 164, 163, 163]
    assert [b - a for a, b in (D.shard_range(3, r, 8) for r in range(8))] == [1, 1, 1, 0, 0, 0, 0, 0]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred, gt = synth.depth_pair((7, 1, 24, 32), 61, border=2)
    gt[3] = 0                                                          # an image without a valid pixel
    a, b = D.shard_range(pred.shape[0], rank, world)
    vals, raws = _per_image(pred[a:b], gt[a:b])
    packed = D.reduce_metric_sums(D.pack_metric_sums(vals, raws))
    out = D.unpack_metric_sums(packed, NAMES)
    q.put((rank, {k: float(v) for k, v in out["image_mean"].items()}, {k: float(v) for k, v in out["pooled"].items()},
           float(out["n_images"]), [float(c) for c in out["delta_counts"]]))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pred, gt = synth.depth_pair((7, 1, 24, 32), 61, border=2)
    gt[3] = 0
    keep = [i for i in range(7) if i != 3]
    ref_mean = ometrics.compute_per_image_mean(pred[keep].double(), gt[keep].double(), NAMES)
    ref_pool = [float(v) for v in ometrics.compute(pred.double(), gt.double(), NAMES)]
    n, c1, c2, c3 = ometrics.delta_counts(pred, gt)
    for rank, image_mean, pooled, n_images, counts in results:
        assert n_images == 6.0
        np.testing.assert_allclose([image_mean[k] for k in NAMES], ref_mean, rtol=1e-12)
        np.testing.assert_allclose([pooled[k] for k in NAMES], ref_pool, rtol=1e-6)
        assert counts == [float(c1), float(c2), float(c3)]


# ---- global-batch loss: the exchange logic over gloo (the kernels are replaced by torch restatements of the partial totals) ----
def _partials_cpu(kind, pred, gt, gmax=None):
    """{S0, S1, N0, N1, max} of one shard as the split-phase kernels define them (csrc/losses_split.cu), in torch."""
    p, t = pred.double(), gt.double()
    out = torch.tensor([0.0, 0.0, 0.0, 0.0, float("-inf"), 0.0, 0.0, 0.0], dtype=torch.float64)
    if kind == "silog":
        m = t > 0.01
        d = torch.log(p[m]) - torch.log(t[m])
        out[0], out[1], out[2] = d.sum(), (d * d).sum(), m.sum()
    elif kind == "berhu":
        if gmax is None:
            out[4] = (p - t).max() if p.numel() else float("-inf")
            return out
        v = t > 0
        ad = (t - p).abs()[v]
        hub = ad > 0.2 * gmax
        out[0], out[1], out[2], out[3], out[4] = ad.sum(), (ad[hub] ** 2).sum(), v.sum(), hub.sum(), gmax
    return out


def _gb_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mono_depth_estimation_b200 import _lib
    pred, gt = synth.depth_pair((5, 1, 24, 32), 62, border=2)
    pred[4, 0, 5, 5] += 20.0                                       # the berHu maximum lives on the last rank
    a, b = D.shard_range(pred.shape[0], rank, world)
    out = {}
    part = _partials_cpu("silog", pred[a:b], gt[a:b])
    D.combine_partials(part, 1)
    out["silog"] = float(D.loss_from_totals(_lib.LOSS_SILOG, part))
    part = _partials_cpu("berhu", pred[a:b], gt[a:b])
    D.combine_partials(part, 0)
    part = _partials_cpu("berhu", pred[a:b], gt[a:b], gmax=float(part[4]))
    D.combine_partials(part, 1)
    out["berhu"] = float(D.loss_from_totals(_lib.LOSS_BERHU, part))
    q.put((rank, out))
    dist.destroy_process_group()


def test_global_batch_exchange_over_gloo():
    """Two ranks, images sharded 3 + 2: MAX then SUM of the partial totals gives every rank the full-batch loss."""
    from oracle import losses as olosses
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gb_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pred, gt = synth.depth_pair((5, 1, 24, 32), 62, border=2)
    pred[4, 0, 5, 5] += 20.0
    want = {n: float(olosses.loss_and_grad(olosses.LOSSES[n], pred.double(), gt.double())[0]) for n in ("silog", "berhu")}
    for rank, out in results:
        for n in want:
            np.testing.assert_allclose(out[n], want[n], rtol=1e-12, err_msg="%s rank %d" % (n, rank))
