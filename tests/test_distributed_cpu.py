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
