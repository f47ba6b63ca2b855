"""Image-sharded evaluation / supervision across the GPUs of one box.

The path shards by image with NO data-path collective: every rank runs the fused kernels on its
own images. The only exchange is one all-reduce of a few dozen doubles (per-rank sums of per-image
metric values + image count, pooled raw sums and counts) - latency-bound, over NCCL/NVLink on GPUs
and over gloo in the CPU tests of the host logic.

The reference itself never synchronises metrics (`self.log` is called without sync_dist,
metrics.py:19-39: each rank averages its own shard); `all_reduce=False` reproduces that.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib

__all__ = ["shard_range", "reduce_metric_sums", "sharded_eval", "finalize_pooled"]


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous balanced partition: the first n % world ranks get one extra item
    (654 images over 8 ranks -> 82 x 6 + 81 x 2)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_metric_sums(per_image_values: torch.Tensor, per_image_raw: torch.Tensor) -> torch.Tensor:
    """[NM + 1 + NQ] doubles: sum over this rank's valid images of per-image values, #valid images,
    pooled raw sums (integer counts stay exact in fp64)."""
    valid = per_image_raw[:, _lib.RAW_INDEX["n_valid"]] > 0
    vsum = torch.where(valid[:, None], per_image_values, torch.zeros_like(per_image_values)).sum(0)
    return torch.cat([vsum, valid.sum().to(torch.float64).reshape(1), per_image_raw.sum(0)])


def reduce_metric_sums(packed: torch.Tensor, group=None) -> torch.Tensor:
    """One all-reduce(sum) of the packed vector (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def finalize_pooled(raw: torch.Tensor) -> torch.Tensor:
    """Metric values [NM] from pooled raw sums [NQ] (same arithmetic as mde_metrics_finalize_host)."""
    n = raw[0]
    r = raw
    return torch.stack([r[1] / n, r[2] / n, r[3] / n, r[4] / n, r[5] / n, r[6] / n, r[7] / n, r[8] / n, r[9] / n,
                        r[10] / n, torch.sqrt(r[5] / n), torch.sqrt(r[11] / n)])


def unpack_metric_sums(packed: torch.Tensor, names):
    NM = _lib.METRIC_NM
    image_mean = packed[:NM] / packed[NM]
    pooled = finalize_pooled(packed[NM + 1:])
    idx = [_lib.METRIC_INDEX[n] for n in names]
    return {"image_mean": {n: image_mean[i] for n, i in zip(names, idx)},
            "pooled": {n: pooled[i] for n, i in zip(names, idx)},
            "n_images": packed[NM], "n_valid": packed[NM + 1], "delta_counts": packed[NM + 2:NM + 5]}


def sharded_eval(pred_shard, target_shard, names, group=None, all_reduce=True):
    """Evaluate this rank's images and combine across ranks.

    Returns dict: 'image_mean' (reference eval-loop semantics: mean over images of per-image means),
    'pooled' (dataset-pooled means), 'n_images', 'n_valid', 'delta_counts' (exact integers in fp64)."""
    from .metrics import fused_metrics
    res = fused_metrics(pred_shard, target_shard, names=names, per_image=True)
    packed = pack_metric_sums(res["per_image"], res["per_image_raw"])
    if all_reduce:
        packed = reduce_metric_sums(packed, group)
    return unpack_metric_sums(packed, names)
