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

__all__ = ["shard_range", "reduce_metric_sums", "sharded_eval", "finalize_pooled", "packed_view", "pack_metric_sums", "unpack_metric_sums"]


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous balanced partition: the first n % world ranks get one extra item
    (654 images over 8 ranks -> 82 x 6 + 81 x 2)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_metric_sums(per_image_values: torch.Tensor, per_image_raw: torch.Tensor) -> torch.Tensor:
    """[NQ + 1 + NM] doubles in the layout of mde_metrics' out_f64[2NM:]: pooled raw sums (integer counts stay exact
    in fp64), #valid images, sum over this rank's valid images of per-image values."""
    valid = per_image_raw[:, _lib.RAW_INDEX["n_valid"]] > 0
    vsum = torch.where(valid[:, None], per_image_values, torch.zeros_like(per_image_values)).sum(0)
    return torch.cat([per_image_raw.sum(0), valid.sum().to(torch.float64).reshape(1), vsum])


def packed_view(out_f64: torch.Tensor) -> torch.Tensor:
    """The same vector as a VIEW of a kernel result (mde_metrics' out_f64): the launch already formed it, so a rank
    all-reduces it in place with no packing launches at all."""
    return out_f64[2 * _lib.METRIC_NM:3 * _lib.METRIC_NM + _lib.METRIC_NQ + 1]


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
    NM, NQ = _lib.METRIC_NM, _lib.METRIC_NQ
    raw, n_img, vsum = packed[:NQ], packed[NQ], packed[NQ + 1:NQ + 1 + NM]
    image_mean = vsum / n_img
    pooled = finalize_pooled(raw)
    idx = [_lib.METRIC_INDEX[n] for n in names]
    return {"image_mean": {n: image_mean[i] for n, i in zip(names, idx)},
            "pooled": {n: pooled[i] for n, i in zip(names, idx)},
            "n_images": n_img, "n_valid": raw[0], "delta_counts": raw[1:4]}


def sharded_eval(pred_shard, target_shard, names, group=None, all_reduce=True, async_op=False):
    """Evaluate this rank's images and combine across ranks: ONE kernel launch per rank, ONE all-reduce of
    NQ + 1 + NM = 25 doubles (in place on a view of the kernel's result vector), nothing else.

    Returns dict: 'image_mean' (reference eval-loop semantics: mean over images of per-image means,
    metrics.py:35-41 + modules/base_module.py:71-76), 'pooled' (dataset-pooled means), 'n_images', 'n_valid',
    'delta_counts' (exact integers in fp64). With `async_op=True` only {'work', 'packed'} come back: the all-reduce
    runs on the communicator's stream while the caller goes on; wait, then unpack_metric_sums(packed, names)."""
    from .metrics import fused_metrics
    if pred_shard.numel() == 0:
        # a rank without images (n_items < world) contributes zeros and STILL enters the collective
        packed = torch.zeros(_lib.METRIC_NM + 1 + _lib.METRIC_NQ, dtype=torch.float64, device=pred_shard.device)
    else:
        packed = packed_view(fused_metrics(pred_shard, target_shard, names=names)["f64"])
    work = None
    if all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if async_op:   # finish with: r["work"].wait(); unpack_metric_sums(r["packed"], names)
        return {"work": work, "packed": packed}
    out = unpack_metric_sums(packed, names)
    out["packed"] = packed
    return out
