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

__all__ = ["PeerComm", "shard_range", "reduce_metric_sums", "sharded_eval", "finalize_pooled", "packed_view", "pack_metric_sums", "unpack_metric_sums",
           "SplitLoss", "global_batch_loss", "combine_partials", "loss_from_totals"]


class PeerComm:
    """The ranks' mailboxes for the in-kernel exchange of mde_metrics_sharded (one process per GPU, one box, <= 8 ranks).

    Every rank allocates 8 KB of device memory through the library (cudaMalloc, zeroed), the CUDA IPC handles travel
    once over the process group (`all_gather_object`), every rank maps the others' blocks into its address space
    (NVLink peer access) and hands the pointer table to the library, which keeps it in a small device-resident
    descriptor. From then on an evaluation is ONE launch per rank and nothing else: the launch's finaliser stores the
    rank's 25 doubles into every peer's mailbox and sums the world's rows of its own. The calls are numbered by a counter
    in the descriptor that the finaliser advances (so a launch can be captured in a CUDA graph and replayed); all ranks
    must make the same calls in the same order (as with any collective). Collective constructor; `close()` (collective)
    releases the mappings."""

    def __init__(self, group=None, timeout_ms=2000):
        import ctypes as C
        self._C = C
        self.lib = _lib.load()
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.seq = 0
        self.handle = None
        self._own = None
        self._peers = []
        if self.world == 1:
            return
        if self.world > _lib.MAX_PEERS:
            raise ValueError("PeerComm serves the GPUs of one box (<= %d ranks)" % _lib.MAX_PEERS)
        own = C.c_void_p()
        _lib.check(self.lib.mde_peer_alloc(_lib.PEER_MAILBOX_BYTES, C.byref(own)))
        self._own = own
        hbuf = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(self.lib.mde_peer_export(own, hbuf))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
        table = (C.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                table[r] = own.value
            else:
                p = C.c_void_p()
                _lib.check(self.lib.mde_peer_open(C.create_string_buffer(handles[r], _lib.PEER_HANDLE_BYTES), C.byref(p)))
                self._peers.append(p)
                table[r] = p.value
        comm = C.c_void_p()
        _lib.check(self.lib.mde_peer_comm_create(table, self.rank, self.world, int(timeout_ms), C.byref(comm)))
        self.handle = comm
        dist.barrier(group)          # nobody launches before every mapping exists

    def next_seq(self):
        return 0                                 # 0 = the communicator's own device-resident counter (graph-capturable)

    def all_reduce_(self, src: torch.Tensor, out: torch.Tensor = None, zero_src: bool = False) -> torch.Tensor:
        """Sum of a float64 vector of <= 31 elements over the ranks (C ABI mde_peer_allreduce_f64): one one-warp launch on
        the current stream, result in `out` (default: in place); `zero_src` clears `src` in the same launch. With one
        rank: a copy."""
        assert src.dtype == torch.float64 and src.is_contiguous() and src.numel() <= 31
        out = src if out is None else out
        if self.handle is None:
            if out is not src:
                out.copy_(src)
                if zero_src:
                    src.zero_()
            return out
        dev = src.device
        with torch.cuda.device(dev):
            ws = _lib.workspace(dev, 1)
            _lib.check(self.lib.mde_peer_allreduce_f64(_lib.ptr(src), _lib.ptr(out), src.numel(), 1 if zero_src else 0, self.handle,
                                                       self.next_seq(), _lib.ptr(ws), _lib.stream_ptr(dev)))
        return out

    def close(self):
        if self.handle is None:
            return
        torch.cuda.synchronize()
        dist.barrier(self.group)     # no launch of any rank may still be writing
        for p in self._peers:
            self.lib.mde_peer_close(p)
        self.lib.mde_peer_comm_destroy(self.handle)
        dist.barrier(self.group)     # every mapping is gone before the blocks are freed
        self.lib.mde_peer_free(self._own)
        self.handle, self._own, self._peers = None, None, []


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


class _EvalViews(dict):
    """The evaluation dict as VIEWS of a kernel result vector whose sums already cover the whole set. The views are made
    on first access (a dozen tensor views cost more host time than the launch of an 80-image shard takes on the GPU)."""

    def __init__(self, out_f64, names):
        super().__init__(packed=packed_view(out_f64), work=None)
        self._f64, self._names = out_f64, list(names)

    def __missing__(self, key):
        NM, NQ, f = _lib.METRIC_NM, _lib.METRIC_NQ, self._f64
        idx = [_lib.METRIC_INDEX[n] for n in self._names]
        if key == "image_mean":
            v = {n: f[NM + i] for n, i in zip(self._names, idx)}
        elif key == "pooled":
            v = {n: f[i] for n, i in zip(self._names, idx)}
        elif key == "n_images":
            v = f[2 * NM + NQ]
        elif key == "n_valid":
            v = f[2 * NM]
        elif key == "delta_counts":
            v = f[2 * NM + 1:2 * NM + 4]
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


def _views_of_result(out_f64: torch.Tensor, names):
    return _EvalViews(out_f64, names)


def sharded_eval(pred_shard, target_shard, names, group=None, all_reduce=True, async_op=False, comm=None):
    """Evaluate this rank's images and combine across ranks.

    With `comm` (a PeerComm): ONE kernel launch per rank and nothing else - the launch's finaliser exchanges the 25 doubles
    {pooled raw sums, #valid images, per-image value sums} with the other ranks' launches over NVLink peer memory and
    finishes the values of the whole set itself (C ABI mde_metrics_sharded); the result dict holds views of its output.
    Without: one launch + ONE all-reduce of the same 25 doubles over the process group (NCCL; gloo in the CPU tests of
    the host logic), in place on a view of the kernel's result vector.

    Returns dict: 'image_mean' (reference eval-loop semantics: mean over images of per-image means,
    metrics.py:35-41 + modules/base_module.py:71-76), 'pooled' (dataset-pooled means), 'n_images', 'n_valid',
    'delta_counts' (exact integers in fp64). With `async_op=True` (process-group path) only {'work', 'packed'} come back:
    the all-reduce runs on the communicator's stream while the caller goes on; wait, then unpack_metric_sums(packed, names)."""
    from .metrics import fused_metrics
    if comm is not None and comm.world > 1 and all_reduce:
        return _views_of_result(fused_metrics(pred_shard, target_shard, names=names, comm=comm)["f64"], names)
    multi = all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi and pred_shard.numel() > 0 and not async_op:
        # one rank (or no exchange asked for): the launch's own result vector is final
        return _views_of_result(fused_metrics(pred_shard, target_shard, names=names)["f64"], names)
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


# ---- global-batch loss ----------------------------------------------------------------------------
# Reference-faithful DDP evaluates every loss on the rank's LOCAL sub-batch (pl.Trainer(gpus=N), train.py:137); the
# normalisations of berHu (a max over all pixels, then a mean), SILog (a variance) and the masked means then differ from
# the single-GPU full-batch call. In global-batch mode the ranks exchange the scalar totals between two launches and
# every rank ends with the loss and the gradient of the full batch.
_GB_KINDS = {"l1": _lib.LOSS_L1, "mse": _lib.LOSS_MSE, "berhu": _lib.LOSS_BERHU, "laina_berhu": _lib.LOSS_LAINA_BERHU,
             "silog": _lib.LOSS_SILOG}
_GB_NEEDS_MAX = (_lib.LOSS_BERHU, _lib.LOSS_LAINA_BERHU)


def _kind_of(criterion):
    """(kind, LossParams) of a criterion module of criteria.py, or of a name in _GB_KINDS."""
    from . import criteria as Cr
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    if isinstance(criterion, str):
        return _GB_KINDS[criterion], lp
    if isinstance(criterion, Cr.silog_loss):
        lp.variance_focus = float(criterion.variance_focus)
        return _lib.LOSS_SILOG, lp
    if isinstance(criterion, Cr.LainaBerHuLoss):
        lp.size_average = int(bool(criterion.size_average)); lp.use_logs = int(bool(criterion.use_log)); lp.clamp_val = float(criterion.clamp_val)
        return _lib.LOSS_LAINA_BERHU, lp
    for cls, kind in ((Cr.MaskedL1Loss, _lib.LOSS_L1), (Cr.MaskedMSELoss, _lib.LOSS_MSE), (Cr.berHuLoss, _lib.LOSS_BERHU)):
        if isinstance(criterion, cls):
            return kind, lp
    raise TypeError("global_batch_loss supports MaskedL1Loss, MaskedMSELoss, berHuLoss, LainaBerHuLoss and silog_loss")


class SplitLoss:
    """The three stages of a split-phase loss on ONE shard (C ABI mde_masked_loss_partials / _from_totals). The
    caller combines `partials` across the shards between the stages: MAX of element 4 after stage_max, SUM of
    elements 0..3 after stage_sums. global_batch_loss() below does that with torch.distributed; tests drive the
    stages by hand to emulate several ranks on one GPU."""

    def __init__(self, criterion, pred, target, mask=None):
        import ctypes as C
        self._C = C
        self.lib = _lib.load()
        self.kind, self.lp = _kind_of(criterion)
        self.dev = _lib.require_cuda(pred, target, mask)
        self.pred = pred
        pc = pred.detach()
        if pc.dtype != torch.float32:
            pc = pc.float()                     # fp32 stash (criteria._compute_copy): the gradient is scaled before it is rounded
        self.pc = pc.contiguous()
        t = target.detach()
        if t.shape != pred.shape:
            t = t.expand_as(pred)
        self.t = t.to(torch.float32).contiguous()
        self.mk = None
        if mask is not None:
            mk = mask.detach()
            if mk.shape != pred.shape:
                mk = mk.expand_as(pred)
            self.mk = (mk != 0).to(torch.uint8).contiguous()
        self.partials = torch.tensor([0.0, 0.0, 0.0, 0.0, float("-inf"), 0.0, 0.0, 0.0], dtype=torch.float64, device=self.dev)

    @property
    def needs_max(self):
        return self.kind in _GB_NEEDS_MAX

    def _stage(self, stage):
        if self.pc.numel() == 0:
            return self.partials                # a rank without images contributes the identity and still joins the collectives
        with torch.cuda.device(self.dev):
            gmax = _lib.ptr(self.partials[4:5]) if (stage == 1 and self.needs_max) else None
            _lib.check(self.lib.mde_masked_loss_partials(self.kind, stage, _lib.ptr(self.pc), _lib.dtype_code(self.pc), _lib.ptr(self.t),
                                                         _lib.ptr(self.mk), self.pc.numel(), self._C.byref(self.lp), gmax,
                                                         _lib.ptr(self.partials), _lib.stream_ptr(self.dev)))
        return self.partials

    def stage_max(self):
        """partials[4] = max over this shard (berHu: max(pred - target) over ALL pixels; Laina: max n_i). No-op otherwise."""
        return self._stage(0) if self.needs_max else self.partials

    def stage_sums(self):
        """partials[0..3] += {S0, S1, N0, N1} of this shard, given the GLOBAL max in partials[4]."""
        return self._stage(1)

    def finish(self):
        """Loss of the global batch (0-dim fp32, the same value on every rank, attached to autograd) given the GLOBAL
        totals in `partials`; backward yields dloss/dpred of THIS shard."""
        from .criteria import _fused_apply
        totals, kind, lp, pc, t, mk, dev, lib, C = self.partials, self.kind, self.lp, self.pc, self.t, self.mk, self.dev, self.lib, self._C

        def launch(p, need_grad):
            with torch.cuda.device(dev):
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if (need_grad and pc.numel() > 0) else None
                if pc.numel() == 0:
                    # the loss value still comes from the totals: one-pixel dummy launch would need data; use the host formula on device
                    loss = loss_from_totals(kind, totals, lp, like_kernel=True).to(torch.float32)
                    return loss, (torch.zeros_like(pc) if need_grad else None)
                _lib.check(lib.mde_masked_loss_from_totals(kind, _lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(t), _lib.ptr(mk), pc.numel(),
                                                           C.byref(lp), _lib.ptr(totals), 1.0, _lib.ptr(loss), _lib.ptr(grad),
                                                           _lib.stream_ptr(dev)))
            if grad is not None and grad.shape != p.shape:
                grad = grad.view(p.shape)
            return loss, grad

        return _fused_apply(self.pred, launch)


def loss_from_totals(kind, totals, lp=None, like_kernel=False):
    """Host-side restatement of the coefficient step: the loss value from the global totals {S0, S1, N0, N1, max}
    (torch ops, any device; used for empty shards and by the CPU tests of the exchange logic). `like_kernel` rounds
    where mde_masked_loss_from_totals rounds (reciprocal first, fp32 square root), so that a rank without images reports
    bit for bit the value the other ranks' launches wrote."""
    S0, S1, N0, N1 = totals[0], totals[1], totals[2], totals[3]
    if like_kernel:
        if kind in (_lib.LOSS_L1, _lib.LOSS_MSE):
            return S0 * (1.0 / N0)
        if kind == _lib.LOSS_SILOG:
            vf = float(torch.tensor(0.85 if lp is None else float(lp.variance_focus), dtype=torch.float32))
            inv = 1.0 / N0
            dm, q = S0 * inv, S1 * inv
            return 10.0 * torch.sqrt((q - vf * dm * dm).to(torch.float32)).to(torch.float64)
        if kind == _lib.LOSS_BERHU:
            return (S0 + S1) * (1.0 / (N0 + N1))
        size_average = True if lp is None else bool(lp.size_average)
        return S0 * (1.0 / N0) if size_average else S0
    if kind in (_lib.LOSS_L1, _lib.LOSS_MSE):
        return S0 / N0
    if kind == _lib.LOSS_SILOG:
        vf = 0.85 if lp is None else float(lp.variance_focus)
        dm, q = S0 / N0, S1 / N0
        return 10.0 * torch.sqrt(q - vf * dm * dm)
    if kind == _lib.LOSS_BERHU:
        return (S0 + S1) / (N0 + N1)
    size_average = True if lp is None else bool(lp.size_average)
    return S0 / N0 if size_average else S0


def combine_partials(partials, stage, group=None):
    """The exchange between two stages, in place: stage 0 -> all-reduce(MAX) of partials[4]; stage 1 -> all-reduce(SUM) of
    partials[0:4]. No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return partials
    if stage == 0:
        dist.all_reduce(partials[4:5], op=dist.ReduceOp.MAX, group=group)
    else:
        dist.all_reduce(partials[0:4], op=dist.ReduceOp.SUM, group=group)
    return partials


def global_batch_loss(criterion, pred_shard, target_shard, mask=None, group=None):
    """criterion(pred, target) of the GLOBAL batch whose images are sharded over the ranks of `group`: two (berHu /
    Laina: three) streaming launches per rank and one (two) all-reduce(s) of at most 4 doubles between them. The returned
    loss is identical on every rank; its backward gives the gradient of the global loss with respect to this rank's
    predictions - together the ranks hold exactly the gradient of the single-GPU full-batch call.
    (Under DDP, which AVERAGES parameter gradients over the ranks, scale the loss by the world size.)"""
    sl = SplitLoss(criterion, pred_shard, target_shard, mask)
    if sl.needs_max:
        combine_partials(sl.stage_max(), 0, group)
    combine_partials(sl.stage_sums(), 1, group)
    return sl.finish()
