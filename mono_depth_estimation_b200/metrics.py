"""Drop-in for the reference's metrics.py (MetricLogger, MetricComputation, METRICS) on B200.

Same names, arguments, return types and error behaviour as reference metrics.py:11-123; the
arithmetic runs in ONE fused CUDA kernel (csrc/metrics.cu via the C ABI `mde_metrics`) instead
of ~20 boolean gathers + ~40 elementwise/reduce launches per call.

Extensions that do not exist in the reference (all opt-in):
  * MetricComputation.compute_batch(pred, target): reference-faithful EVAL-LOOP semantics for a
    whole batch in one launch - the unweighted mean over images of per-image means (what the
    reference's test loop produces with its batch size 1);
  * keys 'rmse_true' and 'rmse_log' (the reference's 'rmse' is mean(sqrt((p-t)^2/t)), metrics.py:106-109);
  * strict=False skips the device->host read that the reference's `assert sum(valid) > 0` implies.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib

__all__ = ["MetricLogger", "MetricComputation", "METRICS", "fused_metrics", "fused_metrics_resized"]


def _prep(pred, target):
    dev = _lib.require_cuda(pred, target)
    if pred.shape != target.shape:
        pred, target = torch.broadcast_tensors(pred, target)
    if pred.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        pred = pred.float()
    pred = pred.detach().contiguous()
    target = target.detach()
    if target.dtype != torch.float32:
        target = target.float()
    target = target.contiguous()
    return dev, pred, target


def fused_metrics(pred, target, names=None, per_image=False, reference_math=False, image_dims=2, comm=None):
    """Launch the fused kernel once.

    `comm` (distributed.PeerComm, world > 1): the launch's finaliser exchanges this rank's sums with the other ranks'
    launches over NVLink peer memory (C ABI mde_metrics_sharded), so 'values' / 'image_mean' / 'f64' describe the WHOLE
    sharded set, identically on every rank; 'per_image*' stay local. An empty shard is legal then. Every rank must call.

    Returns a dict with
      'values'      fp32 [NM] pooled over all valid pixels of the call (reference compute() semantics)
      'image_mean'  fp32 [NM] mean over images of per-image means
      'f64'         float64 [2NM + NQ + 1] (pooled values, image-mean values, pooled raw sums, #valid images)
      'per_image'   float64 [n_img, NM] (only if per_image)
      'per_image_raw' float64 [n_img, NQ] (only if per_image)
    `image_dims` trailing dims form one image (2 -> H,W); every leading dim counts as an image.
    """
    lib = _lib.load()
    dev, pred, target = _prep(pred, target)
    if pred.dim() < image_dims:
        image_dims = pred.dim()
    hw = 1
    for s in pred.shape[pred.dim() - image_dims:]:
        hw *= int(s)
    n_img = pred.numel() // max(hw, 1)
    exchange = comm is not None and comm.world > 1
    if pred.numel() == 0 and not exchange:
        raise AssertionError("invalid target!")
    flags = _lib.METRICS_REFERENCE_MATH if reference_math else 0
    if names is not None:
        g = 0
        for n in names:
            g |= _lib.METRIC_GROUP.get(n, 0)
        if g == 0:
            g = _lib.METRICS_NEED_LOG  # 0 would mean "all"; request the cheapest single group instead
        flags |= g
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, n_img)
        out64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
        out32 = torch.empty(2 * _lib.METRIC_NM, dtype=torch.float32, device=dev)
        piv = pir = None
        if per_image:
            piv = torch.empty((n_img, _lib.METRIC_NM), dtype=torch.float64, device=dev)
            pir = torch.empty((n_img, _lib.METRIC_NQ), dtype=torch.float64, device=dev)
        if exchange:
            _lib.check(lib.mde_metrics_sharded(_lib.ptr(pred) if n_img else None, _lib.dtype_code(pred), _lib.ptr(target) if n_img else None,
                                               n_img, max(hw, 1), flags, _lib.ptr(ws), _lib.ptr(out64), _lib.ptr(out32), _lib.ptr(piv),
                                               _lib.ptr(pir), comm.handle, comm.next_seq(), _lib.stream_ptr(dev)))
        else:
            _lib.check(lib.mde_metrics(_lib.ptr(pred), _lib.dtype_code(pred), _lib.ptr(target), n_img, hw, flags,
                                       _lib.ptr(ws), _lib.ptr(out64), _lib.ptr(out32), _lib.ptr(piv), _lib.ptr(pir),
                                       _lib.stream_ptr(dev)))
    res = {"values": out32[:_lib.METRIC_NM], "image_mean": out32[_lib.METRIC_NM:], "f64": out64}
    if per_image:
        res["per_image"] = piv
        res["per_image_raw"] = pir
    return res


def fused_metrics_resized(pred, target, size=(480, 640), per_image=False):
    """Metrics of `pred` and `target` after BOTH are bilinearly resized to `size` - what the test steps of the eigen,
    dorn and my modules do with two F.interpolate(mode='bilinear') calls before log_test (reference
    modules/eigen.py:49-51, modules/dorn.py:181-183, modules/my.py:64-66) - in one launch that samples the sources on
    the fly (C ABI mde_metrics_resized). pred [B,1,h,w] / target [B,1,h',w'] (or [B,h,w]); same dict as fused_metrics."""
    lib = _lib.load()
    dev = _lib.require_cuda(pred, target)
    p = pred.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    ph, pw = int(p.shape[-2]), int(p.shape[-1])
    th, tw = int(t.shape[-2]), int(t.shape[-1])
    n_img = p.numel() // (ph * pw)
    if t.numel() // (th * tw) != n_img:
        raise AssertionError("inconsistent dimensions")
    oh, ow = int(size[0]), int(size[1])
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, n_img)
        out64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
        out32 = torch.empty(2 * _lib.METRIC_NM, dtype=torch.float32, device=dev)
        piv = pir = None
        if per_image:
            piv = torch.empty((n_img, _lib.METRIC_NM), dtype=torch.float64, device=dev)
            pir = torch.empty((n_img, _lib.METRIC_NQ), dtype=torch.float64, device=dev)
        _lib.check(lib.mde_metrics_resized(_lib.ptr(p), ph, pw, _lib.ptr(t), th, tw, n_img, oh, ow, 0, _lib.ptr(ws),
                                           _lib.ptr(out64), _lib.ptr(out32), _lib.ptr(piv), _lib.ptr(pir), _lib.stream_ptr(dev)))
    res = {"values": out32[:_lib.METRIC_NM], "image_mean": out32[_lib.METRIC_NM:], "f64": out64}
    if per_image:
        res["per_image"] = piv
        res["per_image_raw"] = pir
    return res


def _single(name):
    idx = _lib.METRIC_INDEX[name]

    def fn(pred, target):
        """metric(pred_1d, target_1d) on already gathered vectors (reference metrics.py:75-109)."""
        if pred.numel() == 0:
            return torch.full((), float("nan"), device=pred.device)
        return fused_metrics(pred.reshape(1, -1), target.reshape(1, -1), names=[name], image_dims=1)["values"][idx]

    fn.__name__ = name
    return fn


def _ssim(pred, target):
    # reference metrics.py:123 delegates to torchmetrics (third party, windowed filter): out of scope here
    try:
        import torchmetrics
    except ImportError as e:  # pragma: no cover
        raise NotImplementedError("'ssim' needs torchmetrics (reference metrics.py:123); it is outside the "
                                  "per-pixel hot path this package replaces") from e
    return torchmetrics.functional.structural_similarity_index_measure(pred, target)


METRICS = {name: _single(name) for name in _lib.METRIC_INDEX}
METRICS["ssim"] = _ssim


class MetricComputation(object):
    """reference metrics.py:47-72."""

    def __init__(self, metrics, strict=True, reference_math=False):
        self.names = metrics
        self.metrics = [METRICS[m] for m in metrics]  # KeyError for unknown names, as in the reference
        self.metric_names = metrics
        self.strict = strict
        self.reference_math = reference_math
        self._fused_names = [m for m in metrics if m != "ssim"]
        self._prefetched = None
        self.last_f64 = None   # float64 result vector of the last compute() (layout of mde_metrics' out_f64)
        self.reset()

    # ---- hand-over from a criterion that computed the pooled metrics in its own launch -----------------
    def group_flags(self):
        g = 0
        for n in self._fused_names:
            g |= _lib.METRIC_GROUP.get(n, 0)
        return g if g else _lib.METRICS_NEED_LOG

    @staticmethod
    def _ident(t):
        return (t.data_ptr(), t._version, tuple(t.shape), t.dtype, t.device)

    def offer(self, pred, target, values32, f64, booked=False):
        """A criterion computed the pooled metrics of (pred, target) in its own launch. The offer holds WEAK
        references to the two tensors: it is honoured only while both are still alive (their storage cannot have been
        freed and handed to other data) and unmodified (same version counter), and only by the next compute().
        `booked`: the launch already added the values to the running sums (accum_buffer) - count the call now."""
        self._prefetched = (weakref.ref(pred), self._ident(pred), weakref.ref(target), self._ident(target), values32, f64, booked)
        if booked:
            self._sum_assigned = None
            self.count += 1

    def _take_prefetched(self, pred, target):
        pf, self._prefetched = self._prefetched, None
        if pf is None:
            return None
        rp, idp, rt, idt, values32, f64, booked = pf
        if rp() is None or rt() is None:       # the offered tensors are gone: their addresses may have been reused
            return None
        if self._ident(rp()) != idp or self._ident(rt()) != idt:   # modified in place since the offer
            return None
        # compute() may be handed views / detached aliases of the offered tensors (log_train(y_hat.detach(), y)):
        # same storage address, version counter (shared by aliases), shape and dtype
        if self._ident(pred) != idp or self._ident(target) != idt:
            return None
        return {"values": values32, "f64": f64, "booked": booked}

    def raw_accum_buffer(self, device):
        """fp64 device vector [NQ] of pooled RAW sums (valid count, exact delta counts, float sums) that a booking
        criterion's launches add to: what a rank pools between two all-reduces of the metric sums (distributed.py).
        Created (zeroed) on first use; the caller zeroes it after each exchange."""
        if getattr(self, "_raw_vec", None) is None:
            self._raw_vec = torch.zeros(_lib.METRIC_NQ, dtype=torch.float64, device=device)
        return self._raw_vec

    def accum_buffer(self, device):
        """The running sums as ONE fp32 device vector [NM] (created on first use) - what a fusing criterion's launch
        adds to when it books the metrics itself."""
        if self._sum_vec is None:
            self._sum_vec = torch.zeros(_lib.METRIC_NM, dtype=torch.float32, device=device)
        return self._sum_vec

    def reset(self):
        self.count = 0
        self._sum_vec = None                 # running sums of the fused metrics: ONE device vector
        self._sum_other = {}                 # running sums of anything else ('ssim')
        self._sum_assigned = None            # a caller may assign .sum directly, as with the reference's plain list

    @property
    def sum(self):
        """Running sums in `names` order (reference metrics.py:56,65-66), as 0-dim views of one vector."""
        if self._sum_assigned is not None:
            return self._sum_assigned
        out = []
        for n in self.metric_names:
            if n in _lib.METRIC_INDEX:
                out.append(0.0 if self._sum_vec is None else self._sum_vec[_lib.METRIC_INDEX[n]])
            else:
                out.append(self._sum_other.get(n, 0.0))
        return out

    @sum.setter
    def sum(self, value):
        self._sum_assigned = value

    def _collect(self, res_vec, pred, target, booked=False):
        self._sum_assigned = None
        vals = []
        for n in self.metric_names:
            if n == "ssim":
                v = _ssim(torch.clamp_min(pred, 1e-07).cpu(), target.cpu())  # metrics.py:63
                self._sum_other[n] = self._sum_other.get(n, 0.0) + v
                vals.append(v)
            else:
                vals.append(res_vec[_lib.METRIC_INDEX[n]])
        if booked:      # the criterion's launch added the values and offer() counted the call
            return vals
        self.count += 1
        # one 12-float add instead of one launch per metric
        if self._sum_vec is None:
            self._sum_vec = res_vec.clone()
        else:
            self._sum_vec += res_vec
        return vals

    def compute(self, pred, target):
        """One mean per metric over the valid pixels of the WHOLE call tensor (metrics.py:58-67)."""
        with torch.no_grad():
            res = self._take_prefetched(pred, target)
            if res is None:
                res = fused_metrics(pred, target, names=self._fused_names, reference_math=self.reference_math)
            self.last_f64 = res["f64"]
            if self.strict:
                # the reference's `assert torch.sum(valid_mask) > 0` (metrics.py:61) reads the device too
                assert float(res["f64"][2 * _lib.METRIC_NM + _lib.RAW_INDEX["n_valid"]]) > 0, "invalid target!"
            return self._collect(res["values"], pred, target, booked=res.get("booked", False))

    def compute_resized(self, pred, target, size=(480, 640)):
        """compute() on pred and target bilinearly resized to `size` first (the test_step pattern of the eigen, dorn
        and my modules), without materialising the resized tensors (extension; see fused_metrics_resized)."""
        with torch.no_grad():
            res = fused_metrics_resized(pred, target, size)
            self.last_f64 = res["f64"]
            if self.strict:
                assert float(res["f64"][2 * _lib.METRIC_NM + _lib.RAW_INDEX["n_valid"]]) > 0, "invalid target!"
            return self._collect(res["values"], pred, target)

    def compute_batch(self, pred, target):
        """Mean over images of per-image means for a [B,...,H,W] batch in one launch (extension)."""
        with torch.no_grad():
            res = fused_metrics(pred, target, names=self._fused_names, reference_math=self.reference_math)
            self.last_f64 = res["f64"]
            if self.strict:
                assert float(res["f64"][2 * _lib.METRIC_NM + _lib.METRIC_NQ]) > 0, "invalid target!"
            return self._collect(res["image_mean"], pred, target)

    def avg(self, metric):
        if isinstance(metric, int):
            return self.sum[metric] / self.count
        if isinstance(metric, str):
            return self.sum[self.names.index(metric)] / self.count
        assert False, "metric must be int or str"


class MetricLogger(object):
    """Lightning logging adaptor with the reference's interface (metrics.py:11-44): `.context` is the
    LightningModule whose .log() receives the values, `.computer` the MetricComputation.

    Logged keys (identical to the reference so dashboards / checkpoints monitors keep working):
      train:  "loss", "train_<m>" (logger, on_epoch), "train_<m>(AVG)" (progress bar only)
      val:    "val_<prefix><m>" (logger, on_epoch),  "val_<prefix><m>(AVG)" (progress bar only)
      test:   "<m>" (on_step and on_epoch)
    """

    def __init__(self, metrics, module):
        self.context = module
        self.computer = MetricComputation(metrics)

    def _emit(self, pred, target, key, avg_key, result_key, **log_kwargs):
        out = {}
        for name, value in zip(self.computer.names, self.computer.compute(pred, target)):
            self.context.log(key.format(name), value, **log_kwargs)
            if avg_key is not None:
                self.context.log(avg_key.format(name), self.computer.avg(name), logger=False, prog_bar=True)
            out[result_key.format(name)] = value
        return out

    def log_train(self, pred, target, loss):
        self.context.log("loss", loss)
        result = {"loss": loss}
        result.update(self._emit(pred, target, "train_{}", "train_{}(AVG)", "{}", logger=True, on_epoch=True))
        return result

    def log_val(self, pred, target, prefix=''):
        return self._emit(pred, target, "val_" + prefix + "{}", "val_" + prefix + "{}(AVG)", prefix + "{}",
                          logger=True, on_epoch=True)

    def log_test(self, pred, target):
        return self._emit(pred, target, "{}", None, "{}", on_step=True, on_epoch=True)

    def reset(self):
        self.computer.reset()
