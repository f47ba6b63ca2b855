"""Drop-in for the hot subset of the reference's criteria.py on B200.

Every class keeps the reference's name, constructor arguments, forward signature, `self.loss`
side effect, error behaviour and "no tensors in state_dict" property; the forward and
the backward of each loss run fused in ONE cooperative CUDA launch (csrc/losses.cu, eigen.cu,
dorn.cu, vnl.cu through the C ABI of include/mde_b200.h).

  MaskedDepthLoss        reference criteria.py:17-64
  MaskedMSELoss          reference criteria.py:67-77
  MaskedL1Loss           reference criteria.py:80-90
  berHuLoss              reference criteria.py:111-133
  LainaBerHuLoss         reference criteria.py:476-506
  silog_loss             reference criteria.py:724-732
  ordLoss                reference criteria.py:734-787
  OrdinalRegressionLoss  reference criteria.py:789-836
  VNL_Loss               reference criteria.py:866-1045
  ModelLoss              reference criteria.py:1047-1062 (VNL part on the kernels; WCEL_Loss: wcel.py)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["MaskedDepthLoss", "MaskedMSELoss", "MaskedL1Loss", "berHuLoss", "LainaBerHuLoss", "silog_loss",
           "ordLoss", "OrdinalRegressionLoss", "VNL_Loss", "ModelLoss", "masked_loss",
           "compute_scale_and_shift", "scale_shift", "MidasLoss", "TrimmedProcrustesLoss",
           "normalize_prediction_robust"]


def _scale_grad(grad, grad_output):
    """grad *= grad_output on the device (the kernel returns at once when grad_output == 1)."""
    lib = _lib.load()
    if grad.numel() == 0:
        return grad
    with torch.cuda.device(grad.device):   # backward may run with another current device (one process, several GPUs)
        go = grad_output.detach().to(device=grad.device, dtype=torch.float32).reshape(1).contiguous()
        _lib.check(lib.mde_scale_inplace(_lib.ptr(grad), _lib.dtype_code(grad), grad.numel(), _lib.ptr(go),
                                         _lib.stream_ptr(grad.device)))
    return grad


def _compute_copy(p):
    """The tensor the kernels read. Half-precision predictions (AMP, reference train.py:60,139) are widened to fp32
    first: the kernels store dloss/dpred BEFORE autograd's grad_output is known, i.e. values of order 1/N, which
    underflow fp16 (1/N = 2e-7 at C2, fp16 subnormal step 6e-8) and would defeat GradScaler's loss scaling. With an
    fp32 stash the scale is applied in fp32 and the result is rounded to the prediction's dtype once, as autograd does."""
    pc = p.detach()
    if pc.dtype in (torch.float16, torch.bfloat16):
        pc = pc.float()
    return pc.contiguous()


class _FusedLossFn(torch.autograd.Function):
    """Generic wrapper: `launch(pred, need_grad)` returns (loss 0-dim fp32, grad or None)."""

    @staticmethod
    def forward(ctx, pred, launch, flag=None):
        need_grad = ctx.needs_input_grad[0]
        loss, grad = launch(pred, need_grad)
        ctx.grad = grad
        ctx.pred_dtype = pred.dtype
        ctx.used = False
        ctx.flag = flag
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grad is None:
            return None, None, None
        if ctx.used:
            raise RuntimeError("the fused loss gradient was already consumed; run the forward again "
                               "(retain_graph is not supported by the fused forward+backward kernel)")
        ctx.used = True
        # `loss.backward()` on the criterion's own result: the root gradient is autograd's implicit ones, known on the
        # host (see _fused_apply) - the stashed gradient IS the answer, no launch
        unit = ctx.flag is not None and ctx.flag.unit
        g = ctx.grad if unit else _scale_grad(ctx.grad, grad_output)
        ctx.grad = None
        if g.dtype != ctx.pred_dtype:
            g = g.to(ctx.pred_dtype)          # fp32 stash of a half-precision prediction: one rounding, after the scale
        return g, None, None


class _UnitFlag:
    """Set while `loss.backward()` (no explicit gradient) runs on the tensor a fused criterion returned."""
    __slots__ = ("unit",)

    def __init__(self):
        self.unit = False


class FusedLoss(torch.Tensor):
    """The 0-dim loss a fused criterion returns: a torch.Tensor whose `.backward()` tells the criterion's autograd node
    that the root gradient is the implicit `ones_like(loss)`. Every operation on it returns a plain Tensor."""

    __torch_function__ = torch._C._disabled_torch_function_impl

    def backward(self, gradient=None, retain_graph=None, create_graph=False, inputs=None):
        flag = self.__dict__.get("_mde_flag")
        if flag is None:
            return super().backward(gradient, retain_graph, create_graph, inputs=inputs)
        flag.unit = gradient is None and not create_graph
        if flag.unit:
            # autograd would materialise `ones_like(loss)` with a fill launch per call; a persistent scalar 1 per device
            # serves as the explicit root gradient instead (the value is never read: the flag says what it is)
            gradient = _unit_gradient(self)
        try:
            return super().backward(gradient, retain_graph, create_graph, inputs=inputs)
        finally:
            flag.unit = False


_UNIT_ONES = {}


def _unit_gradient(loss):
    key = (loss.device, loss.dtype)
    one = _UNIT_ONES.get(key)
    if one is None:
        if loss.is_cuda and torch.cuda.is_current_stream_capturing():
            return None                      # never allocate the shared scalar from a graph's private pool
        one = torch.ones((), device=loss.device, dtype=loss.dtype)
        _UNIT_ONES[key] = one
    return one


def _fused_apply(pred, launch):
    """_FusedLossFn.apply + the host-side knowledge that saves the backward's launch in the common case.

    The fused kernels store dloss/dpred in the forward launch; autograd's grad_output is a DEVICE scalar, so its value
    is not known on the host and `mde_scale_inplace` has to be launched to apply it - even when it is the implicit
    `ones_like(loss)` of a plain `loss.backward()` (reference modules/*.py training_step -> Lightning's backward).
    That one case is recognisable without a device read: `.backward()` is called on the very tensor object returned
    here, with `gradient=None`. The object is a FusedLoss (class swapped in place: same tensor, same autograd node) whose
    `backward` raises a flag for the duration of the call; any other route (`(2 * loss).backward()`,
    `torch.autograd.backward`, GradScaler's scaled copy, an explicit gradient) never touches the flag and takes the
    scaling launch as before."""
    flag = _UnitFlag()
    loss = _FusedLossFn.apply(pred, launch, flag)
    if loss.requires_grad and type(loss) is torch.Tensor:
        loss.__class__ = FusedLoss
        loss._mde_flag = flag
    return loss


def _as_images(pred):
    """(n_img, h, w) of a depth tensor: trailing two dims are the image, the rest are images."""
    if pred.dim() >= 2:
        h, w = int(pred.shape[-2]), int(pred.shape[-1])
    else:
        h, w = 1, int(pred.numel())
    n_img = pred.numel() // max(h * w, 1)
    return n_img, h, w


class _FusesMetrics:
    """Mixin: `criterion.fuse_metrics(metric_computation)` makes the loss launch ALSO produce the pooled
    metric suite of `metric_computation` from the same read of pred/target (C ABI
    mde_masked_loss_metrics). The next `metric_computation.compute(pred, target)` on the same tensors -
    what MetricLogger.log_train does right after the criterion call (reference modules/bts.py:106-108) -
    is then served from that launch instead of reading the tensors again. Pass None to undo."""

    _fused_metrics = None
    _fused_book = False

    def fuse_metrics(self, metric_computation, book=False):
        """`book=True`: the launch also ADDS the values to the computer's running sums and counts the call (reference
        metrics.py:64-66) at criterion time, in the kernel's finaliser - the later compute() on the same tensors only
        reads. Use it when every criterion call is followed by that compute() (the reference's training steps);
        with the default the running sums are updated by compute() itself, as in the reference."""
        self._fused_metrics = metric_computation
        self._fused_book = bool(book) and metric_computation is not None
        return self


def masked_loss(kind, pred, target, mask=None, params=None, totals=False, metrics=None, book_metrics=False):
    """Functional entry: fused forward(+backward when pred.requires_grad) of one masked loss.
    `metrics`: an optional metrics.MetricComputation to feed from the same launch (`book_metrics`: see
    _FusesMetrics.fuse_metrics)."""
    lib = _lib.load()
    dev = _lib.require_cuda(pred, target, mask)
    if pred.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError("pred must be fp32/fp16/bf16")
    tgt = target.detach()
    if tgt.shape != pred.shape:
        tgt = tgt.expand_as(pred)
    tgt = tgt.to(torch.float32).contiguous()
    mk = None
    if mask is not None:
        mk = mask.detach()
        if mk.shape != pred.shape:
            mk = mk.expand_as(pred)
        mk = (mk != 0).to(torch.uint8).contiguous()
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    if params:
        for k, v in params.items():
            setattr(lp, k, v)
    tot = torch.zeros(_lib.LOSS_NTOTALS, dtype=torch.float64, device=dev) if totals else None
    fuse = metrics is not None and kind != _lib.LOSS_EIGEN and not getattr(metrics, "reference_math", False)

    def launch(p, need_grad):
        pc = _compute_copy(p)
        n_img, h, w = _as_images(pc)
        with torch.cuda.device(dev):
            ws = _lib.workspace(dev, n_img)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(pc) if need_grad else None
            if fuse:
                m64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
                m32 = torch.empty(2 * _lib.METRIC_NM, dtype=torch.float32, device=dev)
                lp.metrics_accum = _lib.ptr(metrics.accum_buffer(dev)).value if book_metrics else None
                lp.metrics_raw_accum = _lib.ptr(metrics._raw_vec).value if (book_metrics and getattr(metrics, "_raw_vec", None) is not None) else None
                _lib.check(lib.mde_masked_loss_metrics(kind, _lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(tgt),
                                                       _lib.ptr(mk), n_img, h, w, C.byref(lp), 1.0,
                                                       metrics.group_flags(), _lib.ptr(ws), _lib.ptr(loss),
                                                       _lib.ptr(tot), _lib.ptr(grad), _lib.ptr(m64), _lib.ptr(m32),
                                                       _lib.stream_ptr(dev)))
                metrics.offer(pred, target, m32[:_lib.METRIC_NM], m64, booked=book_metrics)
            else:
                _lib.check(lib.mde_masked_loss(kind, _lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(tgt), _lib.ptr(mk),
                                               n_img, h, w, C.byref(lp), 1.0, _lib.ptr(ws), _lib.ptr(loss),
                                               _lib.ptr(tot), _lib.ptr(grad), _lib.stream_ptr(dev)))
        if grad is not None and grad.shape != p.shape:
            grad = grad.view(p.shape)
        return loss, grad

    if pred.numel() == 0:
        return torch.full((), float("nan"), device=dev) + 0 * pred.sum()
    loss = _fused_apply(pred, launch)
    return (loss, tot) if totals else loss


class MaskedDepthLoss(nn.Module):
    """reference criteria.py:17-64 (Eigen scale-invariant + gradient term)."""

    def __init__(self):
        super(MaskedDepthLoss, self).__init__()

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        if pred.dim() == 4 and pred.shape[1] != 1:
            raise ValueError("MaskedDepthLoss expects single-channel depth maps [B,1,H,W] or [B,H,W]")
        self.loss = masked_loss(_lib.LOSS_EIGEN, pred, target)
        return self.loss


class MaskedMSELoss(nn.Module, _FusesMetrics):
    """reference criteria.py:67-77."""

    def __init__(self):
        super(MaskedMSELoss, self).__init__()

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        self.loss = masked_loss(_lib.LOSS_MSE, pred, target, metrics=self._fused_metrics, book_metrics=self._fused_book)
        return self.loss


class MaskedL1Loss(nn.Module, _FusesMetrics):
    """reference criteria.py:80-90."""

    def __init__(self):
        super(MaskedL1Loss, self).__init__()

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        self.loss = masked_loss(_lib.LOSS_L1, pred, target, metrics=self._fused_metrics, book_metrics=self._fused_book)
        return self.loss


class berHuLoss(nn.Module, _FusesMetrics):
    """reference criteria.py:111-133 (threshold = 0.2*max(pred-target) over ALL pixels, signed)."""

    def __init__(self):
        super(berHuLoss, self).__init__()

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"
        self.loss = masked_loss(_lib.LOSS_BERHU, pred, target, metrics=self._fused_metrics, book_metrics=self._fused_book)
        return self.loss


class LainaBerHuLoss(nn.Module, _FusesMetrics):
    """reference criteria.py:476-506 (log-space berHu, differentiable threshold)."""

    def __init__(self, size_average=True, use_logs=True, clamp_val=1e-9):
        super(LainaBerHuLoss, self).__init__()
        self.size_average = size_average
        self.use_log = use_logs
        self.clamp_val = clamp_val

    def forward(self, input, target, mask=None):
        return masked_loss(_lib.LOSS_LAINA_BERHU, input, target, mask=mask, metrics=self._fused_metrics, book_metrics=self._fused_book,
                           params={"size_average": int(bool(self.size_average)), "use_logs": int(bool(self.use_log)),
                                   "clamp_val": float(self.clamp_val)})


class silog_loss(nn.Module, _FusesMetrics):
    """reference criteria.py:724-732 (mask is depth_gt > 1e-2)."""

    def __init__(self, variance_focus):
        super(silog_loss, self).__init__()
        self.variance_focus = variance_focus

    def forward(self, depth_est, depth_gt):
        return masked_loss(_lib.LOSS_SILOG, depth_est, depth_gt, metrics=self._fused_metrics, book_metrics=self._fused_book,
                           params={"variance_focus": float(self.variance_focus)})


# ---- DORN ------------------------------------------------------------------------------------------------
class ordLoss(nn.Module):
    """reference criteria.py:734-787: ord_labels = P [N,K,H,W], target = SID label [N,1,H,W] (float)."""

    def __init__(self):
        super(ordLoss, self).__init__()
        self.loss = 0.0

    def forward(self, ord_labels, target):
        lib = _lib.load()
        dev = _lib.require_cuda(ord_labels, target)
        N, K, H, W = ord_labels.size()
        tgt = target.detach().to(torch.float32).expand(N, 1, H, W).contiguous()

        def launch(p, need_grad):
            pc = p.detach().to(torch.float32).contiguous()
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, 1)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                _lib.check(lib.mde_ord_loss(_lib.ptr(pc), _lib.ptr(tgt), N, K, H * W, 1.0, _lib.ptr(ws),
                                            _lib.ptr(loss), _lib.ptr(grad), _lib.stream_ptr(dev)))
            return loss, grad      # fp32 stash; rounded to p.dtype after the scale (_FusedLossFn.backward)

        self.loss = _fused_apply(ord_labels, launch)
        return self.loss


class OrdinalRegressionLoss(object):
    """reference criteria.py:789-836: prob = log-probabilities [N,2K,H,W] laid out [K '<=' | K '>']."""

    def __init__(self, ord_num, alpha, beta, discretization="SID"):
        self.ord_num = ord_num
        self.alpha = alpha
        self.beta = beta
        self.discretization = discretization

    def __call__(self, prob, gt):
        lib = _lib.load()
        dev = _lib.require_cuda(prob, gt)
        if prob.shape != gt.shape:
            # criteria.py:826-827 (exact identity when the spatial sizes already agree)
            prob = torch.nn.functional.interpolate(prob, size=gt.shape[-2:], mode="bilinear", align_corners=True)
        N, C2, H, W = prob.shape
        K = int(self.ord_num)
        assert C2 == 2 * K, "prob must have 2*ord_num channels"
        gtc = gt.detach().to(torch.float32).reshape(N, H * W).contiguous()
        alpha, beta = float(self.alpha), float(self.beta)
        disc = _lib.DISC_SID if self.discretization == "SID" else _lib.DISC_UD

        def launch(p, need_grad):
            pc = p.detach().to(torch.float32).contiguous()
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, 1)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                _lib.check(lib.mde_ordinal_regression_loss(_lib.ptr(pc), _lib.ptr(gtc), N, K, H * W, alpha, beta,
                                                           disc, 1.0, _lib.ptr(ws), _lib.ptr(loss), _lib.ptr(grad),
                                                           _lib.stream_ptr(dev)))
            return loss, grad      # fp32 stash; rounded to p.dtype after the scale (_FusedLossFn.backward)

        return _fused_apply(prob, launch)


# ---- VNL ---------------------------------------------------------------------------------------------------
class VNL_Loss(nn.Module):
    """reference criteria.py:866-1045 (virtual-normal loss).

    `select_index()` keeps the reference's contract (criteria.py:912-932: a dict p1_x..p3_y of int
    arrays, redrawn on every call) and is the hook for supplying fixed triplets: override it, or
    call `set_triplets(flat_index_tensor[3, n])`.
    """

    def __init__(self, focal_x, focal_y, input_size,
                 delta_cos=0.867, delta_diff_x=0.01,
                 delta_diff_y=0.01, delta_diff_z=0.01,
                 delta_z=0.0001, sample_ratio=0.15):
        super(VNL_Loss, self).__init__()
        self.fx = float(focal_x)
        self.fy = float(focal_y)
        self.input_size = input_size
        self.u0 = float(input_size[1] // 2)
        self.v0 = float(input_size[0] // 2)
        self.delta_cos = delta_cos
        self.delta_diff_x = delta_diff_x
        self.delta_diff_y = delta_diff_y
        self.delta_diff_z = delta_diff_z
        self.delta_z = delta_z
        self.sample_ratio = sample_ratio
        self._fixed = None
        self.last_stats = None

    def set_triplets(self, trip):
        """Fix the triplets: int64 [3, n] flat pixel indices (y*W + x); None restores random sampling."""
        self._fixed = None if trip is None else trip.detach().to(torch.int64).contiguous()

    def select_index(self):
        H, W = int(self.input_size[0]), int(self.input_size[1])
        num = W * H
        n = int(num * self.sample_ratio)
        out = {}
        for k in (1, 2, 3):
            p = np.random.choice(num, n, replace=True)
            np.random.shuffle(p)
            out["p%d_x" % k] = p % W
            out["p%d_y" % k] = (p // W).astype(np.int64)
        return out

    def _triplets(self, device):
        if self._fixed is not None:
            return self._fixed.to(device)
        W = int(self.input_size[1])
        d = self.select_index()
        flat = np.stack([np.asarray(d["p%d_y" % k]).astype(np.int64) * W + np.asarray(d["p%d_x" % k]).astype(np.int64)
                         for k in (1, 2, 3)])
        return torch.from_numpy(flat).to(device)

    def forward(self, gt_depth, pred_depth, select=True):
        lib = _lib.load()
        dev = _lib.require_cuda(gt_depth, pred_depth)
        B, _, H, W = gt_depth.shape
        assert (H, W) == (int(self.input_size[0]), int(self.input_size[1])), "input_size mismatch"
        gt = gt_depth.detach().to(torch.float32).contiguous()
        trip = self._triplets(dev)
        n_trip = int(trip.shape[1])
        fx, fy = self.fx, self.fy
        stats = torch.zeros(8, dtype=torch.float64, device=dev)

        def launch(p, need_grad):
            pc = _compute_copy(p)
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, B)
                scratch = torch.empty(int(lib.mde_vnl_scratch_bytes(B, n_trip, H, W)), dtype=torch.uint8, device=dev)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                _lib.check(lib.mde_vnl_loss(_lib.ptr(gt), _lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(trip), B, H, W,
                                            n_trip, fx, fy, int(bool(select)), 1.0, _lib.ptr(ws), _lib.ptr(scratch),
                                            _lib.ptr(loss), _lib.ptr(stats), _lib.ptr(grad), _lib.stream_ptr(dev)))
            return loss, grad

        loss = _fused_apply(pred_depth, launch)
        self.last_stats = stats
        return loss


def compute_scale_and_shift(prediction, target, mask=None):
    """reference criteria.py:154-176: per-image least-squares (scale, shift) aligning `prediction` to `target` over
    the mask (default `target > 0`); zeros where the 2x2 system is singular. prediction/target [B,H,W] (or any
    [..., H, W]: every leading dim is an image) -> two fp32 tensors of shape [B]. One pass over the inputs
    (C ABI mde_scale_and_shift). Evaluation-side function: the result is detached (the differentiable use inside
    MidasLoss, criteria.py:306-332, is a later row of the scope table)."""
    lib = _lib.load()
    dev = _lib.require_cuda(prediction, target, mask)
    assert prediction.shape == target.shape, "inconsistent dimensions"
    n_img, h, w = _as_images(prediction)
    p = prediction.detach()
    if p.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        p = p.float()
    p = p.contiguous()
    t = target.detach().to(torch.float32).contiguous()
    m = None
    if mask is not None:
        assert mask.shape == target.shape, "inconsistent dimensions"
        m = (mask.detach() != 0).to(torch.uint8).contiguous()
    lead = tuple(prediction.shape[:-2]) if prediction.dim() > 2 else (1,)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, n_img)
        scale = torch.empty(n_img, dtype=torch.float32, device=dev)
        shift = torch.empty(n_img, dtype=torch.float32, device=dev)
        _lib.check(lib.mde_scale_and_shift(_lib.ptr(p), _lib.dtype_code(p), _lib.ptr(t), _lib.ptr(m), n_img, h * w,
                                           _lib.ptr(ws), _lib.ptr(scale), _lib.ptr(shift), None, _lib.stream_ptr(dev)))
    return scale.view(lead), shift.view(lead)


def scale_shift(pred, target):
    """MidasModule.scale_shift (reference modules/midas.py:56-62): align pred to target per image and return both
    as [B,1,H,W] (what the MiDaS module feeds to the metric logger in validation/test)."""
    lib = _lib.load()
    if pred.ndim == 4:
        pred = pred.squeeze(1)
    if target.ndim == 4:
        target = target.squeeze(1)
    scale, shift = compute_scale_and_shift(pred, target)
    dev = pred.device
    p = pred.detach()
    if p.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        p = p.float()
    p = p.contiguous()
    n_img, h, w = _as_images(p)
    with torch.cuda.device(dev):
        out = torch.empty(p.shape, dtype=torch.float32, device=dev)
        _lib.check(lib.mde_apply_scale_shift(_lib.ptr(p), _lib.dtype_code(p), _lib.ptr(scale.contiguous()),
                                             _lib.ptr(shift.contiguous()), n_img, h * w, _lib.ptr(out), _lib.stream_ptr(dev)))
    return out.unsqueeze(1), target.unsqueeze(1)


class MidasLoss(nn.Module):
    """reference criteria.py:306-332: data term + alpha * multi-scale gradient matching.

    `loss` in {'mse', 'l1', 'trim', 'ssimse', 'ssil1', 'ssitrim'} with reduction='batch-based': 'mse' with alpha=0.5
    is the criterion of the registered method `my` (modules/my.py:39), 'ssil1' / 'ssimse' / 'l1' / 'mse' / 'trim' those
    of `midas` (modules/midas.py:30-31; its default 'ssitrim' routes to TrimmedProcrustesLoss below). 'trim' (criteria.py:208-217) trims nothing as the reference is written and equals 'l1'.
    The 'ssi' variants align the prediction per image first (compute_scale_and_shift) and differentiate through
    that 2x2 solve (C ABI mde_midas_ssi_backward). reduction='image-based' is not built (it raises inside the
    reference for 'mse')."""

    def __init__(self, alpha=0.5, scales=4, loss='ssimse', reduction='batch-based'):
        super().__init__()
        self.loss = loss
        if 'trim' in self.loss or 'l1' in self.loss:
            self._kind = 1
        elif 'mse' in self.loss:
            self._kind = 0
        else:
            raise ValueError()
        self._ssi = 'ssi' in self.loss
        if reduction != 'batch-based':
            raise NotImplementedError("MidasLoss(reduction=%r): only 'batch-based' (the reference default) is built" % (reduction,))
        self._alpha = float(alpha)
        self._scales = int(scales)

    def forward(self, prediction, target):
        lib = _lib.load()
        dev = _lib.require_cuda(prediction, target)
        if prediction.ndim == 4:
            prediction = prediction.squeeze(1)
        if target.ndim == 4:
            target = target.squeeze(1)
        assert prediction.shape == target.shape and prediction.ndim == 3, "prediction/target must be [B,H,W] or [B,1,H,W]"
        B, H, W = (int(v) for v in prediction.shape)
        t = target.detach().to(torch.float32).contiguous()
        kind, alpha, scales, ssi = self._kind, self._alpha, self._scales, self._ssi

        def launch(p, need_grad):
            pc = p.detach()
            if pc.dtype != torch.float32:
                pc = pc.float()                      # fp32 stash (see _compute_copy); the backward through the solve needs it anyway
            pc = pc.contiguous()
            sp = _lib.stream_ptr(dev)
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, B)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pc) if need_grad else None
                scale = shift = sums = None
                if ssi:                              # criteria.py:326-328
                    scale = torch.empty(B, dtype=torch.float32, device=dev)
                    shift = torch.empty(B, dtype=torch.float32, device=dev)
                    sums = torch.empty((B, 5), dtype=torch.float64, device=dev)
                    _lib.check(lib.mde_scale_and_shift(_lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(t), None, B, H * W,
                                                       _lib.ptr(ws), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(sums), sp))
                _lib.check(lib.mde_midas_loss(_lib.ptr(pc), _lib.dtype_code(pc), _lib.ptr(t), _lib.ptr(scale), _lib.ptr(shift),
                                              B, H, W, kind, alpha, scales, 1.0, _lib.ptr(ws), _lib.ptr(loss), _lib.ptr(grad), sp))
                if ssi and need_grad:
                    coef = torch.empty((B, 4), dtype=torch.float32, device=dev)
                    _lib.check(lib.mde_midas_ssi_backward(_lib.ptr(pc), _lib.ptr(t), _lib.ptr(scale), _lib.ptr(shift),
                                                          _lib.ptr(sums), B, H * W, _lib.ptr(ws), _lib.ptr(coef), _lib.ptr(grad), sp))
            return loss, grad      # fp32 stash; _FusedLossFn.backward scales it and rounds to the prediction dtype once

        return _fused_apply(prediction, launch)


def normalize_prediction_robust(target, mask=None):
    """reference criteria.py:135-152: per image (x - median(mask * x)) / clamp(mean_mask |x - median|, 1e-6) with the
    default mask `target > 0` (the only one TrimmedProcrustesLoss uses). Evaluation-side function, detached result.
    [B,H,W] fp32 -> [B,H,W] fp32 (C ABI mde_robust_normalize with the tensor as its own mask source)."""
    lib = _lib.load()
    dev = _lib.require_cuda(target, mask)
    n_img, h, w = _as_images(target)
    x = target.detach().to(torch.float32).contiguous()
    # the kernel normalises its first tensor over the pixels where its SECOND tensor is > 0 (and the second over
    # itself): the default mask is the tensor itself, an explicit 0/1 mask takes the second seat
    if mask is None:
        src = x
    else:
        assert mask.shape == target.shape, "inconsistent dimensions"
        src = (mask.detach() != 0).to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        st_a = torch.empty((n_img, 8), dtype=torch.float32, device=dev)
        st_b = torch.empty((n_img, 8), dtype=torch.float32, device=dev)
        out_a, out_b = torch.empty_like(x), torch.empty_like(x)
        scratch = torch.empty(int(lib.mde_robust_scratch_bytes(n_img)) // 8, dtype=torch.float64, device=dev)
        _lib.check(lib.mde_robust_normalize(_lib.ptr(x), _lib.ptr(src), n_img, h * w, _lib.ptr(scratch), _lib.ptr(st_a), _lib.ptr(st_b),
                                            _lib.ptr(out_a), _lib.ptr(out_b), _lib.stream_ptr(dev)))
    return (out_b if mask is None else out_a).view(target.shape)


class TrimmedProcrustesLoss(nn.Module):
    """reference criteria.py:335-363, the default criterion of the registered method `midas` (`--loss ssitrim`,
    modules/midas.py:36-37): both tensors are normalised per image by normalize_prediction_robust (:135-152, median
    and mean absolute deviation over `target > 0`), then TrimmedMAELoss (= l1 as the reference is written, :208-217)
    + alpha * GradientLoss (:283-303) on the normalised pair, batch-based. The backward runs through the median
    (gradient to the element that holds it) and the deviation in closed form (C ABI mde_robust_backward).
    `prediction_ssi` holds the normalised prediction of the last call, as in the reference (:360-363)."""

    def __init__(self, alpha=0.5, scales=4, reduction="batch-based"):
        super().__init__()
        if reduction != 'batch-based':
            raise NotImplementedError("TrimmedProcrustesLoss(reduction=%r): only 'batch-based' (the reference default) is built" % (reduction,))
        self._alpha = float(alpha)
        self._scales = int(scales)
        self._prediction_ssi = None

    @property
    def prediction_ssi(self):
        return self._prediction_ssi

    def forward(self, prediction, target):
        lib = _lib.load()
        dev = _lib.require_cuda(prediction, target)
        if prediction.ndim == 4:
            prediction = prediction.squeeze(1)
        if target.ndim == 4:
            target = target.squeeze(1)
        assert prediction.shape == target.shape and prediction.ndim == 3, "prediction/target must be [B,H,W] or [B,1,H,W]"
        B, H, W = (int(v) for v in prediction.shape)
        t = target.detach().to(torch.float32).contiguous()
        alpha, scales = self._alpha, self._scales

        def launch(p, need_grad):
            pc = p.detach().float().contiguous()
            sp = _lib.stream_ptr(dev)
            with torch.cuda.device(dev):
                ws = _lib.workspace(dev, B)
                loss = torch.empty((), dtype=torch.float32, device=dev)
                st_p = torch.empty((B, 8), dtype=torch.float32, device=dev)
                st_t = torch.empty((B, 8), dtype=torch.float32, device=dev)
                pn, tn = torch.empty_like(pc), torch.empty_like(t)
                grad = torch.empty_like(pc) if need_grad else None
                scratch = torch.empty(int(lib.mde_robust_scratch_bytes(B)) // 8, dtype=torch.float64, device=dev)
                _lib.check(lib.mde_robust_normalize(_lib.ptr(pc), _lib.ptr(t), B, H * W, _lib.ptr(scratch), _lib.ptr(st_p), _lib.ptr(st_t),
                                                    _lib.ptr(pn), _lib.ptr(tn), sp))
                _lib.check(lib.mde_midas_loss_masked(_lib.ptr(pn), _lib.dtype_code(pn), _lib.ptr(tn), _lib.ptr(t), B, H, W, 1,
                                                     alpha, scales, 1.0, _lib.ptr(ws), _lib.ptr(loss), _lib.ptr(grad), sp))
                if need_grad:
                    coef = torch.empty((B, 4), dtype=torch.float32, device=dev)
                    _lib.check(lib.mde_robust_backward(_lib.ptr(pn), _lib.ptr(t), _lib.ptr(st_p), B, H * W, _lib.ptr(ws),
                                                       _lib.ptr(coef), _lib.ptr(grad), sp))
            self._prediction_ssi = pn
            return loss, grad      # fp32 stash; _FusedLossFn.backward scales it and rounds to the prediction dtype once

        return _fused_apply(prediction, launch)


class ModelLoss(nn.Module):
    """reference criteria.py:1047-1062: WCEL(pred_logit, bins, gt) + diff_loss_weight * VNL(gt, pred_depth)."""

    def __init__(self, args):
        super(ModelLoss, self).__init__()
        from .wcel import WCEL_Loss
        self.args = args
        self.weight_cross_entropy_loss = WCEL_Loss(args)
        self.virtual_normal_loss = VNL_Loss(focal_x=args.focal_x, focal_y=args.focal_y, input_size=args.crop_size)

    def forward(self, pred_depth, pred_logit, depth_bins, depth_gt):
        loss_metric = self.weight_cross_entropy_loss(pred_logit, depth_bins, depth_gt)
        loss_normal = self.virtual_normal_loss(depth_gt, pred_depth)
        return loss_metric + self.args.diff_loss_weight * loss_normal
