"""GPU parity of the MiDaS alignment step (evaluation side) against the oracle and the
reference-made golden vectors: compute_scale_and_shift (criteria.py:154-176), scale_shift (modules/midas.py:56-62)."""
import numpy as np
import pytest
import torch

from oracle import midas as om
from oracle import metrics as ometrics
from tests.gpu_util import T, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Cr():
    from mono_depth_estimation_b200 import criteria
    return criteria


def test_golden(Cr, golden):
    g = golden("midas_small.npz")
    pred, target = T(g["pred"]).cuda(), T(g["target"]).cuda()
    s, t = Cr.compute_scale_and_shift(pred, target)
    assert s.shape == (5,) and s.dtype == torch.float32 and s.is_cuda
    close(s, g["scale64"], 1e-5); close(t, g["shift64"], 1e-5, 1e-6)
    assert float(s[3]) == 0.0 and float(t[3]) == 0.0 and float(s[4]) == 0.0 and float(t[4]) == 0.0   # singular -> zeros
    sm, tm = Cr.compute_scale_and_shift(pred, target, T(g["mask"]).cuda())
    close(sm, g["scale_mask64"], 1e-5); close(tm, g["shift_mask64"], 1e-5, 1e-6)
    # second call on the same workspace (self-cleaning accumulators), then a metrics call sharing that region
    s2, _ = Cr.compute_scale_and_shift(pred, target)
    assert torch.equal(s, s2) or bool(torch.allclose(s, s2, rtol=1e-6))
    from mono_depth_estimation_b200 import metrics as M
    res = M.fused_metrics(pred[:3].clamp_min(1e-3), target[:3])
    v64 = [float(v) for v in ometrics.compute(pred[:3].clamp_min(1e-3).cpu().double(), target[:3].cpu().double(), ["mae", "delta1"])]
    close(res["f64"][[3, 0]], v64, 1e-5)


@pytest.mark.parametrize("shape,dtype", [((4, 1, 384, 384), torch.float32), ((3, 33, 41), torch.float32),
                                         ((64, 1, 96, 128), torch.float32), ((2, 1, 384, 384), torch.float16)])
def test_scale_shift_vs_oracle(Cr, shape, dtype):
    g = torch.Generator().manual_seed(5 + shape[0])
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.2] = 0.0
    pred = ((1.0 / target.clamp_min(0.3)) * 0.7 + 0.2 + torch.randn(shape, generator=g) * 0.05).to(dtype)
    p3 = pred.squeeze(1) if pred.ndim == 4 else pred
    t3 = target.squeeze(1) if target.ndim == 4 else target
    s64, t64 = om.compute_scale_and_shift(p3.double(), t3.double())
    s, t = Cr.compute_scale_and_shift(p3.cuda(), t3.cuda())
    close(s, s64, 1e-5); close(t, t64, 1e-5, 1e-6)
    y_hat, y = Cr.scale_shift(pred.cuda(), target.cuda())
    ref_hat, ref_y = om.scale_shift(pred.double(), target.double())
    assert y_hat.shape == ref_hat.shape and y.shape == ref_y.shape and y_hat.dtype == torch.float32
    close(y_hat, ref_hat, 2e-5, 2e-5)
    assert torch.equal(y.cpu(), ref_y.float())
    # the aligned prediction then goes to the metric suite (modules/midas.py:77-80)
    from mono_depth_estimation_b200 import metrics as M
    names = ["delta1", "absrel", "rmse"]
    vals = M.MetricComputation(names, strict=False).compute(y_hat, y)
    v64 = [float(v) for v in ometrics.compute(ref_hat, ref_y, names)]
    close(torch.stack(vals), v64, 2e-4)   # the metrics see an fp32 alignment of an fp32 (or fp16) prediction


@pytest.mark.parametrize("name,kw", [("mse", dict(alpha=0.5, loss="mse")), ("l1", dict(alpha=0.5, loss="l1")),
                                     ("trim", dict(alpha=0.5, loss="trim")), ("mse_a0", dict(alpha=0.0, loss="mse")),
                                     ("mse_s2", dict(alpha=0.25, scales=2, loss="mse")),
                                     ("ssimse", dict(alpha=0.5, loss="ssimse")), ("ssil1", dict(alpha=0.5, loss="ssil1")),
                                     ("ssimse_a0", dict(alpha=0.0, loss="ssimse"))])
def test_midas_loss_golden(Cr, golden, name, kw):
    from tests.gpu_util import LOSS_RTOL, grad_close, run_loss
    g = golden("midas_small.npz")
    pred, target = T(g["ml_pred"]).cuda(), T(g["ml_target"]).cuda()
    if "ssi" in name:
        pred = 0.7 / pred.clamp_min(0.3) + 0.2
    loss, grad = run_loss(Cr.MidasLoss(**kw), pred, target)
    assert loss.dim() == 0 and grad.shape == pred.shape
    close(loss, g[f"ml_{name}_loss64"], LOSS_RTOL)
    grad_close(grad, g[f"ml_{name}_grad64"])
    with torch.no_grad():
        close(Cr.MidasLoss(**kw)(pred, target), g[f"ml_{name}_loss64"], LOSS_RTOL)


@pytest.mark.parametrize("shape", [(8, 1, 384, 384), (3, 1, 33, 41), (2, 1, 480, 640)])
def test_midas_loss_vs_oracle(Cr, shape):
    """The `my` method's criterion (modules/my.py:39) at its training size and at odd sizes; all-invalid images."""
    from tests.gpu_util import LOSS_RTOL, grad_close, run_loss
    g = torch.Generator().manual_seed(17 + shape[0])
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.2] = 0.0
    target[-1, :, : shape[2] // 3] = 0.0
    pred = target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3
    disp = 0.7 / pred.clamp_min(0.3) + 0.2           # disparity-like input for the aligned ('ssi') variants
    for kw in (dict(alpha=0.5, loss="mse"), dict(alpha=0.5, loss="l1"), dict(alpha=0.5, loss="ssimse"),
               dict(alpha=0.5, loss="ssitrim")):
        src = disp if "ssi" in kw["loss"] else pred
        p64 = src.double().requires_grad_(True)
        l64 = om.midas_loss(p64, target.double(), **kw)
        (g64,) = torch.autograd.grad(l64, p64)
        loss, grad = run_loss(Cr.MidasLoss(**kw), src.cuda(), target.cuda())
        close(loss, l64.detach(), 3e-5 if "ssi" in kw["loss"] else LOSS_RTOL, msg=str(kw))
        if "ssi" in kw["loss"]:   # the alignment is an fp32 product + sum of an fp32 prediction: 1e-5-level gradient noise
            close(grad, g64, 1e-3, 2e-5 * float(g64.abs().max()), msg=str(kw))
        else:
            grad_close(grad, g64, msg=str(kw))
    z = torch.zeros(2, 1, 16, 24).cuda()
    assert float(Cr.MidasLoss(alpha=0.5, loss="mse")(torch.ones_like(z), z)) == 0.0      # zero divisors give 0 (criteria.py:185-186)
    with pytest.raises(NotImplementedError):
        Cr.MidasLoss(reduction="image-based")


def _tp_compare(grad, g_ref, pred, target, msg=""):
    """Per-image gradient check of TrimmedProcrustesLoss. torch.median hands the median's gradient to ONE of the
    elements holding the median value; which one is an implementation detail (CPU sort order) when valid pixels
    tie, so on images whose valid predictions are not distinct the comparison is on the gradient with the
    element(s) that differ by the median term removed, plus the image sum."""
    g_ref = g_ref.detach().double().cpu() if torch.is_tensor(g_ref) else torch.from_numpy(np.asarray(g_ref, dtype=np.float64))
    g = grad.detach().double().cpu()
    for b in range(g.shape[0]):
        gb, rb = g[b].reshape(-1), g_ref[b].reshape(-1)
        scale = float(rb.abs().max())
        if scale == 0.0:
            assert float(gb.abs().max()) == 0.0, msg
            continue
        pv = pred[b].reshape(-1)[target[b].reshape(-1) > 0]
        if pv.unique().numel() == pv.numel():
            close(gb, rb, 2e-4, 1e-5 * scale, msg=msg)
        else:
            bad = ((gb - rb).abs() > 1e-4 * scale).nonzero().reshape(-1)
            assert bad.numel() in (0, 2), msg
            close(gb.sum(), rb.sum(), 1e-3, 1e-5 * scale, msg=msg)


@pytest.mark.parametrize("name,kw", [("tp", dict(alpha=0.5)), ("tp_a0", dict(alpha=0.0)), ("tp_s2", dict(alpha=0.25, scales=2))])
def test_trimmed_procrustes_golden(Cr, golden, name, kw):
    from tests.gpu_util import run_loss
    g = golden("midas_small.npz")
    pred, target = T(g["tp_pred"]), T(g["tp_target"])
    mod = Cr.TrimmedProcrustesLoss(**kw)
    loss, grad = run_loss(mod, pred.cuda(), target.cuda())
    assert loss.dim() == 0 and grad.shape == pred.shape
    close(loss, g[f"{name}_loss32"], 1e-5)
    _tp_compare(grad, g[f"{name}_grad32"], pred, target, msg=name)
    if name == "tp":
        close(mod.prediction_ssi, g["tp_ssi32"], 1e-5, 1e-6)
        close(Cr.normalize_prediction_robust(target.squeeze(1).cuda()), g["tp_tnorm32"], 1e-5, 1e-6)
    with torch.no_grad():
        close(Cr.TrimmedProcrustesLoss(**kw)(pred.cuda(), target.cuda()), g[f"{name}_loss32"], 1e-5)


@pytest.mark.parametrize("shape", [(8, 1, 384, 384), (3, 1, 33, 41), (2, 1, 480, 640)])
def test_trimmed_procrustes_vs_oracle(Cr, shape):
    """`midas --loss ssitrim` (modules/midas.py:36-37) at its training size (384 x 384) and at odd sizes."""
    from tests.gpu_util import run_loss
    g = torch.Generator().manual_seed(23 + shape[0])
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.2] = 0.0
    target[-1, :, : shape[2] // 3] = 0.0
    pred = 0.7 / (target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3).clamp_min(0.3) + 0.2
    pred = pred + torch.rand(shape, generator=g) * 1e-3        # distinct values: the median element is unique
    for kw in (dict(alpha=0.5), dict(alpha=0.0)):
        p64 = pred.double().requires_grad_(True)
        l64, ssi64 = om.trimmed_procrustes_loss(p64, target.double(), **kw)
        (g64,) = torch.autograd.grad(l64, p64)
        mod = Cr.TrimmedProcrustesLoss(**kw)
        loss, grad = run_loss(mod, pred.cuda(), target.cuda())
        close(loss, l64.detach(), 2e-5, msg=str(kw))
        close(mod.prediction_ssi, ssi64.detach(), 2e-5, 2e-6, msg=str(kw))
        _tp_compare(grad, g64, pred, target, msg=str(kw))
    # the median itself is exact: (x - m) vanishes at the element that holds it
    ssi = Cr.TrimmedProcrustesLoss()
    ssi(pred.cuda(), target.cuda())
    masked = (pred * (target > 0)).reshape(shape[0], -1)
    med = masked.median(dim=1).values
    s = ((pred - med.view(-1, 1, 1, 1)).abs() * (target > 0)).reshape(shape[0], -1).sum(1) / (target > 0).reshape(shape[0], -1).sum(1)
    close(ssi.prediction_ssi.cpu().reshape(shape), (pred - med.view(-1, 1, 1, 1)) / s.view(-1, 1, 1, 1), 1e-5, 1e-6)
    with pytest.raises(NotImplementedError):
        Cr.TrimmedProcrustesLoss(reduction="image-based")


def test_trimmed_procrustes_many_small_images(Cr):
    """More (image, tensor) pairs than CTAs of the cooperative statistics launch (2 x 80 > 148), tiny images."""
    from tests.gpu_util import run_loss
    shape = (80, 1, 24, 32)
    g = torch.Generator().manual_seed(99)
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.2] = 0.0
    pred = 0.7 / (target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3).clamp_min(0.3) + 0.2
    pred = pred + torch.rand(shape, generator=g) * 1e-3
    p64 = pred.double().requires_grad_(True)
    l64, ssi64 = om.trimmed_procrustes_loss(p64, target.double(), alpha=0.5)
    (g64,) = torch.autograd.grad(l64, p64)
    mod = Cr.TrimmedProcrustesLoss(alpha=0.5)
    loss, grad = run_loss(mod, pred.cuda(), target.cuda())
    close(loss, l64.detach(), 2e-5)
    close(mod.prediction_ssi, ssi64.detach(), 2e-5, 2e-6)
    _tp_compare(grad, g64, pred, target)


@pytest.mark.parametrize("width", [90, 88])     # 88: rows of whole quads, the 128-bit path
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_midas_loss_half_prediction(Cr, dtype, width):
    """AMP: a half-precision prediction is read as is, the gradient comes back in the same dtype (the coarse scales
    add their share to the stored gradient in a second pass)."""
    from tests.gpu_util import run_loss
    shape = (3, 1, 66, width)
    g = torch.Generator().manual_seed(41)
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.2] = 0.0
    pred = (target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3).to(dtype)
    p64 = pred.double().requires_grad_(True)
    l64 = om.midas_loss(p64, target.double(), alpha=0.5, loss="l1")
    (g64,) = torch.autograd.grad(l64, p64)
    loss, grad = run_loss(Cr.MidasLoss(alpha=0.5, loss="l1"), pred.cuda(), target.cuda())
    assert grad.dtype == dtype
    close(loss, l64.detach(), 1e-5)
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    close(grad, g64, tol, tol * float(g64.abs().max()))


@pytest.mark.parametrize("shape", [(2, 1, 37, 64), (5, 1, 16, 4), (3, 1, 1, 128), (2, 1, 9, 132), (2, 1, 40, 43)])
def test_midas_loss_quad_path_edges(Cr, shape):
    """Widths that are multiples of 4 (scale 0 runs on quads): narrow rows, a single row, a width that is not a
    multiple of the warp's 128 pixels; 'l1' and the aligned 'ssimse'."""
    from tests.gpu_util import LOSS_RTOL, grad_close, run_loss
    g = torch.Generator().manual_seed(7 + shape[2])
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.25] = 0.0
    pred = target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3
    for kw in (dict(alpha=0.5, loss="l1"), dict(alpha=0.7, loss="mse", scales=3), dict(alpha=0.3, loss="l1", scales=6),
               dict(alpha=0.5, loss="mse", scales=1)):
        p64 = pred.double().requires_grad_(True)
        l64 = om.midas_loss(p64, target.double(), **kw)
        (g64,) = torch.autograd.grad(l64, p64)
        loss, grad = run_loss(Cr.MidasLoss(**kw), pred.cuda(), target.cuda())
        close(loss, l64.detach(), LOSS_RTOL, msg=str(kw))
        grad_close(grad, g64, msg=str(kw))


def test_normalize_prediction_robust_explicit_mask(Cr):
    """normalize_prediction_robust(target, mask) with an explicit 0/1 mask (reference criteria.py:135-152): per image
    (x - median(mask * x)) / clamp(sum mask |x - median| / sum mask, 1e-6), the masked zeros taking part in the median."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand((3, 40, 56), generator=g) * 8 + 0.3
    mask = (torch.rand((3, 40, 56), generator=g) < 0.6).float()
    mask[2] = 0                                               # an image without a valid pixel: m = 0, s = 1
    if True:
        ssum = mask.sum((1, 2)); valid = ssum > 0
        m = torch.zeros_like(ssum); s = torch.ones_like(ssum)
        m[valid] = torch.median((mask[valid] * x[valid]).view(int(valid.sum()), -1), dim=1).values
        t = x - m.view(-1, 1, 1)
        sq = (mask * t.abs()).sum((1, 2))
        s[valid] = torch.clamp(sq[valid] / ssum[valid], min=1e-6)
        ref = t / s.view(-1, 1, 1)
    out = Cr.normalize_prediction_robust(x.cuda(), mask.cuda())
    close(out, ref, 1e-5, 1e-6)
