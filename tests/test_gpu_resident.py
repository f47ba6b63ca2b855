"""GPU parity of the register-resident small-input loss kernel (csrc/resident_loss.cuh; C ABI mde_masked_loss /
mde_masked_loss_metrics for fp32 inputs of at most 2 quads per thread of one CTA per SM) against the pinned CPU oracle
(reference criteria.py:67-133, :476-506, :724-732; metrics.py:58-67), and against the generic persistent kernel, which
the same entry points take when an explicit mask is passed."""
import ctypes as C

import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import losses as olosses, metrics as ometrics
from tests.gpu_util import LOSS_RTOL, close, grad_close

pytestmark = pytest.mark.gpu
KINDS = ["l1", "mse", "berhu", "laina_berhu", "silog"]
NAMES7 = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
NAMES10 = ["delta1", "delta2", "delta3", "mae", "mse", "msle", "log10", "absrel", "sqrel", "rmse"]


@pytest.fixture(scope="module")
def Cr():
    from mono_depth_estimation_b200 import criteria
    return criteria


def kind_code(name):
    from mono_depth_estimation_b200 import _lib
    return {"l1": _lib.LOSS_L1, "mse": _lib.LOSS_MSE, "berhu": _lib.LOSS_BERHU, "laina_berhu": _lib.LOSS_LAINA_BERHU,
            "silog": _lib.LOSS_SILOG}[name]


def capacity_px():
    """pixels the resident kernel holds with ONE quad per thread (one 1024-thread CTA per SM)"""
    from mono_depth_estimation_b200 import _lib
    sm, coop = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().mde_device_info(C.byref(sm), C.byref(coop)))
    return sm.value * 1024 * 4


def run(Cr, name, pred, gt, names=None, mask=None):
    from mono_depth_estimation_b200 import metrics as M, _lib
    p = pred.cuda().requires_grad_(True)
    g = gt.cuda()                                     # kept alive: the metric hand-over holds weak references
    mc = M.MetricComputation(names, strict=False) if names else None
    _lib.workspace(p.device, p.shape[0])              # first use initialises the workspace (one tiny launch)
    n0 = _lib.launch_count()
    loss = Cr.masked_loss(kind_code(name), p, g, mask=mask, metrics=mc)
    vals = torch.stack(mc.compute(p.detach(), g)) if names else None
    assert _lib.launch_count() - n0 == 1, "loss, gradient and metrics must come from ONE launch"
    loss.backward()
    return loss.detach(), p.grad.detach(), vals


# C1 itself, a ragged shape with an n % 4 tail, a one-CTA input, a two-quads-per-thread input; noise 3.0 puts
# pixels on both sides of the berHu / Laina threshold c = 0.2 max(...)
SHAPES = [((8, 1, 228, 304), 0.5), ((8, 1, 228, 304), 3.0), ((2, 1, 47, 63), 3.0), ((1, 1, 9, 7), 0.5), ((8, 1, 300, 400), 3.0)]


@pytest.mark.parametrize("name", KINDS)
@pytest.mark.parametrize("shape,noise", SHAPES)
def test_resident_vs_oracle(Cr, name, shape, noise):
    pred, gt = synth.depth_pair(shape, 31 + shape[2], border=2, noise=noise)
    l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], pred.double(), gt.double())
    loss, grad, _ = run(Cr, name, pred, gt)
    close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
    # with the metric suite in the same launch: loss and gradient unchanged, metric values within 1e-5,
    # the three threshold counts exact (integers: compared with the fp32 reference arithmetic)
    for names in (NAMES7, NAMES10):
        lf, gf, vals = run(Cr, name, pred, gt, names)
        close(lf, l64, LOSS_RTOL); grad_close(gf, g64)
        v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), names)]
        close(vals, v64, 1e-5)
        n, c1, c2, c3 = ometrics.delta_counts(pred, gt)
        got = (vals[:3].double().cpu() * n).round().long().tolist()
        assert got == [c1, c2, c3], (name, names, got, (c1, c2, c3))


@pytest.mark.parametrize("name", ["l1", "mse", "berhu", "silog"])
def test_resident_agrees_with_the_generic_kernel(Cr, name):
    """An explicit mask sends the call to the generic persistent kernel (these four kinds ignore the mask's content:
    their mask is target > 0 / > 0.01 as in the reference). Different summation orders (and, for SILog, residuals in
    log2 units in the generic fused kernel): 2e-6 on the values, the gradient tolerance on the gradients - not bit-equal."""
    pred, gt = synth.depth_pair((8, 1, 228, 304), 41, border=2, noise=3.0)
    la, ga, va = run(Cr, name, pred, gt, NAMES7)
    lb, gb, vb = run(Cr, name, pred, gt, NAMES7, mask=(gt > 0).cuda())
    close(la, lb, 2e-6); grad_close(ga, gb); close(va, vb, 2e-6)
    assert torch.equal((va[:3].double() * 1e6).round(), (vb[:3].double() * 1e6).round())


def test_resident_capacity_boundary(Cr):
    """Exactly 2 quads per thread still runs in registers; one tile more takes the generic kernel (or, for SILog, the
    shared-memory variant). Both sides of the boundary against the oracle."""
    cap = capacity_px()
    for rows in (2 * cap // 4096, 2 * cap // 4096 + 1):
        pred, gt = synth.depth_pair((1, 1, rows, 4096), 51, border=0, noise=3.0)
        for name in ("berhu", "silog"):
            l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], pred.double(), gt.double())
            loss, grad, _ = run(Cr, name, pred, gt, NAMES7)
            close(loss, l64, LOSS_RTOL); grad_close(grad, g64)


def test_resident_special_values(Cr):
    """NaN in the unmasked maximum poisons berHu as torch.max does (criteria.py:118-119); an all-negative p - t gives
    a negative threshold; no valid pixel gives NaN (0 / 0) exactly like the reference's mean of an empty tensor."""
    pred, gt = synth.depth_pair((2, 1, 40, 64), 61, border=2)
    p2 = pred.clone(); p2[1, 0, 3, 5] = float("nan")
    loss, _, _ = run(Cr, "berhu", p2, gt)
    assert torch.isnan(loss)
    p3 = (gt - 0.25).clamp_min(1e-3)
    g3 = gt.clone(); g3[gt <= 0] = 20.0          # every pixel valid, p - t < 0 everywhere
    for name in ("berhu", "laina_berhu"):
        l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], p3.double(), g3.double())
        loss, grad, _ = run(Cr, name, p3, g3)
        close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
    for name in ("l1", "mse", "berhu"):
        loss, grad, _ = run(Cr, name, pred, torch.zeros_like(gt))
        assert torch.isnan(loss), name


def test_resident_is_bit_reproducible(Cr):
    """Races (slot words recycled between the max and the sum exchange, workspace parity sets) are hunted by determinism:
    the totals are summed in a fixed order, so repeated launches - interleaved with other kinds and sizes that reuse the
    same workspace - give bit-identical loss, gradient and metric values."""
    pred, gt = synth.depth_pair((8, 1, 228, 304), 71, border=2, noise=3.0)
    other_p, other_g = synth.depth_pair((16, 1, 480, 640), 72)
    for name in KINDS:
        ref = None
        for it in range(12):
            out = run(Cr, name, pred, gt, NAMES7)
            if ref is None:
                ref = out
            else:
                assert all(torch.equal(a, b) for a, b in zip(out, ref)), (name, it)
            if it % 4 == 0:
                run(Cr, "silog", other_p, other_g, NAMES7)
                run(Cr, "berhu", pred[:1], gt[:1])


# ---- MaskedDepthLoss (criteria.py:17-64): register-resident variant of csrc/eigen.cu ------------------------------------
# C1 itself (16.9 CTAs per image: most CTAs hold one image, every 17th two), images of exactly one CTA, images of 4420
# pixels (every CTA straddles two images), more images than warps (the per-image gather loops), one image, one CTA
EIGEN_SHAPES = [((8, 1, 228, 304), 0.5), ((8, 1, 228, 304), 3.0), ((5, 1, 120, 160), 3.0), ((3, 1, 64, 64), 0.5),
                ((2, 1, 65, 68), 3.0), ((37, 1, 64, 64), 0.5), ((1, 1, 64, 64), 3.0), ((1, 1, 300, 400), 0.5)]


def run_eigen(Cr, pred, gt, expect_resident=True):
    from mono_depth_estimation_b200 import _lib
    p = pred.requires_grad_(True) if pred.is_cuda else pred.cuda().requires_grad_(True)
    g = gt.cuda()
    _lib.workspace(p.device, p.shape[0])
    n0 = _lib.launch_count()
    loss = Cr.MaskedDepthLoss()(p, g)
    assert _lib.launch_count() - n0 == 1, "loss and gradient must come from ONE launch"
    loss.backward()
    return loss.detach(), p.grad.detach()


@pytest.mark.parametrize("shape,noise", EIGEN_SHAPES)
def test_resident_eigen_vs_oracle_and_generic(Cr, shape, noise):
    pred, gt = synth.depth_pair(shape, 91 + shape[0], border=2, noise=noise)
    l64, g64 = olosses.loss_and_grad(olosses.LOSSES["eigen"], pred.double(), gt.double())
    loss, grad = run_eigen(Cr, pred, gt)
    close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
    with torch.no_grad():                              # forward only: same value, no gradient work
        close(Cr.MaskedDepthLoss()(pred.cuda(), gt.cuda()), l64, LOSS_RTOL)
    # a prediction 4 bytes off a 16-byte boundary takes the cooperative kernel of the same file: same per-pixel
    # expressions, different summation order of the fp64 totals
    buf = torch.zeros(pred.numel() + 1).cuda()
    buf[1:] = pred.flatten().cuda()
    lg, gg = run_eigen(Cr, buf[1:].view(pred.shape).detach(), gt)
    close(loss, lg, 2e-6); close(grad, gg, 2e-6, 1e-7 * float(gg.abs().max()))


def test_resident_eigen_sparse_targets_and_reproducibility(Cr):
    """Images without a valid pixel (n_b = 0 contributes nothing, criteria.py:38-41), an image whose valid pixels have no
    valid neighbour, and bit-identical results over repeated launches interleaved with other kernels that recycle the
    same workspace words (slots, mslots rows, parity sets)."""
    pred, gt = synth.depth_pair((8, 1, 228, 304), 97, border=2, noise=3.0)
    gt[2] = 0.0
    gt[5, :, 1::2, :] = 0.0                            # no vertical pairs in image 5
    l64, g64 = olosses.loss_and_grad(olosses.LOSSES["eigen"], pred.double(), gt.double())
    ref = None
    for it in range(10):
        loss, grad = run_eigen(Cr, pred, gt)
        if ref is None:
            ref = (loss, grad)
            close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
        else:
            assert torch.equal(loss, ref[0]) and torch.equal(grad, ref[1]), it
        if it % 3 == 0:
            run(Cr, "berhu", pred, gt, NAMES7)
            run_eigen(Cr, pred[:3, :, :64, :64].contiguous(), gt[:3, :, :64, :64].contiguous())
    loss, _ = run_eigen(Cr, pred, torch.zeros_like(gt))
    assert torch.isnan(loss)                           # 0 / 0 like the reference
