"""GPU parity of VNL's classification half: WCEL_Loss, depth_to_bins, bins_to_depth and
ModelLoss through the C ABI, against the oracle (oracle/wcel.py) and the reference-made golden vectors."""
import types

import numpy as np
import pytest
import torch

from oracle import wcel as ow
from oracle import vnl as ovnl
from tests.gpu_util import LOSS_RTOL, T, close, grad_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Wc():
    from mono_depth_estimation_b200 import wcel
    return wcel


def _args(C, H=12, W=20):
    p = ow.vnl_params(0.01, 1.1, C)
    return p, types.SimpleNamespace(wce_loss_weight=p["wce_loss_weight"], dec_out_c=C, focal_x=519.0, focal_y=519.0,
                                    crop_size=(H, W), diff_loss_weight=6.0)


def _check_bins(bins_gpu, depth_before, p):
    """Integer bins must equal the reference's except where the fp64 quotient sits within 1e-4 of a bin border
    (CPU and GPU log10f differ in the last ulp there); padding / clamping rules are exact."""
    d = depth_before.double().cpu()
    invalid = d < 0
    dc = d.clamp(p["depth_min"], p["depth_max"])
    q = (torch.log10(dc) - p["depth_min_log"]) / p["depth_bin_interval"]
    ref = q.to(torch.int64)
    ref[ref == p["dec_out_c"]] = p["dec_out_c"] - 1
    ref[invalid] = p["dec_out_c"] + 1
    got = bins_gpu.cpu().to(torch.int64)
    diff = got != ref
    near = (q - q.round()).abs() < 1e-4
    assert not bool((diff & ~near).any()), "bin mismatch away from a border"
    assert int((got - ref).abs().max()) <= 1
    return int(diff.sum())


@pytest.mark.parametrize("tag,C", [("c150", 150), ("c24", 24)])
def test_golden(Wc, golden, tag, C):
    g = golden("wcel_small.npz")
    p, args = _args(C)
    gt = T(g[f"{tag}_gt"]).cuda()
    before = gt.clone()
    bins = Wc.depth_to_bins(gt, p["depth_min"], p["depth_max"], C)
    assert bins.dtype == torch.int32 and bins.shape == gt.shape
    _check_bins(bins, before, p)
    assert np.array_equal(bins.cpu().numpy(), g[f"{tag}_bins"])            # this fixture has no borderline pixel
    assert np.array_equal(gt.cpu().numpy(), g[f"{tag}_gt_after"])          # in-place clamp, padding back to -1
    crit = Wc.WCEL_Loss(args)
    lg = T(g[f"{tag}_logits"]).cuda().requires_grad_(True)
    loss = crit(lg, bins, gt)
    assert loss.dim() == 0 and loss.is_cuda
    loss.backward()
    close(loss, g[f"{tag}_loss64"], LOSS_RTOL)
    grad_close(lg.grad, g[f"{tag}_grad64"])
    with torch.no_grad():
        close(Wc.WCEL_Loss(args)(lg.detach(), bins, gt), g[f"{tag}_loss64"], LOSS_RTOL)
    # bins -> depth and its backward
    sm = T(g[f"{tag}_softmax"]).cuda().requires_grad_(True)
    d = Wc.bins_to_depth(sm, p["depth_bin_border"])
    assert d.shape == (sm.shape[0], 1) + tuple(sm.shape[2:])
    close(d, g[f"{tag}_depth64"], 1e-5)
    d.sum().backward()
    close(sm.grad, g[f"{tag}_depth_gradsum32"], 1e-5, 1e-9)


def test_model_loss_golden(golden):
    from mono_depth_estimation_b200 import criteria, wcel
    g = golden("wcel_small.npz")
    C = 24
    p, args = _args(C)
    gt = T(g["ml_gt"]).cuda()
    bins = wcel.VNLBins(0.01, 1.1, C).depth_to_bins(gt)
    assert np.array_equal(bins.cpu().numpy(), g["ml_bins"]) and np.array_equal(gt.cpu().numpy(), g["ml_gt_after"])
    ml = criteria.ModelLoss(args)
    ml.virtual_normal_loss.set_triplets(T(g["ml_trip"]).cuda())
    lg = T(g["ml_logits"]).cuda().requires_grad_(True)
    pd = T(g["ml_pred"]).cuda().requires_grad_(True)
    total = ml(pd, lg, bins, gt)
    total.backward()
    close(total, g["ml_total32"], 2e-5)
    grad_close(lg.grad, g["ml_grad_logits32"])
    close(pd.grad, g["ml_grad_pred32"], 1e-3, 2e-5 * float(np.abs(g["ml_grad_pred32"]).max()))


@pytest.mark.parametrize("C,dtype", [(150, torch.float32), (37, torch.float32), (240, torch.float32), (150, torch.float16),
                                     (64, torch.bfloat16)])
def test_vs_oracle_shapes_and_dtypes(Wc, C, dtype):
    """Odd plane sizes (hw not a multiple of 4), channel counts off the unroll, the global-memory weight table
    (C = 240 does not fit shared memory) and half-precision logits."""
    p, args = _args(C)
    g = torch.Generator().manual_seed(300 + C)
    B, H, W = 3, 17, 23
    gt = torch.rand((B, 1, H, W), generator=g) * 1.2 + 0.005
    gt[2, :, :4, :] = -1.0
    gt[0, 0, 1, 1] = 0.0
    logits = (torch.randn((B, C, H, W), generator=g) * 4.0).to(dtype)
    gtc = gt.clone()
    bins_ref = ow.depth_to_bins(gtc, p)
    l64 = ow.wcel_loss(logits.double().requires_grad_(True), bins_ref, gtc, p["wce_loss_weight"], C)
    lg64 = logits.double().requires_grad_(True)
    l64 = ow.wcel_loss(lg64, bins_ref, gtc, p["wce_loss_weight"], C)
    (g64,) = torch.autograd.grad(l64, lg64)
    x = logits.cuda().requires_grad_(True)
    loss = Wc.WCEL_Loss(args)(x, bins_ref.cuda(), gtc.cuda())
    loss.backward()
    assert x.grad.dtype == dtype
    close(loss, l64.detach(), LOSS_RTOL)
    if dtype == torch.float32:
        grad_close(x.grad, g64)
    else:
        close(x.grad.float(), g64, 1e-2, 1e-2 * float(g64.abs().max()))
    assert float(x.grad[2, :, :4, :].abs().max()) == 0.0                   # padding rows: exactly zero


def test_all_padding_is_nan(Wc):
    p, args = _args(24)
    gt = torch.full((1, 1, 8, 8), -1.0).cuda()
    bins = torch.full((1, 1, 8, 8), 25, dtype=torch.int32).cuda()
    loss = Wc.WCEL_Loss(args)(torch.randn(1, 24, 8, 8).cuda(), bins, gt)
    assert torch.isnan(loss)                                                # -0 / 0 as in the reference


def test_depth_to_bins_at_scale(Wc):
    """C4-shaped gt: exact away from borders, few borderline pixels, in-place rules."""
    p, _ = _args(150)
    g = torch.Generator().manual_seed(41)
    gt = torch.rand((8, 1, 385, 385), generator=g) * 1.25 + 0.001
    gt[1::2, :, :40, :] = -1.0
    d = gt.clone().cuda()
    bins = Wc.depth_to_bins(d, 0.01, 1.1, 150)
    n_border = _check_bins(bins, gt, p)
    assert n_border < 50
    dc = d.cpu()
    assert bool((dc[gt < 0] == -1.0).all()) and float(dc[gt >= 0].min()) >= np.float32(0.01) and float(dc.max()) <= np.float32(1.1)
    assert int(bins[d < 0].min()) == 151 and int(bins[d >= 0].max()) <= 149 and int(bins.min()) >= 0


def test_config_c4_sized_properties_and_slice_parity(Wc):
    """Full VNL-config size (8 x 150 x 385 x 385 logits, 711 MB): size-independent properties - every valid pixel's
    gradient sums to ~0 over the channels (softmax and the normalised weight row both sum to 1), padding is exactly
    zero, the loss is additive over the batch - and parity with the oracle on one image."""
    C = 150
    p, args = _args(C, 385, 385)
    g = torch.Generator(device="cuda").manual_seed(42)
    B = 8
    gt = torch.rand((B, 1, 385, 385), generator=g, device="cuda") * 1.2 + 0.005
    gt[1::2, :, :40, :] = -1.0
    logits = torch.randn((B, C, 385, 385), generator=g, device="cuda") * 3.0
    bins = Wc.depth_to_bins(gt, 0.01, 1.1, C)
    crit = Wc.WCEL_Loss(args)
    x = logits.requires_grad_(True)
    loss = crit(x, bins, gt)
    loss.backward()
    gsum = x.grad.sum(1)
    assert float(gsum.abs().max()) < 2e-5 * float(x.grad.abs().max())
    assert float(x.grad[1::2, :, :40, :].abs().max()) == 0.0
    n_all = float((gt > 0).sum())
    parts = 0.0
    for sl in (slice(0, 3), slice(3, 8)):
        with torch.no_grad():
            parts += float(crit(logits[sl].detach(), bins[sl], gt[sl])) * float((gt[sl] > 0).sum())
    close(float(loss), parts / n_all, 1e-5)
    b = 1
    lg64 = logits[b:b + 1].detach().cpu().double().requires_grad_(True)
    l64 = ow.wcel_loss(lg64, bins[b:b + 1].cpu(), gt[b:b + 1].cpu(), p["wce_loss_weight"], C)
    (g64,) = torch.autograd.grad(l64, lg64)
    x1 = logits[b:b + 1].detach().clone().requires_grad_(True)
    l1 = crit(x1, bins[b:b + 1], gt[b:b + 1])
    l1.backward()
    close(l1, l64.detach(), LOSS_RTOL)
    grad_close(x1.grad, g64)
