"""GPU parity against the CPU oracle AT THE SIZES BASELINE.json names (configs C2-C5), not only on small shapes.

The oracle restates the reference's ATen chains, so these cases cost it seconds (C2, C4) to tens of seconds (C3, C5) on the
box's host cores. Tolerances are north_star's: decoded bins, valid / threshold counts bit-exact; losses, gradients and float
metrics within 1e-5 relative."""
import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import _lib, synth
from oracle import dorn as odorn
from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import vnl as ovnl
from tests.gpu_util import LOSS_RTOL, close, grad_close

pytestmark = pytest.mark.gpu
TRAIN = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
EVAL = TRAIN + ["absrel", "sqrel", "msle"]
ALL = ["delta1", "delta2", "delta3", "mae", "mse", "log10", "msle", "absrel", "sqrel", "rmse", "rmse_true", "rmse_log"]


@pytest.mark.parametrize("names", [TRAIN, EVAL])
def test_c2_fused_silog_and_metrics_vs_oracle(names):
    """C2, the kernel bench.py times (SILog forward+backward + the metric suite in ONE launch, shared-memory residual
    stash): loss, the FULL gradient, every metric value and the exact valid / delta counts against the oracle."""
    from mono_depth_estimation_b200 import criteria as Cr, metrics as M
    pred, gt = synth.config_inputs("C2")
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double(), 0.85)
    v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), names)]
    mc = M.MetricComputation(names, strict=False)
    crit = Cr.silog_loss(0.85).fuse_metrics(mc)
    p, g = pred.cuda().requires_grad_(True), gt.cuda()
    _lib.workspace(p.device, pred.shape[0])
    n0 = _lib.launch_count()
    loss = crit(p, g)
    vals = mc.compute(p.detach(), g)
    assert _lib.launch_count() - n0 == 1, "loss and metrics must come from one launch"
    loss.backward()
    close(loss, l64, LOSS_RTOL)
    grad_close(p.grad, g64)
    close(torch.stack(vals), v64, 1e-5)
    raw = mc.last_f64[2 * _lib.METRIC_NM:2 * _lib.METRIC_NM + 4].cpu()
    assert [int(x) for x in raw] == list(ometrics.delta_counts(pred, gt))          # n_valid, c1, c2, c3: bit-exact


def test_c3_dorn_fused_vs_oracle():
    """C3 at full size (logits 8 x 136 x 257 x 353 = 395 MB, K = 68): decode bit-exact, depth, ordinal loss and the
    FULL logit gradient against the oracle (network/Dorn.py:292-321, modules/dorn.py:95-107, criteria.py:744-787);
    the oracle's gradient is taken in fp64."""
    from mono_depth_estimation_b200 import dorn as D
    shape = synth.SHAPES["C3"]
    N, C2, H, W = shape
    K = C2 // 2
    x, gt = synth.dorn_inputs(shape, 103)
    # fp32 oracle: decode (the bit-exact artefact), P, label -> depth
    with torch.no_grad():
        dec_o, _ = odorn.ordinal_layer(x)
        depth_o = odorn.label_to_depth(dec_o, 0.001, 1.0, K)
    # fp64 oracle: loss and gradient
    x64 = x.double().requires_grad_(True)
    _, P64 = odorn.ordinal_layer(x64)
    y64 = odorn.depth_to_label(gt, 0.001, 1.0, K).double()
    l64 = odorn.ord_loss(P64, y64)
    (g64,) = torch.autograd.grad(l64, x64)
    del P64, x64
    xr = x.cuda().requires_grad_(True)
    loss, decode, depth, _ = D.dorn_fused(xr, gt.cuda(), K, 0.001, 1.0)
    loss.backward()
    assert decode.dtype == torch.int64 and torch.equal(decode.cpu(), dec_o)          # bit-exact
    close(depth, depth_o, 1e-6)
    close(loss, l64.detach(), LOSS_RTOL)
    grad_close(xr.grad, g64)


def test_c4_vnl_vs_oracle():
    """C4 at full size (8 x 1 x 385 x 385, 100 000 supplied triplets shared by the batch): loss, the FULL gradient and
    the two integers of the trim (valid triplets, dropped quarter) against the oracle (criteria.py:990-1045), fp64."""
    from mono_depth_estimation_b200 import criteria as Cr
    gt, pred, trip = synth.vnl_inputs(synth.SHAPES["C4"], 104)
    p64 = pred.double().clone().requires_grad_(True)
    l64, per, mask = ovnl.vnl_loss(gt.double(), p64, trip, 519.0, 519.0, return_parts=True)
    (g64,) = torch.autograd.grad(l64, p64)
    H, W = gt.shape[-2:]
    v = Cr.VNL_Loss(519.0, 519.0, (H, W))
    v.set_triplets(trip.cuda())
    p = pred.cuda().clone().requires_grad_(True)
    loss = v(gt.cuda(), p)
    loss.backward()
    stats = v.last_stats.cpu()
    assert int(stats[0]) == int(mask.sum())                       # same valid-triplet set
    assert int(stats[1]) == int(int(mask.sum()) * 0.25)           # criteria.py:1042-1043
    close(loss, l64.detach(), LOSS_RTOL)
    if float(stats[4]) == 1.0:                                    # no tie at the trim threshold: the kept set is the oracle's
        grad_close(p.grad, g64)
    else:                                                         # ties: same loss, the tied triplets share the weight
        close(p.grad.double().cpu().sum(), g64.sum(), 1e-4, 1e-6 * float(g64.abs().sum()))


def test_c5_full_nyu_eval_vs_oracle():
    """C5: all 654 images 480 x 640 in one launch (MetricComputation.compute_batch / fused_metrics(per_image=True)) against the
    oracle evaluated image by image as the reference's test loop does (batch size 1, modules/base_module.py:71-76;
    mean over images of per-image means, metrics.py:35-41,58-67): per-image valid / delta counts bit-exact for every image,
    per-image float metrics and the dataset values within 1e-5."""
    from mono_depth_estimation_b200 import metrics as M
    n_img = 654
    # generated in slabs (the oracle walks the images anyway); same tensors on both sides
    slabs = [synth.depth_pair((109, 1, 480, 640), synth.SEEDS["C5"] + i) for i in range(6)]
    pred = torch.cat([s[0] for s in slabs]); gt = torch.cat([s[1] for s in slabs])
    del slabs
    assert pred.shape[0] == n_img
    res = M.fused_metrics(pred.cuda(), gt.cuda(), names=EVAL, per_image=True)
    pir = res["per_image_raw"].cpu()
    piv = res["per_image"].cpu()
    idx = [_lib.METRIC_INDEX[n] for n in EVAL]
    acc = np.zeros(len(EVAL))
    for b in range(n_img):
        p, t = pred[b:b + 1], gt[b:b + 1]
        assert [int(v) for v in pir[b, :4]] == list(ometrics.delta_counts(p, t)), "image %d" % b
        v64 = np.array([float(v) for v in ometrics.compute(p.double(), t.double(), EVAL)])
        np.testing.assert_allclose(piv[b, idx].numpy(), v64, rtol=1e-5, err_msg="image %d" % b)
        acc += v64
    close(res["f64"][_lib.METRIC_NM:2 * _lib.METRIC_NM][idx], acc / n_img, 1e-5)            # mean over images of per-image means
    mc = M.MetricComputation(EVAL, strict=False)
    close(torch.stack(mc.compute_batch(pred.cuda(), gt.cuda())), acc / n_img, 1e-5)
    assert float(res["f64"][2 * _lib.METRIC_NM + _lib.METRIC_NQ]) == n_img
