"""Live cross-check of the CPU oracle against the reference's OWN code, imported by file path from
/root/reference. Runs only where that tree exists (this build container); on the GPU box the
committed golden vectors (tests/test_oracle_vs_golden.py) carry the same evidence."""
import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import _ref_loader as R
from oracle import dorn as odorn
from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import vnl as ovnl

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")


def _grad(fn, pred, *args, **kw):
    p = pred.detach().clone().requires_grad_(True)
    loss = fn(p, *args, **kw)
    (g,) = torch.autograd.grad(loss, p)
    return loss.detach(), g


@pytest.mark.parametrize("seed", [1, 2])
def test_losses_random(seed):
    crit = R.load("criteria")
    pred, gt = synth.depth_pair((3, 1, 40, 56), 100 + seed, border=3)
    pairs = [(crit.MaskedL1Loss(), olosses.masked_l1), (crit.MaskedMSELoss(), olosses.masked_mse),
             (crit.berHuLoss(), olosses.berhu), (crit.LainaBerHuLoss(), olosses.laina_berhu),
             (crit.silog_loss(0.85), lambda p, t: olosses.silog(p, t, 0.85)), (crit.MaskedDepthLoss(), olosses.eigen_masked_depth)]
    for ref_mod, ofn in pairs:
        for dt in (torch.float32, torch.float64):
            lr, gr = _grad(ref_mod, pred.to(dt), gt.to(dt))
            lo, go = _grad(ofn, pred.to(dt), gt.to(dt))
            tol = 1e-12 if dt == torch.float64 else 3e-6
            np.testing.assert_allclose(lo.numpy(), lr.numpy(), rtol=tol)
            np.testing.assert_allclose(go.numpy(), gr.numpy(), rtol=tol * 10, atol=tol * float(gr.abs().max()))


def test_metrics_c1():
    met = R.load("metrics")
    pred, gt = synth.config_inputs("C1")
    names = synth.DEFAULT_EVAL_METRICS
    ref = [float(v) for v in met.MetricComputation(names).compute(pred, gt)]
    mine = [float(v) for v in ometrics.compute(pred, gt, names)]
    np.testing.assert_allclose(mine, ref, rtol=1e-6)
    for k, fn in ((1, met.Delta1_multi_gpu), (2, met.Delta2_multi_gpu), (3, met.Delta3_multi_gpu)):
        p, t = ometrics.gather_valid(pred, gt)
        assert ometrics.delta_count(p, t, k) == int(round(float(fn(p.double(), t.double())) * p.numel()))


def test_dorn_layer_and_losses():
    crit, dn = R.load("criteria"), R.load("dorn_net")
    x, gt = synth.dorn_inputs((2, 24, 17, 23), 7)
    xr = x.clone().requires_grad_(True)
    dec_r, P_r = dn.OrdinalRegressionLayer()(xr)
    xo = x.clone().requires_grad_(True)
    dec_o, P_o = odorn.ordinal_layer(xo)
    assert torch.equal(dec_r, dec_o) and torch.equal(P_r, P_o)
    y = odorn.depth_to_label(gt, 0.001, 1.0, 12)
    lr = crit.ordLoss()(P_r, y); lo = odorn.ord_loss(P_o, y)
    (gr,) = torch.autograd.grad(lr, xr); (go,) = torch.autograd.grad(lo, xo)
    np.testing.assert_allclose(lo.detach().numpy(), lr.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(go.numpy(), gr.numpy(), rtol=1e-5, atol=1e-10)


def test_vnl_random():
    crit = R.load("criteria")
    H, W = 41, 57
    gt, pred, trip = synth.vnl_inputs((2, 1, H, W), 8, n_triplets=900, pad_rows=5, zero_frac=0.01)
    v = crit.VNL_Loss(519.0, 519.0, (H, W))
    v.fx, v.fy, v.u_u0, v.v_v0 = v.fx.double(), v.fy.double(), v.u_u0.double(), v.v_v0.double()
    t = trip.numpy()
    v.select_index = lambda: {"p1_x": t[0] % W, "p1_y": t[0] // W, "p2_x": t[1] % W, "p2_y": t[1] // W,
                              "p3_x": t[2] % W, "p3_y": t[2] // W}
    pr = pred.double().clone().requires_grad_(True)
    lr = v(gt.double(), pr)
    (gr,) = torch.autograd.grad(lr, pr)
    po = pred.double().clone().requires_grad_(True)
    lo = ovnl.vnl_loss(gt.double(), po, trip, 519.0, 519.0)
    (go,) = torch.autograd.grad(lo, po)
    np.testing.assert_allclose(lo.detach().numpy(), lr.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(go.numpy(), gr.numpy(), rtol=1e-9, atol=1e-14)


def test_point_cloud_random_vs_reference_function():
    """Live check (container with /root/reference): the reference's point_cloud, compiled from depth2pointcloud.py by
    oracle/gen_golden.py, against oracle/pointcloud.py on fresh random depth maps - bit for bit."""
    from types import SimpleNamespace
    from oracle import gen_golden, pointcloud as opc
    fn = gen_golden._reference_point_cloud()
    rs = np.random.RandomState(77)
    for shape in ((19, 23), (64, 48), (5, 260)):
        depth = (rs.rand(*shape) * 150).astype(np.float32)
        depth[rs.rand(*shape) < 0.1] = 0.0
        cam = SimpleNamespace(data=SimpleNamespace(angle_x=1.0471975511965976, clip_start=0.5, clip_end=120.0))
        ref = fn(depth, cam)
        out = opc.point_cloud(depth, cam.data.angle_x, cam.data.clip_start, cam.data.clip_end)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(out), np.nan_to_num(ref)) and np.array_equal(np.signbit(out), np.signbit(ref))
