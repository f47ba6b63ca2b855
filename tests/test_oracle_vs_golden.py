"""The CPU oracle restatement must reproduce the golden vectors that oracle/gen_golden.py made by
running the reference's own code (criteria.py, metrics.py, network/Dorn.py). This is what PINS
the oracle; the GPU parity tests then compare the CUDA path with the oracle."""
import numpy as np
import pytest
import torch

from oracle import dorn as odorn
from oracle import losses as olosses
from oracle import metrics as ometrics
from oracle import vnl as ovnl

T = torch.from_numpy


def close(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", ["l1", "mse", "berhu", "laina_berhu", "silog", "eigen"])
def test_losses_fp64_and_fp32(golden, name):
    g = golden("losses_small.npz")
    pred, gt = T(g["pred"]), T(g["gt"])
    fn = olosses.LOSSES[name]
    l64, g64 = olosses.loss_and_grad(fn, pred.double(), gt.double())
    close(l64, g[f"{name}_loss64"], 1e-12)
    close(g64, g[f"{name}_grad64"], 1e-10, 1e-15)
    l32, g32 = olosses.loss_and_grad(fn, pred, gt)
    # same op order as the reference -> fp32 results agree to a few ulp
    close(l32, g[f"{name}_loss32"], 2e-6)
    close(g32, g[f"{name}_grad32"], 1e-5, 1e-9)
    # and the fp32 reference itself is within 1e-5 of its fp64 evaluation (the tolerance budget)
    close(g[f"{name}_loss32"], g[f"{name}_loss64"], 1e-5)


def test_laina_variants(golden):
    g = golden("losses_small.npz")
    pred, gt = T(g["pred"]).double(), T(g["gt"]).double()
    l, gr = olosses.loss_and_grad(olosses.laina_berhu, pred, gt, use_logs=False)
    close(l, g["laina_nolog_loss64"], 1e-12); close(gr, g["laina_nolog_grad64"], 1e-10, 1e-15)
    l, gr = olosses.loss_and_grad(olosses.laina_berhu, pred, gt, size_average=False)
    close(l, g["laina_sum_loss64"], 1e-12); close(gr, g["laina_sum_grad64"], 1e-10, 1e-15)
    l, gr = olosses.loss_and_grad(olosses.laina_berhu, pred, gt, T(g["laina_mask_mask"]))
    close(l, g["laina_mask_loss64"], 1e-12); close(gr, g["laina_mask_grad64"], 1e-10, 1e-15)
    l, gr = olosses.loss_and_grad(olosses.silog, pred, gt, 0.5)
    close(l, g["silog_vf05_loss64"], 1e-12); close(gr, g["silog_vf05_grad64"], 1e-10, 1e-15)


def test_berhu_worked_example(golden):
    g = golden("losses_small.npz")
    l, gr = olosses.loss_and_grad(olosses.berhu, T(g["berhu_ex_pred"]), T(g["berhu_ex_gt"]))
    close(l, g["berhu_ex_loss"], 1e-6)
    close(gr, g["berhu_ex_grad"], 1e-6)
    close(l, 3.15, 1e-6)
    close(gr.flatten(), [0.25, 0.25, 0.0, -1.75], 1e-6)


def test_laina_ties(golden):
    g = golden("losses_small.npz")
    l, gr = olosses.loss_and_grad(olosses.laina_berhu, T(g["laina_tie_pred"]).double(), T(g["laina_tie_gt"]).double())
    close(l, g["laina_tie_loss64"], 1e-12)
    close(gr, g["laina_tie_grad64"], 1e-10)


def test_all_invalid_is_nan(golden):
    g = golden("losses_small.npz")
    t = torch.zeros(1, 1, 4, 4); p = torch.ones(1, 1, 4, 4)
    for name in ("l1", "mse", "silog"):
        assert np.isnan(g[f"{name}_allinvalid"])
        assert torch.isnan(olosses.LOSSES[name](p, t))


def test_dim_mismatch_raises():
    with pytest.raises(AssertionError, match="inconsistent dimensions"):
        olosses.masked_l1(torch.ones(2, 3, 4), torch.ones(2, 1, 3, 4))


def test_metrics(golden):
    g = golden("metrics_small.npz")
    names = [str(n) for n in g["names"]]
    pred, gt = T(g["pred"]), T(g["gt"])
    v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), names)]
    close(v64, g["values64"], 1e-12)
    v32 = [float(v) for v in ometrics.compute(pred, gt, names)]
    close(v32, g["values32"], 2e-6)
    n, c1, c2, c3 = ometrics.delta_counts(pred, gt)
    assert n == int(g["n_valid"])
    assert [c1, c2, c3] == [int(c) for c in g["delta_counts"]]
    per = [[float(v) for v in ometrics.compute(pred[b:b + 1].double(), gt[b:b + 1].double(), names)]
           for b in range(pred.shape[0])]
    close(per, g["per_image64"], 1e-12)
    close(ometrics.compute_per_image_mean(pred.double(), gt.double(), names), g["per_image64"].mean(0), 1e-12)


def test_metric_thresholds_strict(golden):
    g = golden("metrics_small.npz")
    v = ometrics.compute(T(g["thr_pred"]), T(g["thr_gt"]), ["delta1", "delta2", "delta3"])
    close([float(x) for x in v], g["thr_values"], 0, 0)
    n, c1, c2, c3 = ometrics.delta_counts(T(g["thr_pred"]), T(g["thr_gt"]))
    assert (n, c1, c2, c3) == (7, 1, 5, 6)   # only 1.2499999 is < 1.25; exact 1.25^k ratios are NOT counted


def test_metric_running_avg_and_errors(golden):
    g = golden("metrics_small.npz")
    pred, gt = T(g["pred"]), T(g["gt"])
    rm = ometrics.RunningMetrics(["absrel", "mae"])
    rm.compute(pred[:2], gt[:2]); rm.compute(pred[2:], gt[2:])
    close([float(rm.avg("absrel")), float(rm.avg(1))], g["running_avg"], 1e-6)
    with pytest.raises(AssertionError, match="invalid target!"):
        ometrics.compute(pred, torch.zeros_like(gt), ["mae"])
    with pytest.raises(KeyError):
        ometrics.compute(pred, gt, ["rmsle"])     # reference test.py:71 names a key that does not exist


def test_dorn_layer_and_losses(golden):
    g = golden("dorn_small.npz")
    x, gt = T(g["logits"]), T(g["gt"])
    K = x.shape[1] // 2
    xr = x.clone().requires_grad_(True)
    decode, P = odorn.ordinal_layer(xr)
    assert decode.dtype == torch.int64 and tuple(decode.shape) == (x.shape[0], 1) + tuple(x.shape[2:])
    assert np.array_equal(decode.numpy(), g["decode"])             # bit-exact
    assert np.array_equal(P.detach().numpy(), g["P"])              # same ops -> same bits on CPU
    depth = odorn.label_to_depth(decode, 0.001, 1.0, K)
    y = odorn.depth_to_label(gt, 0.001, 1.0, K)
    assert np.array_equal(depth.numpy(), g["depth"])
    assert np.array_equal(y.numpy(), g["y_sid"])
    loss = odorn.ord_loss(P, y)
    (gx,) = torch.autograd.grad(loss, xr)
    close(loss.detach(), g["ordloss32"], 2e-6)
    close(gx, g["ordloss_gradx32"], 1e-5, 1e-10)
    xd = x.double().clone().requires_grad_(True)
    _, P64 = odorn.ordinal_layer(xd)
    y64 = K * torch.log(gt.double() / 0.001) / np.log(1.0 / 0.001)
    l64 = odorn.ord_loss(P64, y64)
    (g64,) = torch.autograd.grad(l64, xd)
    close(l64.detach(), g["ordloss64"], 1e-12)
    close(g64, g["ordloss_gradx64"], 1e-10, 1e-16)
    Pl = T(g["P"]).clone().requires_grad_(True)
    (gP,) = torch.autograd.grad(odorn.ord_loss(Pl, y), Pl)
    close(gP, g["ordloss_gradP32"], 1e-6, 0)


@pytest.mark.parametrize("disc", ["SID", "UD"])
def test_ordinal_regression_loss(golden, disc):
    g = golden("dorn_small.npz")
    prob, gt = T(g["orl_prob"]), T(g["orl_gt"])
    K = prob.shape[1] // 2
    pr = prob.clone().requires_grad_(True)
    l = odorn.ordinal_regression_loss(pr, gt, K, torch.tensor(0.001), torch.tensor(1.0), disc)
    (gr,) = torch.autograd.grad(l, pr)
    close(l.detach(), g[f"orl_{disc}_loss"], 2e-6)
    close(gr, g[f"orl_{disc}_grad"], 1e-6, 0)


@pytest.mark.parametrize("sel", [True, False])
def test_vnl(golden, sel):
    g = golden("vnl_small.npz")
    gt, pred, trip = T(g["gt"]), T(g["pred"]), T(g["trip"])
    for tag, dt, rt in (("64", torch.float64, 1e-11), ("32", torch.float32, 1e-5)):
        p = pred.to(dt).clone().requires_grad_(True)
        l = ovnl.vnl_loss(gt.to(dt), p, trip, 519.0, 519.0, select=sel)
        (gr,) = torch.autograd.grad(l, p)
        close(l.detach(), g[f"loss{tag}_sel{int(sel)}"], rt)
        scale = np.abs(g[f"grad{tag}_sel{int(sel)}"]).max()
        close(gr, g[f"grad{tag}_sel{int(sel)}"], rt * 10, rt * scale)


# ---- VNL's ModelLoss other half: WCEL_Loss, depth_to_bins, bins_to_depth -----------------
@pytest.mark.parametrize("tag,C", [("c150", 150), ("c24", 24)])
def test_wcel_and_bins(golden, tag, C):
    from oracle import wcel as ow
    g = golden("wcel_small.npz")
    p = ow.vnl_params(0.01, 1.1, C)
    gt = T(g[f"{tag}_gt"]).clone()
    bins = ow.depth_to_bins(gt, p)
    assert bins.dtype == torch.int32 and np.array_equal(bins.numpy(), g[f"{tag}_bins"])        # integer: exact
    assert np.array_equal(gt.numpy(), g[f"{tag}_gt_after"])                                    # in-place clamp / -1 restore
    logits = T(g[f"{tag}_logits"])
    for dt, sfx, rt in ((torch.float64, "64", 1e-12), (torch.float32, "32", 2e-6)):
        lg = logits.to(dt).requires_grad_(True)
        loss = ow.wcel_loss(lg, bins, gt, p["wce_loss_weight"], C)
        (gr,) = torch.autograd.grad(loss, lg)
        close(loss.detach(), g[f"{tag}_loss{sfx}"], rt)
        close(gr, g[f"{tag}_grad{sfx}"], 1e-5 if sfx == "32" else 1e-10, 1e-9 if sfx == "32" else 1e-15)
    close(g[f"{tag}_loss32"], g[f"{tag}_loss64"], 1e-5)
    sm = T(g[f"{tag}_softmax"]).requires_grad_(True)
    d = ow.bins_to_depth(sm, p)
    close(d.detach(), g[f"{tag}_depth32"], 1e-6)
    close(d.detach(), g[f"{tag}_depth64"], 1e-5)
    (gx,) = torch.autograd.grad(d.sum(), sm)
    close(gx, g[f"{tag}_depth_gradsum32"], 1e-5, 1e-9)


def test_model_loss(golden):
    from oracle import wcel as ow
    g = golden("wcel_small.npz")
    C = 24
    p = ow.vnl_params(0.01, 1.1, C)
    gt = T(g["ml_gt"]).clone()
    bins = ow.depth_to_bins(gt, p)
    assert np.array_equal(bins.numpy(), g["ml_bins"]) and np.array_equal(gt.numpy(), g["ml_gt_after"])
    lg = T(g["ml_logits"]).double().requires_grad_(True)
    pd = T(g["ml_pred"]).double().requires_grad_(True)
    H, W = gt.shape[-2:]
    vnl = lambda gt_, pred_: ovnl.vnl_loss(gt_, pred_, T(g["ml_trip"]), 519.0, 519.0)
    total = ow.model_loss(pd, lg, bins, gt.double(), p, vnl, 6.0)
    g_lg, g_pd = torch.autograd.grad(total, (lg, pd))
    close(total.detach(), g["ml_total32"], 2e-5)
    close(g_lg, g["ml_grad_logits32"], 1e-4, 1e-8)
    close(g_pd, g["ml_grad_pred32"], 1e-3, 1e-6 * float(np.abs(g["ml_grad_pred32"]).max()))


def test_midas_scale_and_shift(golden):
    from oracle import midas as om
    g = golden("midas_small.npz")
    pred, target = T(g["pred"]), T(g["target"])
    s64, t64 = om.compute_scale_and_shift(pred.double(), target.double())
    close(s64, g["scale64"], 1e-12); close(t64, g["shift64"], 1e-12)
    s32, t32 = om.compute_scale_and_shift(pred, target)
    close(s32, g["scale32"], 1e-6); close(t32, g["shift32"], 1e-6, 1e-7)
    assert float(s64[3]) == 0.0 and float(t64[3]) == 0.0 and float(s64[4]) == 0.0      # singular systems
    sm, tm = om.compute_scale_and_shift(pred.double(), target.double(), T(g["mask"]).double())
    close(sm, g["scale_mask64"], 1e-12); close(tm, g["shift_mask64"], 1e-12)
    # the fp32 reference is within the tolerance budget of its fp64 evaluation
    close(g["scale32"][:3], g["scale64"][:3], 1e-4); close(g["shift32"][:3], g["shift64"][:3], 1e-4, 1e-5)


@pytest.mark.parametrize("name,kw", [("mse", dict(alpha=0.5, loss="mse")), ("l1", dict(alpha=0.5, loss="l1")),
                                     ("trim", dict(alpha=0.5, loss="trim")), ("mse_a0", dict(alpha=0.0, loss="mse")),
                                     ("mse_s2", dict(alpha=0.25, scales=2, loss="mse")),
                                     ("ssimse", dict(alpha=0.5, loss="ssimse")), ("ssil1", dict(alpha=0.5, loss="ssil1")),
                                     ("ssimse_a0", dict(alpha=0.0, loss="ssimse"))])
def test_midas_loss(golden, name, kw):
    from oracle import midas as om
    g = golden("midas_small.npz")
    pred, target = T(g["ml_pred"]), T(g["ml_target"])
    if "ssi" in name:
        pred = 0.7 / pred.clamp_min(0.3) + 0.2
    p = pred.double().requires_grad_(True)
    l = om.midas_loss(p, target.double(), **kw)
    (gr,) = torch.autograd.grad(l, p)
    close(l.detach(), g[f"ml_{name}_loss64"], 1e-12)
    close(gr, g[f"ml_{name}_grad64"], 1e-10, 1e-15)
    close(g[f"ml_{name}_loss32"], g[f"ml_{name}_loss64"], 1e-5)
    if name == "trim":   # as written the reference trims nothing: same value and gradient as l1
        close(g["ml_trim_loss64"], g["ml_l1_loss64"], 1e-12)
        close(g["ml_trim_grad64"], g["ml_l1_grad64"], 1e-12, 1e-15)


@pytest.mark.parametrize("name,kw", [("tp", dict(alpha=0.5)), ("tp_a0", dict(alpha=0.0)), ("tp_s2", dict(alpha=0.25, scales=2))])
def test_trimmed_procrustes(golden, name, kw):
    """TrimmedProcrustesLoss (criteria.py:335-363). The reference only runs in fp32 (its mask and statistics are
    hard-coded float32, :137 / :348), so the golden vectors are fp32; the oracle reproduces them with the same op
    sequence and its fp64 evaluation stays within the 1e-5 budget of them."""
    from oracle import midas as om
    g = golden("midas_small.npz")
    pred, target = T(g["tp_pred"]), T(g["tp_target"])
    p = pred.clone().requires_grad_(True)
    l, ssi = om.trimmed_procrustes_loss(p, target, **kw)
    (gr,) = torch.autograd.grad(l, p)
    close(l.detach(), g[f"{name}_loss32"], 1e-6)
    close(gr, g[f"{name}_grad32"], 1e-5, 1e-6 * float(np.abs(g[f"{name}_grad32"]).max()))
    if name == "tp":
        close(ssi.detach(), g["tp_ssi32"], 1e-6, 1e-7)
        close(om.normalize_prediction_robust(target.squeeze(1)), g["tp_tnorm32"], 1e-6, 1e-7)
    p64 = pred.double().requires_grad_(True)
    l64, _ = om.trimmed_procrustes_loss(p64, target.double(), **kw)
    (g64,) = torch.autograd.grad(l64, p64)
    close(l64.detach(), g[f"{name}_loss32"], 1e-5)
    for b in range(pred.shape[0]):
        close(g64[b], g[f"{name}_grad32"][b], 1e-4, 2e-6 * float(g64[b].abs().max()))
    assert float(g64[4].abs().max()) == 0.0                      # image without a valid pixel
    assert float(g64[3].abs().max()) > 1e3                       # constant prediction: scale clamped to 1e-6


STDEPTH_CASES = [("silma", "silma", 10), ("silms", "silms", 10), ("mse", "mse", 10), ("mae", "mae", 10),
                 ("silma_fb", "silma+fbdivergence", 10), ("mae_mse_fb", "mae+mse+fbdivergence", 10), ("silma20", "silma", 20),
                 ("mae20_fb", "mae+fbdivergence", 20)]


@pytest.mark.parametrize("name,loss_name,C", STDEPTH_CASES)
def test_stdepth_base_criterion(golden, name, loss_name, C):
    """BaseModule.setup_criterion's closure (modules/base_module.py:124-208), masked-reduction terms."""
    from oracle import stdepth as ost
    g = golden("stdepth_small.npz")
    pred, targ, rgba = T(g[f"pred{C}"]), T(g[f"targ{C}"]), T(g[f"rgba{C}"])
    kw = dict(variance_focus=0.85, depth_w=0.7, fbdiv_w=0.3, single_layer=(C == 10))
    p = pred.double().requires_grad_(True)
    l64, d64 = ost.stdepth_loss(p, targ.double(), rgba.double(), loss_name, **kw)
    (g64,) = torch.autograd.grad(l64, p)
    close(l64.detach(), g[f"{name}_loss64"], 1e-12)
    close(g64, g[f"{name}_grad64"], 1e-10, 1e-15)
    for k, v in d64.items():
        close(v.detach(), g[f"{name}_{k}64"], 1e-12)
    p32 = pred.clone().requires_grad_(True)
    l32, _ = ost.stdepth_loss(p32, targ, rgba, loss_name, **kw)
    (g32,) = torch.autograd.grad(l32, p32)
    close(l32.detach(), g[f"{name}_loss32"], 2e-6)
    close(g32.reshape(-1)[::7], g[f"{name}_grad32_s7"], 1e-5, 1e-9)
    close(g[f"{name}_loss32"], g[f"{name}_loss64"], 1e-5)


def test_stdepth_empty_depth_mask(golden):
    from oracle import stdepth as ost
    g = golden("stdepth_small.npz")
    l, d = ost.stdepth_loss(T(g["e_pred"]), T(g["e_targ"]), T(g["e_rgba"]), "silma", depth_w=0.7)
    assert float(d["depth_silog"]) == 0.0 and float(g["e_depth_silog32"]) == 0.0       # NaN -> nan_to_num -> 0
    close(l, g["e_loss32"], 2e-6)


def test_pointcloud_oracle_vs_reference_golden(golden):
    """depth -> point cloud: the golden vectors were produced by the REFERENCE's own point_cloud (depth2pointcloud.py:12-31,
    compiled from the reference source by oracle/gen_golden.py); the numpy restatement must reproduce them bit for bit,
    including the NaN pattern (clip planes, strict comparisons) and the sign of zero at invalid pixels."""
    from oracle import pointcloud as opc
    g = golden("pointcloud.npz")
    angle_x, clip_start, clip_end = (float(v) for v in g["camera"])
    for tag in "abc":
        ref = g["points_" + tag]
        out = opc.point_cloud(g["depth_" + tag], angle_x, clip_start, clip_end)
        assert out.dtype == ref.dtype == np.float64 and out.shape == ref.shape
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(out), np.nan_to_num(ref))
        assert np.array_equal(np.signbit(out), np.signbit(ref))
