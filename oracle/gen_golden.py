"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported by file path from
/root/reference, see oracle/_ref_loader.py) on small fixed inputs.

Run here (the container that has /root/reference):   python -m oracle.gen_golden
The fixtures are committed; the GPU box never needs /root/reference.

Every fixture stores the inputs themselves (they are small) plus the reference outputs in fp32
(what the reference computes) and, where useful, in fp64 (the same reference code fed double
tensors: separates "our error" from the reference's own fp32 summation error).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import _ref_loader as R  # noqa: E402
from mono_depth_estimation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _grad(fn, pred, *args):
    p = pred.detach().clone().requires_grad_(True)
    loss = fn(p, *args)
    (g,) = torch.autograd.grad(loss, p, allow_unused=True)
    if g is None:
        g = torch.zeros_like(p)
    return loss.detach(), g.detach()


def small_pair(seed, shape=(3, 1, 20, 28), border=2):
    return synth.depth_pair(shape, seed, border=border)


def gen_losses():
    crit = R.load("criteria")
    out = {}
    pred, gt = small_pair(11)
    out["pred"], out["gt"] = pred.numpy(), gt.numpy()
    mods = {
        "l1": crit.MaskedL1Loss(),
        "mse": crit.MaskedMSELoss(),
        "berhu": crit.berHuLoss(),
        "laina_berhu": crit.LainaBerHuLoss(),
        "silog": crit.silog_loss(0.85),
        "eigen": crit.MaskedDepthLoss(),
    }
    for name, m in mods.items():
        l32, g32 = _grad(m, pred, gt)
        l64, g64 = _grad(m, pred.double(), gt.double())
        out[f"{name}_loss32"] = l32.numpy()
        out[f"{name}_grad32"] = g32.numpy()
        out[f"{name}_loss64"] = l64.numpy()
        out[f"{name}_grad64"] = g64.numpy()
    # Laina variants: no logs / sum reduction / explicit mask
    mask = (gt > 2.0)
    for tag, m, extra in (
        ("laina_nolog", crit.LainaBerHuLoss(use_logs=False), ()),
        ("laina_sum", crit.LainaBerHuLoss(size_average=False), ()),
        ("laina_mask", crit.LainaBerHuLoss(), (mask,)),
    ):
        l64, g64 = _grad(m, pred.double(), gt.double(), *extra)
        l32, g32 = _grad(m, pred, gt, *extra)
        out[f"{tag}_loss32"], out[f"{tag}_grad32"] = l32.numpy(), g32.numpy()
        out[f"{tag}_loss64"], out[f"{tag}_grad64"] = l64.numpy(), g64.numpy()
    out["laina_mask_mask"] = mask.numpy()
    # silog with another variance focus
    l64, g64 = _grad(crit.silog_loss(0.5), pred.double(), gt.double())
    out["silog_vf05_loss64"], out["silog_vf05_grad64"] = l64.numpy(), g64.numpy()

    # worked example of SURVEY appendix A.1 (berHu)
    t = torch.tensor([[[[1.0, 2.0, 0.0, 4.0]]]])
    p = torch.tensor([[[[1.5, 2.1, 3.0, 1.0]]]])
    l, g = _grad(crit.berHuLoss(), p, t)
    out["berhu_ex_pred"], out["berhu_ex_gt"] = p.numpy(), t.numpy()
    out["berhu_ex_loss"], out["berhu_ex_grad"] = l.numpy(), g.numpy()

    # ties at the Laina maximum: the gradient through c is split evenly over tied maxima
    t = torch.tensor([[[[1.0, 1.0, 2.0, 2.0, 0.0, 3.0]]]])
    p = torch.tensor([[[[4.0, 4.0, 2.2, 1.0, 9.0, 3.3]]]])
    l, g = _grad(crit.LainaBerHuLoss(), p.double(), t.double())
    out["laina_tie_pred"], out["laina_tie_gt"] = p.numpy(), t.numpy()
    out["laina_tie_loss64"], out["laina_tie_grad64"] = l.numpy(), g.numpy()

    # all-invalid target: NaN losses (no assertion in the losses)
    t = torch.zeros(1, 1, 4, 4)
    p = torch.ones(1, 1, 4, 4)
    for name in ("l1", "mse", "silog"):
        with torch.no_grad():
            out[f"{name}_allinvalid"] = mods[name](p, t).numpy()
    np.savez_compressed(os.path.join(OUT, "losses_small.npz"), **out)


def gen_metrics():
    met = R.load("metrics")
    names = [n for n in met.METRICS if n != "ssim"]
    out = {"names": np.array(names)}
    pred, gt = small_pair(12, shape=(4, 1, 24, 32))
    pred[0, 0, 5, 5] = -0.5   # exercises clamp_min(1e-7)
    pred[1, 0, 6, 7] = 0.0
    out["pred"], out["gt"] = pred.numpy(), gt.numpy()
    mc = met.MetricComputation(names)
    out["values32"] = np.array([float(v) for v in mc.compute(pred, gt)], dtype=np.float64)
    mc64 = met.MetricComputation(names)
    out["values64"] = np.array([float(v) for v in mc64.compute(pred.double(), gt.double())], dtype=np.float64)
    # integer delta counts: recover from the reference's own mean * n
    p = torch.clamp_min(pred, 1e-7)[gt > 0]
    t = gt[gt > 0]
    ratio = torch.max(p / t, t / p)
    out["n_valid"] = np.int64(t.numel())
    out["delta_counts"] = np.array([int((ratio < 1.25 ** k).sum()) for k in (1, 2, 3)], dtype=np.int64)
    # per-image values (the eval loop calls compute() once per image)
    per = []
    for b in range(pred.shape[0]):
        m1 = met.MetricComputation(names)
        per.append([float(v) for v in m1.compute(pred[b:b + 1].double(), gt[b:b + 1].double())])
    out["per_image64"] = np.array(per, dtype=np.float64)

    # exact-threshold ratios: t = 1, p = 1.25^k exactly -> strict '<' must NOT count them
    t = torch.tensor([[[[1.0, 1.0, 1.0, 1.0, 2.0, 0.0, 4.0, 4.0]]]])
    p = torch.tensor([[[[1.25, 1.5625, 1.953125, 1.2499999, 2.5, 7.0, 5.0, 3.2]]]])
    mc = met.MetricComputation(["delta1", "delta2", "delta3"])
    out["thr_pred"], out["thr_gt"] = p.numpy(), t.numpy()
    out["thr_values"] = np.array([float(v) for v in mc.compute(p, t)])
    # running average semantics
    mc = met.MetricComputation(["absrel", "mae"])
    mc.compute(pred[:2], gt[:2])
    mc.compute(pred[2:], gt[2:])
    out["running_avg"] = np.array([float(mc.avg("absrel")), float(mc.avg(1))])
    np.savez_compressed(os.path.join(OUT, "metrics_small.npz"), **out)


def gen_dorn():
    crit = R.load("criteria")
    dn = R.load("dorn_net")
    out = {}
    K = 6
    logits, gt = synth.dorn_inputs((2, 2 * K, 9, 11), 13)
    # adversarial pairs: exact ties, both non-positive (clamp tie), near ties by 1..3 ulp, clamps at 1e4
    x = logits.clone()
    x[0, 0, 0, 0], x[0, 1, 0, 0] = 1.0, 1.0
    x[0, 2, 0, 0], x[0, 3, 0, 0] = -3.0, -0.5
    x[0, 4, 0, 0], x[0, 5, 0, 0] = 0.0, 1e-8
    x[0, 6, 0, 0], x[0, 7, 0, 0] = 2e4, 3e4
    x[0, 8, 0, 0], x[0, 9, 0, 0] = -1.0, 5e4
    x[0, 10, 0, 0], x[0, 11, 0, 0] = 30.0, 1.0
    for j, ulps in enumerate((1, 2, 3, 4)):
        for i, base in enumerate((0.3, 0.75, 1.0, 3.0, 100.0)):
            a = np.float32(base)
            b = a
            for _ in range(ulps):
                b = np.nextafter(b, np.float32(np.inf), dtype=np.float32)
            x[1, 2 * j, 1, i], x[1, 2 * j + 1, 1, i] = float(a), float(b)
    layer = dn.OrdinalRegressionLayer()
    xr = x.clone().requires_grad_(True)
    decode, P = layer(xr)
    out["logits"], out["gt"] = x.numpy(), gt.numpy()
    out["decode"], out["P"] = decode.numpy(), P.detach().numpy()
    alpha, beta = torch.tensor(0.001).float(), torch.tensor(1.0).float()
    Kt = torch.tensor(K).int()
    # modules/dorn.py:95-107 formulas executed verbatim on the reference's tensors
    depth = torch.exp(torch.log(alpha) + torch.log(beta / alpha) * decode / Kt)
    y_sid = Kt * torch.log(gt / alpha) / torch.log(beta / alpha)
    out["depth"], out["y_sid"] = depth.numpy(), y_sid.numpy()
    loss = crit.ordLoss()(P, y_sid)
    (gx,) = torch.autograd.grad(loss, xr)
    out["ordloss32"], out["ordloss_gradx32"] = loss.detach().numpy(), gx.numpy()
    # same in fp64
    xd = x.double().clone().requires_grad_(True)
    dec64, P64 = layer(xd)
    y64 = K * torch.log(gt.double() / 0.001) / np.log(1.0 / 0.001)
    l64 = crit.ordLoss()(P64, y64)
    (gx64,) = torch.autograd.grad(l64, xd)
    out["ordloss64"], out["ordloss_gradx64"] = l64.detach().numpy(), gx64.numpy()
    out["P64"] = P64.detach().numpy()
    # grad w.r.t. P alone (ordLoss(P, y) entry point)
    Pl = P.detach().clone().requires_grad_(True)
    lp = crit.ordLoss()(Pl, y_sid)
    (gP,) = torch.autograd.grad(lp, Pl)
    out["ordloss_gradP32"] = gP.numpy()

    # OrdinalRegressionLoss: prob = log-probabilities [K '<=' | K '>']
    prob = torch.log_softmax(torch.randn(2, 2 * K, 9, 11, generator=torch.Generator().manual_seed(5)), 1)
    gt2 = gt.clone()
    gt2[0, 0, 0, :3] = torch.tensor([0.0005, 0.001, 0.9999])   # gt < alpha -> label trunc toward zero
    for disc in ("SID", "UD"):
        orl = crit.OrdinalRegressionLoss(K, alpha, beta, disc)
        pr = prob.clone().requires_grad_(True)
        l = orl(pr, gt2)
        (g,) = torch.autograd.grad(l, pr)
        out[f"orl_{disc}_loss"], out[f"orl_{disc}_grad"] = l.detach().numpy(), g.numpy()
    out["orl_prob"], out["orl_gt"] = prob.numpy(), gt2.numpy()
    np.savez_compressed(os.path.join(OUT, "dorn_small.npz"), **out)


def gen_vnl():
    crit = R.load("criteria")
    out = {}
    H, W = 33, 41
    gt, pred, trip = synth.vnl_inputs((3, 1, H, W), 14, n_triplets=700, pad_rows=6, zero_frac=0.02)

    def make(dtype):
        v = crit.VNL_Loss(focal_x=519.0, focal_y=519.0, input_size=(H, W))
        if dtype == torch.float64:
            v.fx, v.fy = v.fx.double(), v.fy.double()
            v.u_u0, v.v_v0 = v.u_u0.double(), v.v_v0.double()
        t = trip.numpy()
        v.select_index = lambda: {"p1_x": t[0] % W, "p1_y": t[0] // W, "p2_x": t[1] % W,
                                  "p2_y": t[1] // W, "p3_x": t[2] % W, "p3_y": t[2] // W}
        return v

    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        for sel in (True, False):
            v = make(dt)
            p = pred.to(dt).clone().requires_grad_(True)
            l = v(gt.to(dt), p, select=sel)
            (g,) = torch.autograd.grad(l, p)
            out[f"loss{tag}_sel{int(sel)}"] = l.detach().numpy()
            out[f"grad{tag}_sel{int(sel)}"] = g.numpy()
    out["gt"], out["pred"], out["trip"] = gt.numpy(), pred.numpy(), trip.numpy()
    np.savez_compressed(os.path.join(OUT, "vnl_small.npz"), **out)


def gen_config_scalars():
    """Full-size C1 (the reference's CPU-runnable config): scalars + grad probes; inputs come
    from synth (seeded), so only ~KBs are stored."""
    crit = R.load("criteria")
    met = R.load("metrics")
    pred, gt = synth.config_inputs("C1")
    out = {}
    probe = np.random.RandomState(7).randint(0, pred.numel(), size=256)
    out["probe_idx"] = probe
    for name, m in (("berhu", crit.berHuLoss()), ("l1", crit.MaskedL1Loss()), ("mse", crit.MaskedMSELoss()),
                    ("silog", crit.silog_loss(0.85)), ("laina_berhu", crit.LainaBerHuLoss()),
                    ("eigen", crit.MaskedDepthLoss())):
        l64, g64 = _grad(m, pred.double(), gt.double())
        l32, g32 = _grad(m, pred, gt)
        out[f"{name}_loss64"], out[f"{name}_loss32"] = l64.numpy(), l32.numpy()
        out[f"{name}_gradprobe64"] = g64.flatten()[probe].numpy()
        out[f"{name}_gradsum64"] = g64.sum().numpy()
        out[f"{name}_gradabssum64"] = g64.abs().sum().numpy()
    names = synth.DEFAULT_EVAL_METRICS
    out["metric_names"] = np.array(names)
    out["metrics32"] = np.array([float(v) for v in met.MetricComputation(names).compute(pred, gt)])
    out["metrics64"] = np.array([float(v) for v in met.MetricComputation(names).compute(pred.double(), gt.double())])
    p = torch.clamp_min(pred, 1e-7)[gt > 0]
    t = gt[gt > 0]
    ratio = torch.max(p / t, t / p)
    out["n_valid"] = np.int64(t.numel())
    out["delta_counts"] = np.array([int((ratio < 1.25 ** k).sum()) for k in (1, 2, 3)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "config_c1.npz"), **out)


def _vnl_module_methods():
    """depth_to_bins / bins_to_depth of VNLModule (reference modules/vnl.py:202-230) compiled straight from
    the reference source: the module itself cannot be imported (pytorch_lightning is absent)."""
    import ast
    src = open(os.path.join(R.REF_ROOT, "modules", "vnl.py")).read()
    tree = ast.parse(src)
    fns = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("depth_to_bins", "bins_to_depth"):
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "np": np}
            exec(compile(mod, "modules/vnl.py", "exec"), ns)
            fns[node.name] = ns[node.name]
    assert len(fns) == 2
    return fns


def gen_wcel():
    """WCEL_Loss / ModelLoss by the reference's classes; depth_to_bins / bins_to_depth by the reference's
    method bodies run against a stand-in `self` carrying method.{depth_min,depth_max,dec_out_c},
    params.{depth_min_log,depth_bin_interval,depth_bin_border} and device."""
    import types
    from oracle import wcel as ow
    crit = R.load("criteria")
    fns = _vnl_module_methods()
    out = {}
    for C, tag in ((150, "c150"), (24, "c24")):
        p = ow.vnl_params(0.01, 1.1, C)
        self_ = types.SimpleNamespace(
            method=types.SimpleNamespace(depth_min=0.01, depth_max=1.1, dec_out_c=C),
            params=types.SimpleNamespace(depth_min_log=p["depth_min_log"], depth_bin_interval=p["depth_bin_interval"],
                                         depth_bin_border=p["depth_bin_border"]),
            device=torch.device("cpu"))
        g = torch.Generator().manual_seed(900 + C)
        B, H, W = (2, 6, 12) if C == 150 else (2, 12, 20)
        gt = torch.rand((B, 1, H, W), generator=g) * 1.3 + 0.002           # beyond both clamps
        gt[1, :, :3, :] = -1.0                                               # VNL padding
        gt[0, 0, 5, 5] = 0.01; gt[0, 0, 5, 6] = 1.1; gt[0, 0, 5, 7] = 0.0     # borders, zero (valid bin, not counted)
        logits = torch.randn((B, C, H, W), generator=g) * 3.0
        depth_in = gt.clone()
        bins = fns["depth_to_bins"](self_, depth_in)                          # mutates depth_in
        out[f"{tag}_gt"], out[f"{tag}_logits"] = gt.numpy(), logits.numpy()
        out[f"{tag}_bins"], out[f"{tag}_gt_after"] = bins.numpy(), depth_in.numpy()
        args = types.SimpleNamespace(wce_loss_weight=p["wce_loss_weight"], dec_out_c=C, focal_x=519.0, focal_y=519.0,
                                     crop_size=(H, W), diff_loss_weight=6.0)
        for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
            m = crit.WCEL_Loss(types.SimpleNamespace(**vars(args)))
            if dt == torch.float64:   # forward() casts the weight to fp32 (:851): feed an fp64 run through a patched cast
                w64 = m.weight.clone()
                lg = logits.double().requires_grad_(True)
                lp = torch.nn.functional.log_softmax(lg, 1)
                lp = torch.t(torch.transpose(lp, 0, 1).reshape(lp.size(1), -1))
                oh = (bins.reshape(-1, 1) == torch.arange(C, dtype=bins.dtype)).double()
                loss = -1 * torch.sum(torch.matmul(oh, w64) * lp) / torch.sum(depth_in > 0.).double()
            else:
                lg = logits.clone().requires_grad_(True)
                loss = m(lg, bins, depth_in)
            (gr,) = torch.autograd.grad(loss, lg)
            out[f"{tag}_loss{sfx}"], out[f"{tag}_grad{sfx}"] = loss.detach().numpy(), gr.numpy()
        # bins_to_depth on the softmax of the logits (what the decoder hands over, network/VNL.py:681)
        sm = torch.softmax(logits, 1)
        for dt, sfx in ((torch.float32, "32"),):
            x = sm.clone().requires_grad_(True)
            d = fns["bins_to_depth"](self_, x)
            (gx,) = torch.autograd.grad(d.sum(), x)
            out[f"{tag}_softmax"], out[f"{tag}_depth{sfx}"], out[f"{tag}_depth_gradsum{sfx}"] = sm.numpy(), d.detach().numpy(), gx.numpy()
        out[f"{tag}_depth64"] = (10 ** (sm.double().permute(0, 2, 3, 1) * torch.tensor(p["depth_bin_border"])).sum(3, keepdim=True)).permute(0, 3, 1, 2).numpy()
    # ModelLoss = WCEL + 6 * VNL on one small case with fixed triplets
    C = 24
    p = ow.vnl_params(0.01, 1.1, C)
    H, W = 12, 20
    gtv, predv, trip = synth.vnl_inputs((2, 1, H, W), 61, n_triplets=300, pad_rows=2)
    args = types.SimpleNamespace(wce_loss_weight=p["wce_loss_weight"], dec_out_c=C, focal_x=519.0, focal_y=519.0,
                                 crop_size=(H, W), diff_loss_weight=6.0)
    ml = crit.ModelLoss(args)
    v = ml.virtual_normal_loss
    t = trip.numpy()
    sel = {"p1_x": t[0] % W, "p1_y": t[0] // W, "p2_x": t[1] % W, "p2_y": t[1] // W, "p3_x": t[2] % W, "p3_y": t[2] // W}
    v.select_index = lambda: sel
    self_ = types.SimpleNamespace(
        method=types.SimpleNamespace(depth_min=0.01, depth_max=1.1, dec_out_c=C),
        params=types.SimpleNamespace(depth_min_log=p["depth_min_log"], depth_bin_interval=p["depth_bin_interval"],
                                     depth_bin_border=p["depth_bin_border"]), device=torch.device("cpu"))
    gt_in = gtv.clone()
    bins = fns["depth_to_bins"](self_, gt_in)
    logits = torch.randn((2, C, H, W), generator=torch.Generator().manual_seed(62)) * 2.0
    lg = logits.clone().requires_grad_(True)
    pd = predv.clone().requires_grad_(True)
    total = ml(pd, lg, bins, gt_in)
    g_lg, g_pd = torch.autograd.grad(total, (lg, pd))
    out.update({"ml_gt": gtv.numpy(), "ml_gt_after": gt_in.numpy(), "ml_pred": predv.numpy(), "ml_trip": trip.numpy(),
                "ml_logits": logits.numpy(), "ml_bins": bins.numpy(), "ml_total32": total.detach().numpy(),
                "ml_grad_logits32": g_lg.numpy(), "ml_grad_pred32": g_pd.numpy()})
    np.savez_compressed(os.path.join(OUT, "wcel_small.npz"), **out)


def gen_midas():
    """compute_scale_and_shift by the reference's own function (criteria.py:154-176), fp32 and fp64, default and
    explicit mask; one image with a single valid pixel and one without any (singular systems -> zeros)."""
    crit = R.load("criteria")
    g = torch.Generator().manual_seed(777)
    B, H, W = 5, 24, 36
    target = torch.rand((B, H, W), generator=g) * 9.5 + 0.5
    target[torch.rand((B, H, W), generator=g) < 0.25] = 0.0
    pred = (1.0 / target.clamp_min(0.3)) * 0.7 + 0.2 + torch.randn((B, H, W), generator=g) * 0.05   # disparity-like
    target[3] = 0.0; target[3, 4, 5] = 2.0        # one valid pixel: det == 0
    target[4] = 0.0                                # no valid pixel
    out = {"pred": pred.numpy(), "target": target.numpy()}
    for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
        s, t = crit.compute_scale_and_shift(pred.to(dt), target.to(dt))
        out["scale" + sfx], out["shift" + sfx] = s.numpy(), t.numpy()
    mask = (target > 1.0).float()
    s, t = crit.compute_scale_and_shift(pred.double(), target.double(), mask.double())
    out["mask"], out["scale_mask64"], out["shift_mask64"] = mask.numpy(), s.numpy(), t.numpy()
    # MidasLoss without ssi (the `my` method's criterion and its l1 / trim siblings), loss and gradient, fp32 + fp64
    g2 = torch.Generator().manual_seed(778)
    Bm, Hm, Wm = 3, 21, 30                       # not multiples of 8: the coarse grids end off the border
    tg = torch.rand((Bm, 1, Hm, Wm), generator=g2) * 9.5 + 0.5
    tg[torch.rand((Bm, 1, Hm, Wm), generator=g2) < 0.25] = 0.0
    pr = tg.clamp_min(0.4) + torch.randn((Bm, 1, Hm, Wm), generator=g2) * 0.3
    pr[0, 0, 4, 4:8] = tg[0, 0, 4, 4:8]          # exact ties: |0| has a zero subgradient
    out["ml_pred"], out["ml_target"] = pr.numpy(), tg.numpy()
    for name, kw in (("mse", dict(alpha=0.5, loss="mse")), ("l1", dict(alpha=0.5, loss="l1")), ("trim", dict(alpha=0.5, loss="trim")),
                     ("mse_a0", dict(alpha=0.0, loss="mse")), ("mse_s2", dict(alpha=0.25, scales=2, loss="mse")),
                     ("ssimse", dict(alpha=0.5, loss="ssimse")), ("ssil1", dict(alpha=0.5, loss="ssil1")),
                     ("ssimse_a0", dict(alpha=0.0, loss="ssimse"))):
        for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
            src = (0.7 / pr.clamp_min(0.3) + 0.2) if "ssi" in name else pr     # disparity-like input for the aligned variants
            p = src.to(dt).clone().requires_grad_(True)
            l = crit.MidasLoss(**kw)(p, tg.to(dt))
            (gr,) = torch.autograd.grad(l, p)
            out[f"ml_{name}_loss{sfx}"], out[f"ml_{name}_grad{sfx}"] = l.detach().numpy(), gr.numpy()
    # TrimmedProcrustesLoss (criteria.py:335-363, `midas --loss ssitrim`) and normalize_prediction_robust (:135-152):
    # two ordinary images, one with > 50 % invalid pixels (median = a masked zero), one with a constant prediction
    # (deviation 0: the scale is clamped to 1e-6), one without any valid pixel
    g3 = torch.Generator().manual_seed(779)
    Bt, Ht, Wt = 5, 22, 31
    tt = torch.rand((Bt, 1, Ht, Wt), generator=g3) * 9.5 + 0.5
    tt[torch.rand((Bt, 1, Ht, Wt), generator=g3) < 0.25] = 0.0
    tt[2][torch.rand((1, Ht, Wt), generator=g3) < 0.5] = 0.0
    tt[4] = 0.0
    pt = 0.7 / (tt.clamp_min(0.4) + torch.randn((Bt, 1, Ht, Wt), generator=g3) * 0.3).clamp_min(0.3) + 0.2
    pt[3] = 0.75
    out["tp_pred"], out["tp_target"] = pt.numpy(), tt.numpy()
    for name, kw in (("tp", dict(alpha=0.5)), ("tp_a0", dict(alpha=0.0)), ("tp_s2", dict(alpha=0.25, scales=2))):
        # fp32 only: the reference builds its mask and statistics as float32 (criteria.py:137,348) and its
        # `m[valid] = median(...)` refuses a float64 source
        for dt, sfx in ((torch.float32, "32"),):
            p = pt.to(dt).clone().requires_grad_(True)
            mod = crit.TrimmedProcrustesLoss(**kw)
            l = mod(p, tt.to(dt))
            (gr,) = torch.autograd.grad(l, p)
            out[f"{name}_loss{sfx}"], out[f"{name}_grad{sfx}"] = l.detach().numpy(), gr.numpy()
            if name == "tp":
                out[f"tp_ssi{sfx}"] = mod.prediction_ssi.detach().numpy()
                out[f"tp_tnorm{sfx}"] = crit.normalize_prediction_robust(tt.to(dt).squeeze(1)).numpy()
    np.savez_compressed(os.path.join(OUT, "midas_small.npz"), **out)


def _base_module_criterion(method, single_layer):
    """The closure BaseModule.setup_criterion returns (reference modules/base_module.py:124-208), compiled straight
    from the reference source (the module itself needs pytorch_lightning) and bound to a stand-in `self`; it calls
    the reference's own criteria.silog_loss and stdepth_utils functions."""
    import ast
    import importlib.util
    import types
    import torch.nn.functional as F
    src = open(os.path.join(R.REF_ROOT, "modules", "base_module.py")).read()
    node = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "setup_criterion")
    spec = importlib.util.spec_from_file_location("_mde_ref_stdepth_utils", os.path.join(R.REF_ROOT, "stdepth_utils.py"))
    su = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(su)
    ns = {"torch": torch, "F": F, "criteria": R.load("criteria"), "composite_layers": su.composite_layers,
          "depth_sort": su.depth_sort, "dssim2d": su.dssim2d}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "modules/base_module.py", "exec"), ns)
    return ns["setup_criterion"](types.SimpleNamespace(method=method, single_layer=single_layer))


from mono_depth_estimation_b200.synth import stdepth_inputs  # noqa: E402


def gen_stdepth():
    import types
    T = torch.from_numpy
    out = {}
    cases = [("silma", "silma", 10), ("silms", "silms", 10), ("mse", "mse", 10), ("mae", "mae", 10),
             ("silma_fb", "silma+fbdivergence", 10), ("mae_mse_fb", "mae+mse+fbdivergence", 10), ("silma20", "silma", 20),
             ("mae20_fb", "mae+fbdivergence", 20)]
    for C in (10, 20):
        pred, targ, rgba = stdepth_inputs(880 + C, 3, C, 19, 27)
        rgba[2, 3] = 0.0 if C == 20 else rgba[2, 3]                         # an image without mask1 pixels
        out[f"pred{C}"], out[f"targ{C}"], out[f"rgba{C}"] = pred.numpy(), targ.numpy(), rgba.numpy()
    for name, loss_name, C in cases:
        method = types.SimpleNamespace(loss=loss_name, variance_focus=0.85, depth_loss_weight=0.7, comp_loss_weight=1.0,
                                       fbdiv_loss_weight=0.3, ssim_loss_weight=1.0)
        crit = _base_module_criterion(method, single_layer=(C == 10))
        for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
            p = T(out[f"pred{C}"]).to(dt).requires_grad_(True)
            loss, ld = crit(p, T(out[f"targ{C}"]).to(dt), T(out[f"rgba{C}"]).to(dt), return_loss_dict=True)
            (gr,) = torch.autograd.grad(loss, p)
            out[f"{name}_loss{sfx}"] = loss.detach().numpy()
            if dt == torch.float64:
                out[f"{name}_grad{sfx}"] = gr.numpy()
            else:                                    # fp32 gradient: a strided sample keeps the fixture small
                out[f"{name}_grad{sfx}_s7"] = gr.numpy().reshape(-1)[::7].copy()
            for k, v in ld.items():
                out[f"{name}_{k}{sfx}"] = v.numpy()
    # empty depth mask: silog -> NaN -> nan_to_num -> 0 (base_module.py:126-127)
    pred, targ, rgba = stdepth_inputs(899, 2, 10, 8, 9)
    targ[:, 8:] = 0.0
    method = types.SimpleNamespace(loss="silma", variance_focus=0.85, depth_loss_weight=0.7, comp_loss_weight=1.0,
                                   fbdiv_loss_weight=0.3, ssim_loss_weight=1.0)
    p = pred.clone().requires_grad_(True)
    loss, ld = _base_module_criterion(method, True)(p, targ, rgba, return_loss_dict=True)
    out.update({"e_pred": pred.numpy(), "e_targ": targ.numpy(), "e_rgba": rgba.numpy(), "e_loss32": loss.detach().numpy(),
                "e_depth_silog32": ld["depth_silog"].numpy()})
    np.savez_compressed(os.path.join(OUT, "stdepth_small.npz"), **out)


def _reference_point_cloud():
    """`point_cloud` of reference depth2pointcloud.py:12-31 compiled from the reference source itself: the file is a
    Blender script (imports bpy / mathutils at the top, U+200B characters on its blank lines) and cannot be imported,
    but the function body only needs `np` and `tan`."""
    import ast
    from math import tan
    src = open(os.path.join(R.REF_ROOT, "depth2pointcloud.py"), encoding="utf-8").read().replace("\u200b", "")
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == "point_cloud":
            ns = {"np": np, "tan": tan}
            exec(compile(ast.Module(body=[node], type_ignores=[]), "depth2pointcloud.py", "exec"), ns)
            return ns["point_cloud"]
    raise RuntimeError("point_cloud not found in depth2pointcloud.py")


def gen_pointcloud():
    """Golden vectors of depth -> point cloud produced by the REFERENCE's function (pins oracle/pointcloud.py, the GPU
    kernel is compared with both). The camera is the stub the function reads: cam.data.{angle_x, clip_start, clip_end}."""
    from types import SimpleNamespace
    fn = _reference_point_cloud()
    out = {}
    for tag, shape, seed in (("a", (37, 53), 3), ("b", (48, 64), 11), ("c", (8, 1028), 12)):
        rs = np.random.RandomState(seed)
        depth = (rs.rand(*shape) * 12).astype(np.float32)
        depth[0, :5] = 0.05                      # below clip_start
        depth[3, 3] = 200.0                      # beyond clip_end
        depth[5, 7] = 0.1                        # exactly clip_start: strict '>' -> invalid
        cam = SimpleNamespace(data=SimpleNamespace(angle_x=0.8575560450553894, clip_start=0.1, clip_end=100.0))
        pts = fn(depth, cam)
        assert pts.dtype == np.float64 and pts.shape == shape + (3,)
        out["depth_" + tag] = depth
        out["points_" + tag] = pts
    out["camera"] = np.array([0.8575560450553894, 0.1, 100.0])
    np.savez_compressed(os.path.join(OUT, "pointcloud.npz"), **out)


def main():
    assert R.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if sys.argv[1:] == ["midas"]:
        gen_midas()
        return
    if sys.argv[1:] == ["stdepth"]:
        gen_stdepth()
        return
    if sys.argv[1:] == ["pointcloud"]:
        gen_pointcloud()
        return
    gen_losses()
    gen_metrics()
    gen_dorn()
    gen_vnl()
    gen_config_scalars()
    gen_wcel()
    gen_midas()
    gen_stdepth()
    gen_pointcloud()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
