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


def main():
    assert R.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gen_losses()
    gen_metrics()
    gen_dorn()
    gen_vnl()
    gen_config_scalars()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
