"""GPU parity of the fused masked losses (C ABI mde_masked_loss) against the pinned CPU oracle."""
import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import losses as olosses
from tests.gpu_util import LOSS_RTOL, T, close, grad_close, run_loss

pytestmark = pytest.mark.gpu
NAMES = ["l1", "mse", "berhu", "laina_berhu", "silog", "eigen"]


@pytest.fixture(scope="module")
def Cr():
    from mono_depth_estimation_b200 import criteria
    return criteria


def make(Cr, name):
    return {"l1": Cr.MaskedL1Loss, "mse": Cr.MaskedMSELoss, "berhu": Cr.berHuLoss, "laina_berhu": Cr.LainaBerHuLoss,
            "silog": lambda: Cr.silog_loss(0.85), "eigen": Cr.MaskedDepthLoss}[name]()


@pytest.mark.parametrize("name", NAMES)
def test_small_golden(Cr, golden, name):
    g = golden("losses_small.npz")
    pred, gt = T(g["pred"]).cuda(), T(g["gt"]).cuda()
    m = make(Cr, name)
    loss, grad = run_loss(m, pred, gt)
    assert loss.dim() == 0 and loss.is_cuda and grad.shape == pred.shape and grad.dtype == pred.dtype
    close(loss, g[f"{name}_loss64"], LOSS_RTOL)
    grad_close(grad, g[f"{name}_grad64"])
    if name not in ("laina_berhu", "silog"):
        assert m.loss is not None                     # side effect callers may read (criteria.py:63,76,89,131)
    # forward only (no autograd) gives the same value and no gradient work
    with torch.no_grad():
        close(make(Cr, name)(pred, gt), g[f"{name}_loss64"], LOSS_RTOL)


@pytest.mark.parametrize("name", NAMES)
def test_config_c1_golden(Cr, golden, name):
    g = golden("config_c1.npz")
    pred, gt = synth.config_inputs("C1")
    loss, grad = run_loss(make(Cr, name), pred.cuda(), gt.cuda())
    close(loss, g[f"{name}_loss64"], LOSS_RTOL)
    gd = grad.double().cpu()
    probe = torch.from_numpy(g["probe_idx"])
    scale = float(g[f"{name}_gradabssum64"]) / pred.numel()
    close(gd.flatten()[probe], g[f"{name}_gradprobe64"], 1e-5, 2e-5 * scale)
    close(gd.abs().sum(), g[f"{name}_gradabssum64"], 1e-5)
    close(gd.sum(), g[f"{name}_gradsum64"], 1e-5, 1e-5 * float(g[f"{name}_gradabssum64"]))


@pytest.mark.parametrize("name", ["silog", "berhu", "l1"])
def test_config_c2_vs_oracle(Cr, name):
    pred, gt = synth.config_inputs("C2")
    l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], pred.double(), gt.double())
    loss, grad = run_loss(make(Cr, name), pred.cuda(), gt.cuda())
    close(loss, l64, LOSS_RTOL)
    grad_close(grad, g64)


def test_laina_variants_and_silog_focus(Cr, golden):
    g = golden("losses_small.npz")
    pred, gt = T(g["pred"]).cuda(), T(g["gt"]).cuda()
    for tag, mod, extra in (("laina_nolog", Cr.LainaBerHuLoss(use_logs=False), ()),
                            ("laina_sum", Cr.LainaBerHuLoss(size_average=False), ()),
                            ("laina_mask", Cr.LainaBerHuLoss(), (T(g["laina_mask_mask"]).cuda(),))):
        loss, grad = run_loss(mod, pred, gt, *extra)
        close(loss, g[f"{tag}_loss64"], LOSS_RTOL, msg=tag)
        grad_close(grad, g[f"{tag}_grad64"], msg=tag)
    loss, grad = run_loss(Cr.silog_loss(0.5), pred, gt)
    close(loss, g["silog_vf05_loss64"], LOSS_RTOL)
    grad_close(grad, g["silog_vf05_grad64"])


def test_worked_examples(Cr, golden):
    g = golden("losses_small.npz")
    loss, grad = run_loss(Cr.berHuLoss(), T(g["berhu_ex_pred"]).cuda(), T(g["berhu_ex_gt"]).cuda())
    close(loss, 3.15, 1e-6)
    close(grad.flatten(), [0.25, 0.25, 0.0, -1.75], 1e-6)
    loss, grad = run_loss(Cr.LainaBerHuLoss(), T(g["laina_tie_pred"]).cuda(), T(g["laina_tie_gt"]).cuda())
    close(loss, g["laina_tie_loss64"], LOSS_RTOL)
    grad_close(grad, g["laina_tie_grad64"])          # gradient through c is split evenly over tied maxima


def test_empty_mask_and_errors(Cr):
    t = torch.zeros(1, 1, 4, 4).cuda(); p = torch.ones(1, 1, 4, 4).cuda()
    for m in (Cr.MaskedL1Loss(), Cr.MaskedMSELoss(), Cr.silog_loss(0.85)):
        assert torch.isnan(m(p, t))                   # reference: empty mask -> NaN, no exception
    with pytest.raises(AssertionError, match="inconsistent dimensions"):
        Cr.MaskedMSELoss()(torch.ones(2, 3, 4).cuda(), torch.ones(2, 1, 3, 4).cuda())


@pytest.mark.parametrize("name", ["l1", "silog", "berhu", "eigen"])
def test_layouts_tails_and_alignment(Cr, name):
    fn = olosses.LOSSES[name]
    # odd sizes (n % 4 != 0), 3-D input, non-contiguous view, 4-byte-offset storage (scalar path)
    for shape in ((3, 1, 33, 41), (2, 1, 7, 5), (4, 1, 16, 24)):
        pred, gt = synth.depth_pair(shape, 31, border=1)
        l64, g64 = olosses.loss_and_grad(fn, pred.double(), gt.double())
        loss, grad = run_loss(make(Cr, name), pred.cuda(), gt.cuda())
        close(loss, l64, LOSS_RTOL, msg=str(shape)); grad_close(grad, g64, msg=str(shape))
    pred, gt = synth.depth_pair((4, 1, 16, 24), 32, border=1)
    l64, g64 = olosses.loss_and_grad(fn, pred.double(), gt.double())
    buf = torch.zeros(pred.numel() + 1).cuda()
    buf[1:] = pred.flatten().cuda()
    shifted = buf[1:].view(pred.shape)                # data_ptr is 4 bytes off a 16-byte boundary
    loss, grad = run_loss(make(Cr, name), shifted, gt.cuda())
    close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
    wide = torch.zeros(4, 1, 16, 48).cuda()
    wide[..., ::2] = pred.cuda()
    loss, grad = run_loss(make(Cr, name), wide[..., ::2], gt.cuda())   # non-contiguous
    close(loss, l64, LOSS_RTOL); grad_close(grad, g64)
    if name != "eigen":
        p3, t3 = pred[:, 0], gt[:, 0]
        loss, grad = run_loss(make(Cr, name), p3.cuda(), t3.cuda())
        close(loss, l64, LOSS_RTOL); grad_close(grad, g64[:, 0])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_amp_pred_dtypes(Cr, dtype):
    pred, gt = synth.depth_pair((2, 1, 24, 32), 33, border=1)
    ph = pred.to(dtype)
    for name in ("silog", "berhu", "eigen"):
        l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], ph.double(), gt.double())
        loss, grad = run_loss(make(Cr, name), ph.cuda(), gt.cuda())
        assert grad.dtype == dtype
        close(loss, l64, LOSS_RTOL, msg=name)
        eps = 1e-3 if dtype == torch.float16 else 8e-3   # the gradient is ROUNDED to pred's dtype
        close(grad, g64, eps, eps * float(g64.abs().max()), msg=name)


def test_grad_output_scaling_and_reuse(Cr):
    pred, gt = synth.depth_pair((2, 1, 24, 32), 34, border=1)
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double())
    p = pred.cuda().requires_grad_(True)
    loss = Cr.silog_loss(0.85)(p, gt.cuda())
    (loss * 65536.0).backward()                        # GradScaler-style grad_output
    grad_close(p.grad / 65536.0, g64)
    # the loss composes with other autograd ops
    p2 = pred.cuda().requires_grad_(True)
    total = Cr.MaskedL1Loss()(p2 * 1.0, gt.cuda()) + 0.5 * Cr.MaskedMSELoss()(p2, gt.cuda())
    total.backward()
    _, ga = olosses.loss_and_grad(olosses.masked_l1, pred.double(), gt.double())
    _, gb = olosses.loss_and_grad(olosses.masked_mse, pred.double(), gt.double())
    grad_close(p2.grad, ga + 0.5 * gb)
    # many calls in a row on one workspace (parity sets alternate)
    m = Cr.berHuLoss()
    ref, _ = olosses.loss_and_grad(olosses.berhu, pred.double(), gt.double())
    for _ in range(5):
        close(run_loss(m, pred.cuda(), gt.cuda())[0], ref, LOSS_RTOL)


def test_shift_property_at_scale(Cr):
    """Size-independent checks at a size the CPU oracle is not run on: MaskedL1 and SILog are invariant
    under the transforms the math says they are, and the gradient integrates the loss."""
    pred, gt = synth.depth_pair((64, 1, 480, 640), 35, device="cuda")
    l1 = Cr.MaskedL1Loss()
    a = l1(pred, gt)
    b = l1(pred * 2, gt * 2)
    close(b, 2 * a, 1e-6)
    si = Cr.silog_loss(1.0)                            # lambda = 1: invariant to a global scale of pred
    close(si(pred * 1.7, gt), si(pred, gt), 1e-5)
    # MSE is quadratic in pred: L(p + h g) = L(p) + h |g|^2 + h^2 |g|^2 / n_valid  (g = dL/dp, zero off-mask)
    loss, grad = run_loss(Cr.MaskedMSELoss(), pred, gt)
    g2 = float((grad.double() ** 2).sum())
    n_valid = float((gt > 0).sum())
    h = 0.05 * float(loss) / g2
    with torch.no_grad():
        l2 = Cr.MaskedMSELoss()(pred + h * grad, gt)
    close(float(l2) - float(loss), h * g2 + h * h * g2 / n_valid, 1e-3)


@pytest.mark.parametrize("name", ["silog", "l1", "mse", "berhu", "laina_berhu"])
@pytest.mark.parametrize("names", [["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"],
                                   ["delta1", "absrel", "sqrel", "msle", "rmse_log", "rmse_true"]])
def test_loss_with_fused_metrics(Cr, name, names):
    """criterion.fuse_metrics(mc): ONE launch gives loss, gradient and the pooled metric suite; the following
    mc.compute(pred, gt) is served from it (no second pass) and equals the stand-alone results."""
    from mono_depth_estimation_b200 import metrics as M, _lib
    from oracle import metrics as ometrics
    for shape, seed in (((3, 1, 40, 56), 71), ((2, 1, 33, 41), 72)):
        pred, gt = synth.depth_pair(shape, seed, border=2)
        pred[0, 0, 5, 5] = 1e-8                       # below the metrics clamp (1e-7), above Laina clamp (1e-9)
        l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], pred.double(), gt.double())
        v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), names)]
        mc = M.MetricComputation(names)
        crit = make(Cr, name).fuse_metrics(mc)
        p = pred.cuda().requires_grad_(True)
        g = gt.cuda()
        _lib.workspace(p.device, shape[0])             # first use initialises the workspace (one tiny launch)
        n0 = _lib.launch_count()
        loss = crit(p, g)
        vals = mc.compute(p.detach(), g)
        assert _lib.launch_count() - n0 == 1, "metrics must come from the loss launch"
        loss.backward()
        close(loss, l64, LOSS_RTOL, msg=name)
        grad_close(p.grad, g64, msg=name)
        close(torch.stack(vals), v64, 1e-5, msg=name)
        assert mc.count == 1
        # a different tensor is NOT served from the cache
        vals2 = mc.compute((p.detach() * 1.0), g)
        assert _lib.launch_count() - n0 >= 2           # the loss launch (no scaling launch for a plain backward) + this one
        close(torch.stack(vals2), v64, 1e-5)


def _silog_both(Cr, pred, gt, names=("delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse")):
    """(plain loss, grad), (fused loss, grad, metric values) of SILog on the same tensors."""
    from mono_depth_estimation_b200 import metrics as M
    loss, grad = run_loss(Cr.silog_loss(0.85), pred.cuda(), gt.cuda())
    mc = M.MetricComputation(list(names), strict=False)
    crit = Cr.silog_loss(0.85).fuse_metrics(mc)
    p = pred.cuda().requires_grad_(True)
    lf = crit(p, gt.cuda())
    vals = mc.compute(p.detach(), gt.cuda())
    lf.backward()
    return (loss, grad), (lf.detach(), p.grad.detach(), torch.stack(vals))


@pytest.mark.parametrize("shape", [(2, 1, 40, 64), (4, 1, 480, 640)])
def test_silog_shared_memory_variant_rare_quads(Cr, shape):
    """The fast path of the SS kernels assumes valid targets > 0.01 and predictions >= 1e-7 per QUAD and sends
    every other quad through the exact arithmetic: targets in (0, 0.01] (valid for the metrics, masked for the
    loss, criteria.py:730), exactly 0.01, a subnormal target, predictions below the metrics' clamp.
    The small shape runs in the register-resident kernel (resident_loss.cuh: up to 2 quads per thread of one CTA per
    SM), 4x480x640 is just beyond its capacity and takes the shared-memory variant: the same cases for both."""
    from oracle import metrics as ometrics
    names = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
    pred, gt = synth.depth_pair(shape, 81, border=2)
    gt[0, 0, 10, 8:20] = torch.linspace(1e-4, 0.0099, 12)
    gt[0, 0, 11, 8] = 0.01
    gt[0, 0, 12, 9] = 0.010001
    pred[1, 0, 21, 31] = 3e-8                          # below the metrics clamp, a legal SILog operand
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double())
    v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), names)]
    (lp, gp), (lf, gf, vf) = _silog_both(Cr, pred, gt, names)
    close(lp, l64, LOSS_RTOL); grad_close(gp, g64)
    close(lf, l64, LOSS_RTOL); grad_close(gf, g64)
    close(vf, v64, 1e-5)
    # a subnormal (valid) target: masked for the loss, still a pixel of the metric suite ('rmse' would overflow
    # fp32 there in the reference as well, so it is left out of this list)
    names2 = ["delta1", "delta2", "delta3", "mae", "log10"]
    gt2 = gt.clone(); gt2[1, 0, 20, 30] = 1e-39
    l64b, g64b = olosses.loss_and_grad(olosses.silog, pred.double(), gt2.double())
    v64b = [float(v) for v in ometrics.compute(pred.double(), gt2.double(), names2)]
    (lp, gp), (lf, gf, vf) = _silog_both(Cr, pred, gt2, names2)
    close(lp, l64b, LOSS_RTOL); grad_close(gp, g64b)
    close(lf, l64b, LOSS_RTOL); grad_close(gf, g64b)
    close(vf, v64b, 1e-5)
    # non-positive predictions under the mask poison the loss exactly as the reference's log does
    for bad in (0.0, -1.0, float("nan")):
        p2 = pred.clone(); p2[0, 0, 30, 40] = bad
        assert gt[0, 0, 30, 40] > 0.01
        (lp, _), (lf, _, _) = _silog_both(Cr, p2, gt, names)
        assert torch.isnan(lp) and torch.isnan(lf), bad


@pytest.mark.parametrize("extra_tiles", [0, 1])
def test_silog_shared_memory_variant_capacity_boundary(Cr, extra_tiles):
    """The SS kernels hold at most 10 tiles of 2048 px per CTA: a batch of exactly grid x 10 tiles runs there,
    one tile more takes the generic kernel; both must agree with the oracle (and with each other)."""
    from mono_depth_estimation_b200 import _lib
    import ctypes as C
    sm, coop = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().mde_device_info(C.byref(sm), C.byref(coop)))
    tiles = coop.value * 10 + extra_tiles
    shape = (1, 1, tiles, 2048)
    pred, gt = synth.depth_pair(shape, 82, border=0)
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double())
    (lp, gp), (lf, gf, _) = _silog_both(Cr, pred, gt)
    close(lp, l64, LOSS_RTOL); grad_close(gp, g64)
    close(lf, l64, LOSS_RTOL); grad_close(gf, g64)


def test_silog_shared_memory_variant_is_bit_reproducible(Cr):
    """compute-sanitizer is not available on the GPU pool, so races in the SS kernels (bulk-copy ring, in-place
    residual slots, dynamic tile claims, the ticket-free all-reduce) are hunted by determinism: its loss totals are
    summed in a fixed order and every gradient element is a pure function of (p_i, t_i, totals), so 40 launches on
    the same C2 batch - interleaved with launches on other data that recycle the workspace parity sets and the
    slot words - must give bit-identical loss and gradient, for the plain and the fused kernel."""
    from mono_depth_estimation_b200 import metrics as M
    pred, gt = synth.depth_pair((16, 1, 480, 640), 91, device="cuda")
    other_p, other_g = synth.depth_pair((8, 1, 228, 304), 92, device="cuda")
    for fused in (False, True):
        ref_loss = ref_grad = ref_vals = None
        for it in range(40):
            mc = M.MetricComputation(["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"], strict=False)
            crit = Cr.silog_loss(0.85)
            if fused:
                crit = crit.fuse_metrics(mc)
            p = pred.detach().requires_grad_(True)
            loss = crit(p, gt)
            vals = torch.stack(mc.compute(p.detach(), gt)) if fused else None
            loss.backward()
            if ref_loss is None:
                ref_loss, ref_grad, ref_vals = loss.detach().clone(), p.grad.clone(), vals
            else:
                assert torch.equal(loss.detach(), ref_loss), (fused, it)
                assert torch.equal(p.grad, ref_grad), (fused, it)
                if fused:   # the metric sums go through fp64 atomics: equal to ~1e-15, the integer counts exactly
                    close(vals, ref_vals, 1e-12)
            if it % 3 == 0:
                run_loss(Cr.silog_loss(0.85), other_p, other_g)


@pytest.mark.parametrize("shape,dtype", [((2, 1, 37, 41), torch.float32), ((16, 1, 480, 640), torch.float32),
                                         ((3, 1, 64, 80), torch.float16), ((1, 1, 5, 7), torch.bfloat16)])
def test_upstream_gradient_scales_the_stashed_gradient(shape, dtype):
    """loss.backward() stashes dL/dpred at forward time; an upstream gradient other than 1 (a GradScaler, a weighted
    sum of losses) multiplies it in place (mde_scale_inplace: 128-bit path and the scalar tail)."""
    from mono_depth_estimation_b200 import criteria
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(shape, generator=g) * 9.5 + 0.5
    gt[torch.rand(shape, generator=g) < 0.2] = 0.0
    pred = (gt.clamp_min(0.5) + torch.randn(shape, generator=g) * 0.3).clamp_min(1e-3).to(dtype).cuda()
    gt = gt.cuda()
    p1 = pred.clone().requires_grad_(True)
    criteria.MaskedL1Loss()(p1, gt).backward()
    p2 = pred.clone().requires_grad_(True)
    (criteria.MaskedL1Loss()(p2, gt) * 2.5).backward()
    assert p2.grad.dtype == dtype
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(p2.grad.float(), p1.grad.float() * 2.5, rtol=tol, atol=0.0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", ["silog", "l1", "mse", "berhu"])
def test_amp_grad_scaler_at_training_size(Cr, name, dtype):
    """Half-precision predictions under GradScaler (reference train.py:60 --precision 16, :139) at a real batch size:
    dloss/dpred is of order 1/N (5e-7 at 8x416x544) - below fp16's normal range - until the scale (65536) is applied.
    The stash is fp32, the scale is applied in fp32 and the result rounded ONCE to the prediction dtype, so the scaled
    gradient keeps the dtype's relative precision (a half-precision stash flushes most of it to zero)."""
    shape = (8, 1, 416, 544)                           # 1.81 M px
    pred, gt = synth.depth_pair(shape, 36)
    ph = pred.to(dtype)
    _, g64 = olosses.loss_and_grad(olosses.LOSSES[name], ph.double(), gt.double())
    p = ph.cuda().requires_grad_(True)
    loss = make(Cr, name)(p, gt.cuda())
    (loss * 65536.0).backward()
    assert p.grad.dtype == dtype
    ref = g64 * 65536.0
    got = p.grad.double().cpu()
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7     # one rounding to the dtype (+ margin)
    big = ref.abs() > 1e-3 * ref.abs().max()
    rel = ((got - ref).abs() / ref.abs().clamp_min(1e-30))[big]
    assert float(rel.max()) < eps, (name, float(rel.max()))
    assert float((got != 0).double().mean()) > 0.5 * float((ref != 0).double().mean()), "gradient flushed to zero"


def test_fused_metrics_offer_is_not_served_to_other_data(Cr):
    """The hand-over from a fusing criterion is keyed on the LIVE tensor objects: once the offered prediction is freed,
    a new tensor that the caching allocator places at the same address (same shape, version 0) must be evaluated on its
    own data, not served the previous batch's metrics."""
    from mono_depth_estimation_b200 import metrics as M
    from oracle import metrics as ometrics
    names = ["delta1", "mse", "mae", "log10", "rmse"]
    shape = (2, 1, 64, 96)
    mc = M.MetricComputation(names, strict=False)
    crit = Cr.silog_loss(0.85).fuse_metrics(mc)
    pred, gt = synth.depth_pair(shape, 91, border=2)
    g = gt.cuda()
    p = pred.cuda().requires_grad_(True)
    addr = p.data_ptr()
    crit(p, g).backward()                              # offer pending; the step never calls compute() (e.g. logs every k steps)
    del p
    other, _ = synth.depth_pair(shape, 92, border=2)
    q = other.cuda()                                   # typically lands on the freed block
    vals = mc.compute(q, g)
    want = [float(v) for v in ometrics.compute(other.double(), gt.double(), names)]
    close(torch.stack(vals), want, 1e-5, msg="same address: %s" % (q.data_ptr() == addr))
    # an in-place update of the offered tensor also invalidates the offer
    p = pred.cuda().requires_grad_(True)
    crit(p, g).backward()
    with torch.no_grad():
        p.mul_(1.5)
    vals = mc.compute(p.detach(), g)
    want = [float(v) for v in ometrics.compute(pred.double() * 1.5, gt.double(), names)]
    close(torch.stack(vals), want, 1e-5)


def test_fused_metrics_booked_in_the_launch(Cr):
    """fuse_metrics(mc, book=True): the criterion's launch adds the values to the running sums and counts the call;
    the following compute() on the same tensors only reads (no launch at all), avg() matches the reference recipe."""
    from mono_depth_estimation_b200 import metrics as M, _lib
    from oracle import metrics as ometrics
    names = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]
    mc = M.MetricComputation(names, strict=False)
    crit = Cr.silog_loss(0.85).fuse_metrics(mc, book=True)
    acc = np.zeros(len(names))
    for k, seed in enumerate((93, 94, 95)):
        pred, gt = synth.depth_pair((2, 1, 48, 64), seed, border=2)
        p, g = pred.cuda().requires_grad_(True), gt.cuda()
        loss = crit(p, g)
        n0 = _lib.launch_count()
        vals = mc.compute(p.detach(), g)
        assert _lib.launch_count() == n0 and mc.count == k + 1
        loss.backward()
        want = np.array([float(v) for v in ometrics.compute(pred.double(), gt.double(), names)])
        close(torch.stack(vals), want, 1e-5)
        acc += want
    close(torch.stack([mc.avg(n) for n in names]), acc / 3, 1e-5)


def test_plain_backward_needs_no_scaling_launch(Cr):
    """`loss.backward()` on the criterion's result takes the stashed gradient as it is (the root gradient is autograd's
    implicit ones, known on the host); a scaled loss goes through mde_scale_inplace and gives the scaled gradient."""
    from mono_depth_estimation_b200 import _lib
    pred, gt = synth.depth_pair((2, 1, 48, 64), 11, border=2)
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double())
    crit = Cr.silog_loss(0.85)
    p = pred.cuda().requires_grad_(True)
    _lib.workspace(p.device, 2)
    n0 = _lib.launch_count()
    crit(p, gt.cuda()).backward()
    assert _lib.launch_count() - n0 == 1
    grad_close(p.grad, g64)
    q = pred.cuda().requires_grad_(True)
    n0 = _lib.launch_count()
    (crit(q, gt.cuda()) * 4.0).backward()
    assert _lib.launch_count() - n0 == 2
    grad_close(q.grad, 4.0 * g64)
