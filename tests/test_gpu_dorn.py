"""GPU parity of the DORN ordinal head kernels against the pinned CPU oracle."""
import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import dorn as odorn
from tests.gpu_util import LOSS_RTOL, T, close, grad_close, run_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    from mono_depth_estimation_b200 import dorn
    return dorn


@pytest.fixture(scope="module")
def Cr():
    from mono_depth_estimation_b200 import criteria
    return criteria


def _gradx_ref(g):
    """fp64 reference gradient, except on knife-edge logits where the fp32 and fp64 evaluations of the
    reference are different functions: a logit equal to fp32(1e-8) passes clamp(min=1e-8) in fp32
    (x >= min) but not after widening to fp64 (fp32(1e-8) < 1e-8). The kernel follows fp32."""
    g32, g64 = g["ordloss_gradx32"].astype(np.float64), g["ordloss_gradx64"]
    knife = np.abs(g32 - g64) > 1e-4 * np.abs(g64).max()
    assert knife.sum() <= 2
    return np.where(knife, g32, g64)


def test_small_golden_layer_and_ordloss(D, Cr, golden):
    g = golden("dorn_small.npz")
    x, gt = T(g["logits"]).cuda(), T(g["gt"]).cuda()
    K = x.shape[1] // 2
    xr = x.clone().requires_grad_(True)
    decode, P = D.OrdinalRegressionLayer()(xr)
    assert decode.dtype == torch.int64 and tuple(decode.shape) == (x.shape[0], 1) + tuple(x.shape[2:])
    assert tuple(P.shape) == (x.shape[0], K) + tuple(x.shape[2:])
    assert np.array_equal(decode.cpu().numpy(), g["decode"])                  # bit-exact incl. ties / near ties
    close(P, g["P64"], 1e-5, 1e-7)
    alpha, beta = torch.tensor(0.001), torch.tensor(1.0)
    depth = D.label_to_depth(decode, alpha, beta, torch.tensor(K).int())
    y = D.depth_to_label(gt, alpha, beta, torch.tensor(K).int())
    close(depth, g["depth"], 1e-5)
    yref = torch.from_numpy(g["y_sid"])
    fin = torch.isfinite(yref)
    close(y.cpu()[fin], yref[fin], 1e-5, 1e-5)
    assert torch.equal(torch.isinf(y.cpu()), torch.isinf(yref))                # gt = 0 -> -inf label
    loss = Cr.ordLoss()(P, y)                                                   # the reference's two-module path
    loss.backward()
    close(loss, g["ordloss64"], LOSS_RTOL)
    grad_close(xr.grad, _gradx_ref(g))
    Pl = T(g["P"]).cuda().requires_grad_(True)
    lp = Cr.ordLoss()(Pl, T(g["y_sid"]).cuda())
    lp.backward()
    close(lp, g["ordloss32"], LOSS_RTOL)
    grad_close(Pl.grad, g["ordloss_gradP32"])


def test_small_golden_fused(D, golden):
    g = golden("dorn_small.npz")
    x, gt = T(g["logits"]).cuda(), T(g["gt"]).cuda()
    K = x.shape[1] // 2
    xr = x.clone().requires_grad_(True)
    loss, decode, depth, P = D.dorn_fused(xr, gt, K, 0.001, 1.0, "SID", want_prob=True)
    loss.backward()
    assert np.array_equal(decode.cpu().numpy(), g["decode"])
    close(depth, g["depth"], 1e-5)
    close(P, g["P64"], 1e-5, 1e-7)
    close(loss, g["ordloss64"], LOSS_RTOL)
    grad_close(xr.grad, _gradx_ref(g))
    head = D.DornOrdinalHead(K, 0.001, 1.0)
    l2, d2, dec2 = head(x, gt)                                                   # no grad requested
    close(l2, g["ordloss64"], LOSS_RTOL)
    assert torch.equal(dec2, decode)


@pytest.mark.parametrize("disc", ["SID", "UD"])
def test_ordinal_regression_loss(Cr, golden, disc):
    g = golden("dorn_small.npz")
    prob, gt = T(g["orl_prob"]).cuda(), T(g["orl_gt"]).cuda()
    K = prob.shape[1] // 2
    orl = Cr.OrdinalRegressionLoss(K, torch.tensor(0.001), torch.tensor(1.0), disc)
    loss, grad = run_loss(orl, prob, gt)
    close(loss, g[f"orl_{disc}_loss"], LOSS_RTOL)
    grad_close(grad, g[f"orl_{disc}_grad"])


@pytest.mark.parametrize("shape,K", [((2, 136, 64, 80), 68), ((3, 20, 33, 41), 10), ((1, 142, 17, 19), 71)])
def test_random_vs_oracle(D, shape, K):
    x, gt = synth.dorn_inputs(shape, 41)
    x[0, :, 0, 0] = 0.0                      # all-tie pixel
    x[0, 0::2, 0, 1] = 5.0; x[0, 1::2, 0, 1] = -5.0
    x[0, 0::2, 0, 2] = -5.0; x[0, 1::2, 0, 2] = 5.0      # all pairs counted -> decode == K
    dec_ref, P_ref = odorn.ordinal_layer(x)
    xd = x.double().clone().requires_grad_(True)
    _, P64 = odorn.ordinal_layer(xd)
    y64 = odorn.depth_to_label(gt.double(), 0.001, 1.0, K)
    l64 = odorn.ord_loss(P64, y64)
    (g64,) = torch.autograd.grad(l64, xd)
    xr = x.cuda().requires_grad_(True)
    loss, decode, depth, _ = D.dorn_fused(xr, gt.cuda(), K, 0.001, 1.0)
    loss.backward()
    assert torch.equal(decode.cpu(), dec_ref)                                    # bit-exact
    assert int(decode[0, 0, 0, 0]) == 0 and int(decode[0, 0, 0, 2]) == K
    close(depth, odorn.label_to_depth(dec_ref, 0.001, 1.0, K), 1e-5)
    close(loss, l64.detach(), LOSS_RTOL)
    grad_close(xr.grad, g64)
    dec2, P = D.OrdinalRegressionLayer()(x.cuda())
    assert torch.equal(dec2.cpu(), dec_ref)
    close(P, P64.detach(), 1e-5, 1e-7)


def test_decode_near_ties_bit_exact(D):
    """Adversarial: 1..5-ulp near ties at many magnitudes, clamp ties, clamp saturation."""
    g = torch.Generator().manual_seed(9)
    M = 300_000
    base = torch.rand(M, generator=g) * 8 + 1e-3
    base[: M // 4] = torch.rand(M // 4, generator=g) * 0.5
    ulps = torch.randint(0, 6, (M,), generator=g)
    b = base.clone()
    for k in range(1, 6):
        b = torch.where(ulps >= k, torch.nextafter(b, torch.tensor(float("inf"))), b)
    swap = torch.rand(M, generator=g) < 0.5
    A, B = torch.where(swap, b, base), torch.where(swap, base, b)
    extra = torch.tensor([[0.0, 0.0], [-1.0, -2.0], [1e-8, 2e-8], [2e4, 3e4], [9999.0, 2e4], [-5.0, 1e-8], [1e-9, 1.1e-8]])
    A, B = torch.cat([A, extra[:, 0]]), torch.cat([B, extra[:, 1]])
    x = torch.stack([A, B], 0).view(1, 2, 1, -1).contiguous()
    dec_ref, P_ref = odorn.ordinal_layer(x)
    dec, P = D.OrdinalRegressionLayer()(x.cuda())
    assert torch.equal(dec.cpu(), dec_ref)
    close(P, P_ref, 1e-6, 1e-7)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_logits(D, dtype):
    x, gt = synth.dorn_inputs((2, 24, 9, 16), 42)
    xh = x.to(dtype)
    dec_ref, _ = odorn.ordinal_layer(xh.float())
    xd = xh.double().clone().requires_grad_(True)
    _, P64 = odorn.ordinal_layer(xd)
    l64 = odorn.ord_loss(P64, odorn.depth_to_label(gt.double(), 0.001, 1.0, 12))
    (g64,) = torch.autograd.grad(l64, xd)
    xr = xh.cuda().requires_grad_(True)
    loss, decode, depth, _ = D.dorn_fused(xr, gt.cuda(), 12, 0.001, 1.0)
    loss.backward()
    assert torch.equal(decode.cpu(), dec_ref) and xr.grad.dtype == dtype
    close(loss, l64.detach(), LOSS_RTOL)
    eps = 1e-3 if dtype == torch.float16 else 8e-3
    close(xr.grad, g64, eps, eps * float(g64.abs().max()))


def test_sid_tables(D):
    lab = torch.arange(0, 69).view(1, 1, 3, 23)
    for ds in ("kitti", "nyu", "floorplan3d"):
        close(D.get_depth_sid(ds, lab.cuda()), odorn.sid_table_depth(ds, lab), 1e-5)
    d = torch.rand(1, 1, 8, 8) * 9 + 0.5
    a, b = D.get_labels_sid("nyu", d.cuda()).cpu(), odorn.sid_table_labels("nyu", d)
    assert a.dtype == torch.int32 and torch.equal(a, b)
    close(D.label_to_depth(lab.cuda().float(), 0.5, 10.0, 68, "UD"), odorn.label_to_depth(lab.float(), 0.5, 10.0, 68, "UD"), 1e-6)
    close(D.depth_to_label(d.cuda(), 0.5, 10.0, 68, "UD"), odorn.depth_to_label(d, 0.5, 10.0, 68, "UD"), 1e-5, 1e-5)


def test_config_c3_full_size_consistency(D, Cr):
    """C3 at full size (8 x 136 x 257 x 353 logits, 395 MB), size-independent properties: the fused step equals the
    layer followed by ordLoss (decode bit-exact, loss and gradient to fp32 rounding); the batch loss is the mean of the
    per-image losses; decode never exceeds K and counts exactly the pairs with b' - a' above the tie margin."""
    shape = synth.SHAPES["C3"]
    N, C2, H, W = shape
    K = C2 // 2
    x, gt = synth.dorn_inputs(shape, 103, device="cuda")
    xr = x.clone().requires_grad_(True)
    loss, decode, depth, _ = D.dorn_fused(xr, gt, K, 0.001, 1.0)
    loss.backward()
    dec2, P = D.OrdinalRegressionLayer()(x)
    assert torch.equal(dec2, decode) and decode.dtype == torch.int64
    assert int(decode.min()) >= 0 and int(decode.max()) <= K
    a, b = x[:, 0::2].clamp(1e-8, 1e4), x[:, 1::2].clamp(1e-8, 1e4)
    assert torch.equal(decode, ((b - a) > 1.5 * 2.0 ** -24).sum(1, keepdim=True))
    del a, b
    Pl = P.detach().requires_grad_(True)
    l2 = Cr.ordLoss()(Pl, D.depth_to_label(gt, 0.001, 1.0, K))
    close(l2, loss.detach(), 2e-6)
    per_image = [float(D.dorn_fused(x[i:i + 1], gt[i:i + 1], K, 0.001, 1.0)[0]) for i in range(N)]
    close(loss.detach(), sum(per_image) / N, 2e-6)
    assert bool(torch.isfinite(xr.grad).all())
    # channels of a pair receive opposite gradients wherever both logits are inside the clamp
    ga, gb = xr.grad[:, 0::2], xr.grad[:, 1::2]
    inside = (x[:, 0::2] > 1e-8) & (x[:, 0::2] < 1e4) & (x[:, 1::2] > 1e-8) & (x[:, 1::2] < 1e4)
    assert torch.equal(ga[inside], -gb[inside])


def test_sid_integer_labels_exact_outside_the_rounding_band(D):
    """get_labels_sid truncates an fp32 expression whose logarithm comes from the platform's logf (modules/dorn.py:43-71:
    CPU libm in the oracle, the CUDA math library here - both within 1 ulp, neither correctly rounded, and the
    reference's own CPU and GPU runs disagree the same way). The integers must be EQUAL wherever the exact label (fp64)
    is further than 4 fp32 ulps from an integer; inside that band - where the two fp32 evaluations may land on different
    sides - they differ by at most one. 2 M depths, plus depths constructed to sit ON the integer labels."""
    g = torch.Generator().manual_seed(5)
    for ds, (alpha, beta, K) in (("nyu", (0.02, 10.0, 68)), ("kitti", (0.001, 80.0, 71)), ("floorplan3d", (0.0552, 10.0, 68))):
        d = (torch.rand(1 << 21, generator=g, dtype=torch.float64) * (beta - alpha) + alpha).float()
        k = torch.arange(0, K + 1, dtype=torch.float64)
        on_border = (alpha * (beta / alpha) ** (k / K)).float()          # depths whose label is an integer up to rounding
        d[:on_border.numel()] = on_border
        got = D.get_labels_sid(ds, d.cuda()).cpu()
        ref = odorn.sid_table_labels(ds, d)
        exact = K * torch.log(d.double() / float(torch.tensor(alpha).float())) / torch.log(torch.tensor(beta).float().double() / torch.tensor(alpha).float().double())
        band = (exact - exact.round()).abs() <= 4 * 2.0 ** -23 * exact.abs().clamp_min(1.0)
        assert torch.equal(got[~band], ref[~band]), ds
        assert int((got - ref).abs().max()) <= 1, ds
        assert float(band.double().mean()) < 1e-3, ds                    # the band is tiny: this is an exactness test


@pytest.mark.parametrize("shape,K", [((2, 136, 40, 52), 68), ((2, 22, 33, 41), 11)])
def test_nan_labels_and_the_wide_index_path(D, Cr, shape, K):
    """A negative or NaN target has a NaN SID label: neither k <= y nor k > y holds (criteria.py:769-770), the pixel
    contributes no loss term and its logits get a zero gradient (the kernel runs such pixels through their own code
    path). Then the same call with 64-bit per-pixel pointers instead of 32-bit element indices (the path of tensors with
    >= 2^32 elements, MDE_DORN_NO_INDEX32=1): bit-identical outputs."""
    import os
    x, gt = synth.dorn_inputs(shape, 43)
    gt[0, 0, 1, :7] = -1.0
    gt[1, 0, 2, 3:9] = float("nan")
    gt[1, 0, -1, -1] = -0.5
    xd = x.double().clone().requires_grad_(True)
    dec_ref, P64 = odorn.ordinal_layer(xd)
    l64 = odorn.ord_loss(P64, odorn.depth_to_label(gt.double(), 0.001, 1.0, K))
    (g64,) = torch.autograd.grad(l64, xd)
    assert torch.isfinite(l64) and float(g64[0, :, 1, :7].abs().sum()) == 0.0

    def once():
        xr = x.cuda().requires_grad_(True)
        loss, decode, depth, P = D.dorn_fused(xr, gt.cuda(), K, 0.001, 1.0, want_prob=True)
        loss.backward()
        # the reference's own sequence: layer, depth_to_label, ordLoss, backward through both modules
        x2 = x.cuda().requires_grad_(True)
        dec2, P2 = D.OrdinalRegressionLayer()(x2)
        y2 = D.depth_to_label(gt.cuda(), torch.tensor(0.001), torch.tensor(1.0), torch.tensor(K).int())
        l2 = Cr.ordLoss()(P2, y2)
        l2.backward()
        return loss.detach(), decode, depth, P.detach(), xr.grad, l2.detach(), dec2, P2.detach(), x2.grad
    out = once()
    close(out[5], l64.detach(), LOSS_RTOL)
    assert torch.equal(out[6].cpu(), dec_ref)
    grad_close(out[8], g64)
    close(out[0], l64.detach(), LOSS_RTOL)
    assert torch.equal(out[1].cpu(), dec_ref)
    grad_close(out[4], g64)
    assert float(out[4][0, :, 1, :7].abs().sum()) == 0.0 and float(out[4][1, :, 2, 3:9].abs().sum()) == 0.0
    os.environ["MDE_DORN_NO_INDEX32"] = "1"
    try:
        wide = once()
    finally:
        del os.environ["MDE_DORN_NO_INDEX32"]
    assert all(torch.equal(a, b) for a, b in zip(out, wide))
    assert float(out[8][0, :, 1, :7].abs().sum()) == 0.0
