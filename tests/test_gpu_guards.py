"""Out-of-bounds hunt without compute-sanitizer: the C ABI is called on buffers that sit inside larger allocations
whose guard bands hold NaN (inputs: a stray read that is USED poisons the result) or a sentinel (outputs: a stray
write is seen). Covers the kernels with neighbour / multi-row addressing: MidasLoss (quad and scalar paths),
the robust statistics, the layered-depth criterion."""
import pytest
import torch

from tests.gpu_util import close

pytestmark = pytest.mark.gpu

GUARD = 8192          # elements either side, a multiple of 4: the payload keeps its 16-byte alignment
SENTINEL = 12345.0


def guarded(t, fill):
    t = t.contiguous()
    buf = torch.full((t.numel() + 2 * GUARD,), fill, dtype=t.dtype, device="cuda")
    view = buf[GUARD:GUARD + t.numel()].view(t.shape)
    view.copy_(t)
    return buf, view


def guards_intact(buf, n, fill):
    lo, hi = buf[:GUARD], buf[GUARD + n:]
    if fill != fill:
        return bool(torch.isnan(lo).all()) and bool(torch.isnan(hi).all())
    return bool((lo == fill).all()) and bool((hi == fill).all())


def _inputs(shape, seed):
    g = torch.Generator().manual_seed(seed)
    target = torch.rand(shape, generator=g) * 9.5 + 0.5
    target[torch.rand(shape, generator=g) < 0.25] = 0.0
    pred = target.clamp_min(0.4) + torch.randn(shape, generator=g) * 0.3
    return pred, target


@pytest.mark.parametrize("shape", [(2, 37, 64), (3, 40, 43), (1, 8, 8), (2, 5, 4), (1, 64, 132)])
@pytest.mark.parametrize("scales", [4, 6])
def test_midas_loss_stays_inside_its_buffers(shape, scales):
    from mono_depth_estimation_b200 import _lib, criteria
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    pred, target = _inputs(shape, 5 + shape[1])
    B, H, W = shape
    ref = criteria.MidasLoss(alpha=0.5, loss="l1", scales=scales)
    p = pred.cuda().requires_grad_(True)
    l_ref = ref(p, target.cuda())
    l_ref.backward()
    pb, pv = guarded(pred.cuda(), float("nan"))
    tb, tv = guarded(target.cuda(), float("nan"))
    gb, gv = guarded(torch.zeros(shape, device=dev), SENTINEL)
    loss = torch.empty((), device=dev)
    ws = _lib.workspace(dev, B)
    _lib.check(lib.mde_midas_loss(_lib.ptr(pv), 0, _lib.ptr(tv), None, None, B, H, W, 1, 0.5, scales, 1.0, _lib.ptr(ws),
                                  _lib.ptr(loss), _lib.ptr(gv), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert guards_intact(gb, pred.numel(), SENTINEL), "gradient written outside its buffer"
    assert guards_intact(pb, pred.numel(), float("nan")) and guards_intact(tb, pred.numel(), float("nan"))
    close(loss, l_ref.detach(), 1e-6)
    assert torch.equal(gv, p.grad), "a guard value leaked into the gradient"


@pytest.mark.parametrize("shape", [(3, 24, 32), (2, 33, 41), (5, 7, 9)])
def test_robust_normalize_stays_inside_its_buffers(shape):
    from mono_depth_estimation_b200 import _lib
    from oracle import midas as om
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    pred, target = _inputs(shape, 11 + shape[1])
    B, H, W = shape
    pb, pv = guarded(pred.cuda(), float("nan"))
    tb, tv = guarded(target.cuda(), float("nan"))
    ob, ov = guarded(torch.zeros(shape, device=dev), SENTINEL)
    qb, qv = guarded(torch.zeros(shape, device=dev), SENTINEL)
    sb, sv = guarded(torch.zeros((B, 8), device=dev), SENTINEL)
    ub, uv = guarded(torch.zeros((B, 8), device=dev), SENTINEL)
    nscr = int(lib.mde_robust_scratch_bytes(B)) // 8
    scb, scv = guarded(torch.zeros(nscr, dtype=torch.float64, device=dev), SENTINEL)
    _lib.check(lib.mde_robust_normalize(_lib.ptr(pv), _lib.ptr(tv), B, H * W, _lib.ptr(scv), _lib.ptr(sv), _lib.ptr(uv),
                                        _lib.ptr(ov), _lib.ptr(qv), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    for buf, n in ((ob, pred.numel()), (qb, pred.numel()), (sb, B * 8), (ub, B * 8), (scb, nscr)):
        assert guards_intact(buf, n, SENTINEL), "write outside an output buffer"
    mask = (target > 0).double()
    close(ov, om.normalize_prediction_robust(pred.double(), mask), 1e-5, 1e-6)
    close(qv, om.normalize_prediction_robust(target.double(), mask), 1e-5, 1e-6)


@pytest.mark.parametrize("shape", [(2, 10, 19, 27), (1, 20, 8, 12), (3, 10, 7, 33)])
def test_stdepth_stays_inside_its_buffers(shape):
    from mono_depth_estimation_b200 import _lib
    from oracle import stdepth as ost
    from mono_depth_estimation_b200.synth import stdepth_inputs
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    B, C, H, W = shape
    pred, targ, rgba = stdepth_inputs(77 + C, B, C, H, W)
    pb, pv = guarded(pred.cuda(), float("nan"))
    tb, tv = guarded(targ.cuda(), float("nan"))
    xb, xv = guarded(rgba.cuda(), float("nan"))
    gb, gv = guarded(torch.zeros(shape, device=dev), SENTINEL)
    ob, ov = guarded(torch.zeros(8, device=dev), SENTINEL)
    ws = _lib.workspace(dev, B)
    flags = 1 | 2 | 8 | 32      # depth_silog + color_mae + all_mse + fb_divergence
    _lib.check(lib.mde_stdepth_loss(_lib.ptr(pv), 0, _lib.ptr(tv), _lib.ptr(xv), 4, B, C, H * W, flags, 0.7, 0.3, 0.85, 1.0,
                                    _lib.ptr(ws), _lib.ptr(ov), _lib.ptr(gv), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert guards_intact(gb, pred.numel(), SENTINEL) and guards_intact(ob, 8, SENTINEL), "write outside an output buffer"
    p64 = pred.double().requires_grad_(True)
    l64, _ = ost.stdepth_loss(p64, targ.double(), rgba.double(), "silma+mse+fbdivergence", variance_focus=0.85, depth_w=0.7,
                              fbdiv_w=0.3, single_layer=(C == 10))
    (g64,) = torch.autograd.grad(l64, p64)
    close(ov[0], l64.detach(), 1e-5)
    close(gv, g64, 1e-5, 2e-6 * float(g64.abs().max()))
