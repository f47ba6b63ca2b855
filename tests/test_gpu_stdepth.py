"""GPU parity of the layered-depth base criterion (reference modules/base_module.py:124-208, the
criterion of the registered methods `bts` and `laina`) against the reference-made golden vectors and the oracle."""
import types

import numpy as np
import pytest
import torch

from oracle import stdepth as ost
from tests.gpu_util import LOSS_RTOL, T, close, grad_close
from tests.test_oracle_vs_golden import STDEPTH_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from mono_depth_estimation_b200 import stdepth
    return stdepth


def _method(loss_name, depth_w=0.7, fbdiv_w=0.3):
    return types.SimpleNamespace(loss=loss_name, variance_focus=0.85, depth_loss_weight=depth_w, comp_loss_weight=1.0,
                                 fbdiv_loss_weight=fbdiv_w, ssim_loss_weight=1.0)


def _run(crit, pred, targ, rgba, **kw):
    p = pred.detach().clone().requires_grad_(True)
    ret = crit(p, targ, rgba, **kw)
    ret[0].backward()
    return ret, p.grad.detach()


@pytest.mark.parametrize("name,loss_name,C", STDEPTH_CASES)
def test_golden(S, golden, name, loss_name, C):
    g = golden("stdepth_small.npz")
    pred, targ, rgba = T(g[f"pred{C}"]).cuda(), T(g[f"targ{C}"]).cuda(), T(g[f"rgba{C}"]).cuda()
    crit = S.setup_criterion(_method(loss_name), single_layer=(C == 10))
    ret, grad = _run(crit, pred, targ, rgba, return_loss_dict=True)
    assert len(ret) == 2 and ret[0].dim() == 0 and grad.shape == pred.shape
    close(ret[0], g[f"{name}_loss64"], LOSS_RTOL)
    grad_close(grad, g[f"{name}_grad64"])
    for k in S.TERM_FLAGS:                                               # the same dict keys as the reference's loss_dict
        assert (k in ret[1]) == (f"{name}_{k}64" in g.files), k
    for k, v in ret[1].items():
        close(v, g[f"{name}_{k}64"], LOSS_RTOL)
        assert not v.requires_grad
    with torch.no_grad():
        (l,) = crit(pred, targ, rgba)
        close(l, g[f"{name}_loss64"], LOSS_RTOL)


def test_empty_depth_mask_and_errors(S, golden):
    g = golden("stdepth_small.npz")
    pred, targ, rgba = T(g["e_pred"]).cuda(), T(g["e_targ"]).cuda(), T(g["e_rgba"]).cuda()
    ret, grad = _run(S.setup_criterion(_method("silma")), pred, targ, rgba, return_loss_dict=True)
    assert float(ret[1]["depth_silog"]) == 0.0                          # NaN -> nan_to_num -> 0 (base_module.py:126-127)
    close(ret[0], g["e_loss32"], LOSS_RTOL)
    assert float(grad[:, 8:].abs().max()) == 0.0 and bool(torch.isfinite(grad).all())
    with pytest.raises(NotImplementedError):
        S.setup_criterion(_method("mae+composite"))                     # compositing term: stdepth_utils, out of scope
    with pytest.raises(NotImplementedError):
        S.setup_criterion(_method("silma+colorssim"))
    with pytest.raises(ValueError):
        S.setup_criterion(_method("silma"), single_layer=False)(pred, targ, rgba)
    with pytest.raises(NotImplementedError):
        S.setup_criterion(_method("silma"))(pred, targ, rgba, return_composited=True)
    # compositing for the visualisation is the caller's function (the reference's stdepth_utils.composite_layers)
    calls = []
    def comp(layers):
        calls.append(tuple(layers.shape))
        return layers[:, 0, :4]
    ret = S.setup_criterion(_method("silma"), composite_layers=comp)(pred, targ, rgba, return_composited=True)
    assert len(ret) == 2 and calls == [(2, 2, 4, 8, 9)] and ret[1].shape == (2, 4, 8, 9)


@pytest.mark.parametrize("shape,loss_name,dtype", [((4, 10, 256, 256), "silma", torch.float32),
                                                   ((2, 10, 97, 131), "silms+fbdivergence", torch.float32),
                                                   ((2, 20, 128, 160), "mae+mse+fbdivergence", torch.float32),
                                                   ((3, 10, 64, 80), "silma", torch.float16),
                                                   ((2, 20, 64, 96), "silma", torch.float32),       # quad kernel, three layers
                                                   ((2, 10, 64, 96), "mae+mse", torch.float32),     # quad kernel, all-channel terms
                                                   ((2, 20, 40, 52), "silms", torch.float32),
                                                   ((2, 10, 64, 96), "silms+fbdivergence", torch.float32)])   # quad kernel with the front/back term
def test_vs_oracle(S, shape, loss_name, dtype):
    """BTS's default 'silma' at a training-like size, odd sizes, three layers, AMP (fp16 prediction)."""
    from mono_depth_estimation_b200.synth import stdepth_inputs
    B, C, H, W = shape
    pred, targ, rgba = stdepth_inputs(31 + B + C, B, C, H, W)
    pred = pred.to(dtype)
    p64 = pred.double().requires_grad_(True)
    l64, d64 = ost.stdepth_loss(p64, targ.double(), rgba.double(), loss_name, variance_focus=0.85, depth_w=0.7, fbdiv_w=0.3,
                                single_layer=(C == 10))
    (g64,) = torch.autograd.grad(l64, p64)
    crit = S.setup_criterion(_method(loss_name), single_layer=(C == 10))
    ret, grad = _run(crit, pred.cuda(), targ.cuda(), rgba.cuda(), return_loss_dict=True)
    close(ret[0], l64.detach(), LOSS_RTOL, msg=loss_name)
    for k, v in ret[1].items():
        close(v, d64[k].detach(), LOSS_RTOL, msg=k)
    if dtype == torch.float32:
        grad_close(grad, g64, msg=loss_name)
    else:
        assert grad.dtype == dtype
        close(grad, g64, 2e-3, 2e-3 * float(g64.abs().max()), msg=loss_name)
    # second call on the same workspace (parity sets of the cooperative accumulators)
    ret2, grad2 = _run(crit, pred.cuda(), targ.cuda(), rgba.cuda())
    close(ret2[0], ret[0], 1e-6)
