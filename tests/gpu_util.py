"""Shared helpers of the -m gpu parity tests."""
import numpy as np
import torch

T = torch.from_numpy


def close(a, b, rtol, atol=0.0, msg=""):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


# Tolerances of BASELINE.json: losses, gradients and float metrics within 1e-5 relative (fp32).
LOSS_RTOL = 1e-5


def grad_close(g, g_ref, msg=""):
    """Gradient check: every element within 1e-5 relative, with an absolute floor of 2e-6 of the largest
    |gradient| (elements that are differences of nearly equal terms carry no relative precision in the
    fp32 reference either)."""
    g_ref = g_ref.detach().double().cpu() if torch.is_tensor(g_ref) else torch.from_numpy(np.asarray(g_ref, dtype=np.float64))
    scale = float(g_ref.abs().max())
    close(g, g_ref, rtol=1e-5, atol=2e-6 * scale, msg=msg)


def run_loss(module, pred, *args, **kw):
    p = pred.detach().clone().requires_grad_(True)
    loss = module(p, *args, **kw)
    loss.backward()
    return loss.detach(), p.grad.detach()
