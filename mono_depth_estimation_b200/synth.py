"""Deterministic synthetic NYU / Structured3D-shaped inputs for the five BASELINE.json configs.

The distributions follow SURVEY.md section 8(d): metric depth in U(0.5, 10) m (the
Structured3D/Floorplan loaders clip to [0, 10], reference datasets/structured3d_dataset.py:45),
a FIXED invalid-pixel pattern (20 % seeded Bernoulli holes plus a 10-px zero frame, as NYU's
white border) that encodes the mask as gt == 0 (the reference never takes a separate mask
tensor, criteria.py:25,73,86,121), and a dense prediction = gt + N(0, 0.5^2).

Everything is generated with an explicit torch.Generator so that the same seed gives the same
tensors in the golden-vector generator, the parity tests and the benchmark.
"""
from __future__ import annotations

import numpy as np
import torch

SEEDS = {"C1": 101, "C2": 102, "C3": 103, "C4": 104, "C5": 105}

SHAPES = {
    "C1": (8, 1, 228, 304),     # FCRN/Laina NYU training crop (datasets/nyu_dataloader.py:96)
    "C2": (16, 1, 480, 640),    # BTS batch at NYU full resolution
    "C3": (8, 136, 257, 353),   # DORN logits, K=68 (modules/dorn.py:208-215)
    "C4": (8, 1, 385, 385),     # VNL crop (modules/vnl.py:342-349)
    "C5": (654, 1, 480, 640),   # NYU test split (datasets/nyu_dataloader.py:146,277)
}

DEFAULT_EVAL_METRICS = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse",
                        "absrel", "sqrel", "msle"]


def _gen(seed: int, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def depth_pair(shape, seed, device="cpu", border=10, hole_frac=0.2, lo=0.5, hi=10.0,
               noise=0.5, dtype=torch.float32):
    """(pred, gt) of `shape` [B,1,H,W]: gt has zeros at invalid pixels, pred is dense and >= 1e-3."""
    g = _gen(seed, device)
    B, C, H, W = shape
    gt_full = torch.rand(shape, generator=g, device=device, dtype=torch.float32) * (hi - lo) + lo
    holes = torch.rand(shape, generator=g, device=device, dtype=torch.float32) < hole_frac
    eps = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * noise
    pred = torch.clamp_min(gt_full + eps, 1e-3)
    gt = gt_full.masked_fill(holes, 0.0)
    b = min(border, H // 4, W // 4)
    if b > 0:
        gt[..., :b, :] = 0
        gt[..., H - b:, :] = 0
        gt[..., :, :b] = 0
        gt[..., :, W - b:] = 0
    return pred.to(dtype).contiguous(), gt.contiguous()


def dorn_inputs(shape, seed, device="cpu", alpha=0.001, beta=1.0, hole_frac=0.2):
    """(logits [N,2K,H,W] ~ N(0, 2^2), gt [N,1,H,W] ~ U(alpha, beta) with fixed zeros)."""
    g = _gen(seed, device)
    N, C2, H, W = shape
    logits = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * 2.0
    gt = torch.rand((N, 1, H, W), generator=g, device=device, dtype=torch.float32) * (beta - alpha) + alpha
    holes = torch.rand((N, 1, H, W), generator=g, device=device, dtype=torch.float32) < hole_frac
    gt = gt.masked_fill(holes, 0.0)
    return logits.contiguous(), gt.contiguous()


def vnl_inputs(shape, seed, n_triplets=100_000, device="cpu", pad_rows=40, zero_frac=1e-3):
    """(gt, pred, triplets int64 [3,n] of flat pixel indices) for the virtual-normal loss.

    gt ~ U(0.01, 1.1); on every second image the top `pad_rows` rows are -1 (the VNL padding
    value, reference modules/vnl.py:104,209-215); pred = clamp_min(|gt| + N(0,0.05^2), 1e-3)
    with `zero_frac` exact zeros injected (exercises the z==0 fix-up, criteria.py:1004).
    The triplets come from numpy RandomState(seed).randint so they do not depend on torch.
    """
    g = _gen(seed, device)
    B, C, H, W = shape
    gt = torch.rand(shape, generator=g, device=device, dtype=torch.float32) * 1.09 + 0.01
    eps = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * 0.05
    zeros = torch.rand(shape, generator=g, device=device, dtype=torch.float32) < zero_frac
    pr = min(pad_rows, H // 4)
    if pr > 0:
        gt[1::2, :, :pr, :] = -1.0
    pred = torch.clamp_min(gt.abs() + eps, 1e-3).masked_fill(zeros, 0.0)
    rs = np.random.RandomState(int(seed))
    trip = torch.from_numpy(rs.randint(0, H * W, size=(3, n_triplets)).astype(np.int64)).to(device)
    return gt.contiguous(), pred.contiguous(), trip


def config_inputs(name: str, device="cpu", batch=None):
    """Inputs of a BASELINE.json config ('C1'..'C5'); `batch` overrides the leading dim."""
    shape = list(SHAPES[name])
    if batch is not None:
        shape[0] = int(batch)
    shape = tuple(shape)
    seed = SEEDS[name]
    if name in ("C1", "C2", "C5"):
        return depth_pair(shape, seed, device)
    if name == "C3":
        return dorn_inputs(shape, seed, device)
    if name == "C4":
        return vnl_inputs(shape, seed, device=device)
    raise KeyError(name)


def stdepth_inputs(seed, B, C, H, W):
    """Layered-depth batch: 8 (or 16) colour/alpha channels in [0,1], depth channels in (0, 1] with holes, an alpha
    plane with holes; pred = targ + noise, positive on the depth channels."""
    g = torch.Generator().manual_seed(seed)
    targ = torch.rand((B, C, H, W), generator=g)
    d0 = 8 if C == 10 else 16
    targ[:, d0:] = targ[:, d0:] * 0.95 + 0.05
    targ[:, d0:][torch.rand((B, C - d0, H, W), generator=g) < 0.3] = 0.0
    targ[:, d0][torch.rand((B, H, W), generator=g) < 0.05] = 0.005          # in maskD, outside silog's own mask (> 1e-2)
    rgba = torch.rand((B, 4, H, W), generator=g)
    rgba[:, 3][torch.rand((B, H, W), generator=g) < 0.35] = 0.0
    pred = targ + torch.randn((B, C, H, W), generator=g) * 0.1
    pred[:, d0:] = pred[:, d0:].abs() + 0.02
    pred[0, 0, 2, 3:6] = targ[0, 0, 2, 3:6]                                 # exact ties: sign(0) = 0
    pred[0, :3, 5, 5] = 0.0                                                 # zero front vector: norm gradient 0
    return pred, targ, rgba
