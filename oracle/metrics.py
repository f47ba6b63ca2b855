"""ORACLE (test infrastructure, not product code): CPU restatement of the masked error metrics.

Follows reference metrics.py:58-67 (MetricComputation.compute) and the metric functions
metrics.py:75-109. `mae`/`mse`/`msle` are torchmetrics 0.7.3 functionals in the reference
(metrics.py:116-119, requirements.txt:2 -- a third-party dependency that is NOT under
/root/reference and not installed): their published closed forms sum(|p-t|)/numel,
sum((p-t)^2)/numel, sum((log1p p - log1p t)^2)/numel are restated here; the reference holds no
test that pins them, so for those three keys parity is "unpinned by the reference" and anchored
on the closed form. `ssim` is out of scope. `rmse_true` / `rmse_log` do not exist in the
reference: fp64 closed forms only.
"""
from __future__ import annotations

import torch

DELTA_THRESHOLDS = (1.25 ** 1, 1.25 ** 2, 1.25 ** 3)  # metrics.py:77,82,87 (python doubles, exact in fp32)


def _max_ratio(p, t):
    return torch.max(p / t, t / p)  # metrics.py:76,81,86


def delta_count(p, t, k):
    """Integer count behind Delta{k}_multi_gpu (metrics.py:75-87): the bit-exact artefact."""
    return int((_max_ratio(p, t) < DELTA_THRESHOLDS[k - 1]).sum().item())


METRIC_FNS = {
    "delta1": lambda p, t: (_max_ratio(p, t) < DELTA_THRESHOLDS[0]).float().mean(),
    "delta2": lambda p, t: (_max_ratio(p, t) < DELTA_THRESHOLDS[1]).float().mean(),
    "delta3": lambda p, t: (_max_ratio(p, t) < DELTA_THRESHOLDS[2]).float().mean(),
    "mae": lambda p, t: torch.sum(torch.abs(p - t)) / t.numel(),
    "mse": lambda p, t: torch.sum((p - t) * (p - t)) / t.numel(),
    "msle": lambda p, t: torch.sum((torch.log1p(p) - torch.log1p(t)) ** 2) / t.numel(),
    "log10": lambda p, t: (torch.log10(p) - torch.log10(t)).abs().mean(),      # metrics.py:90-91
    "absrel": lambda p, t: (torch.abs(p - t) / t).mean(),                      # metrics.py:94-97
    "sqrel": lambda p, t: ((p - t) ** 2 / t).mean(),                           # metrics.py:100-103
    "rmse": lambda p, t: torch.sqrt((p - t) ** 2 / t).mean(),                  # metrics.py:106-109 (quirk)
    # additions (no reference counterpart)
    "rmse_true": lambda p, t: torch.sqrt(torch.sum((p - t) * (p - t)) / t.numel()),
    "rmse_log": lambda p, t: torch.sqrt(torch.sum((torch.log(p) - torch.log(t)) ** 2) / t.numel()),
}


def gather_valid(pred, target):
    """metrics.py:59-63: clamp pred at 1e-7, keep target > 0; returns the two 1-D gathered vectors."""
    pred = torch.clamp_min(pred, 1e-07)
    valid = target > 0
    assert torch.sum(valid) > 0, "invalid target!"
    return pred[valid], target[valid]


def compute(pred, target, names):
    """One mean per metric over the valid pixels of the WHOLE call tensor (metrics.py:58-67)."""
    p, t = gather_valid(pred, target)
    return [METRIC_FNS[n](p, t) for n in names]


def delta_counts(pred, target):
    """(n_valid, c1, c2, c3) integers for the call tensor."""
    p, t = gather_valid(pred, target)
    return (int(p.numel()), delta_count(p, t, 1), delta_count(p, t, 2), delta_count(p, t, 3))


def compute_per_image_mean(pred, target, names):
    """Dataset value as the reference's eval loop produces it: the test loader has
    batch size 1 (modules/base_module.py:71-76), compute() gives a per-image mean, Lightning then
    averages over steps -> unweighted mean over images of per-image means."""
    B = pred.shape[0]
    acc = [0.0] * len(names)
    for b in range(B):
        vals = compute(pred[b:b + 1], target[b:b + 1], names)
        for i, v in enumerate(vals):
            acc[i] += float(v)
    return [a / B for a in acc]


class RunningMetrics:
    """metrics.py:47-72 (MetricComputation): running mean of per-call values."""

    def __init__(self, names):
        self.names = list(names)
        self.reset()

    def reset(self):
        self.count = 0
        self.sum = [0.0 for _ in self.names]

    def compute(self, pred, target):
        vals = compute(pred, target, self.names)
        self.count += 1
        for i, v in enumerate(vals):
            self.sum[i] += v
        return vals

    def avg(self, metric):
        if isinstance(metric, int):
            return self.sum[metric] / self.count
        if isinstance(metric, str):
            return self.sum[self.names.index(metric)] / self.count
        assert False, "metric must be int or str"


RAW_NAMES = ("n_valid", "d1", "d2", "d3", "abs", "sq", "log10", "sle", "absrel", "sqrel", "rsq", "lnsq")


def raw_sums(pred, target):
    """The 12 raw per-call sums of include/mde_b200.h (MDE_Q_*) in the input dtype: counts from the
    reference's fp32 ratio test when given fp32 tensors, float sums as plain sums of the reference terms."""
    p, t = gather_valid(pred, target)
    r = _max_ratio(p, t)
    d = p - t
    out = [p.numel(), int((r < DELTA_THRESHOLDS[0]).sum()), int((r < DELTA_THRESHOLDS[1]).sum()),
           int((r < DELTA_THRESHOLDS[2]).sum()), d.abs().sum(), (d * d).sum(),
           (torch.log10(p) - torch.log10(t)).abs().sum(), ((torch.log1p(p) - torch.log1p(t)) ** 2).sum(),
           (d.abs() / t).sum(), (d * d / t).sum(), torch.sqrt(d * d / t).sum(), ((torch.log(p) - torch.log(t)) ** 2).sum()]
    return torch.tensor([float(v) for v in out], dtype=torch.float64)
