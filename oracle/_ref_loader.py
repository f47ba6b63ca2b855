"""Load the upstream reference modules BY FILE PATH (test infrastructure only).

Used by `oracle/gen_golden.py`, `tests/test_oracle_vs_reference.py` (container with `/root/reference`)
and `oracle/ref_step.py` (the CPU baseline of bench.py, which on the GPU box finds the unmodified copy
under `baseline/_ref/`). Nothing in the product package imports it.

Shims:
  * never put the reference directory on sys.path (its statistics.py shadows the stdlib);
  * `torchmetrics` is not installed -> a stand-in module with the closed forms of
    torchmetrics 0.7.3 `mean_absolute_error`, `mean_squared_error`,
    `mean_squared_log_error` (requirements.txt:2, call sites metrics.py:116-119);
  * `np.int` was removed from numpy>=1.24 (criteria.py:924-930).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """The reference tree itself when present (this container), else the unmodified copy of the hot-path files that
    oracle/install_ref.py put under git-ignored baseline/_ref (the GPU box receives only the repository)."""
    for cand in (os.environ.get("MDE_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "criteria.py")):
            return cand
    return os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "criteria.py"))


def _install_torchmetrics_standin():
    if "torchmetrics" in sys.modules:
        return
    tm = types.ModuleType("torchmetrics")
    fn = types.ModuleType("torchmetrics.functional")
    reg = types.ModuleType("torchmetrics.functional.regression")

    def mean_absolute_error(preds, target):
        return torch.sum(torch.abs(preds - target)) / target.numel()

    def mean_squared_error(preds, target):
        d = preds - target
        return torch.sum(d * d) / target.numel()

    def mean_squared_log_error(preds, target):
        d = torch.log1p(preds) - torch.log1p(target)
        return torch.sum(d * d) / target.numel()

    def structural_similarity_index_measure(*a, **k):  # out of scope
        raise NotImplementedError("ssim is out of scope")

    reg.mean_absolute_error = mean_absolute_error
    reg.mean_squared_error = mean_squared_error
    reg.mean_squared_log_error = mean_squared_log_error
    fn.regression = reg
    fn.structural_similarity_index_measure = structural_similarity_index_measure
    tm.functional = fn
    sys.modules["torchmetrics"] = tm
    sys.modules["torchmetrics.functional"] = fn
    sys.modules["torchmetrics.functional.regression"] = reg


def _load(name, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location("_mde_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load(name):
    """name in {'criteria', 'metrics', 'dorn_net'}"""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(np, "int"):
        np.int = int  # criteria.py:924-930
    if name == "criteria":
        mod = _load("criteria", "criteria.py")
    elif name == "metrics":
        _install_torchmetrics_standin()
        mod = _load("metrics", "metrics.py")
    elif name == "dorn_net":
        mod = _load("dorn_net", os.path.join("network", "Dorn.py"))
    else:
        raise KeyError(name)
    _cache[name] = mod
    return mod
