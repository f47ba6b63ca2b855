"""CPU-only checks of the C-ABI boundary and the host-side logic (no kernel is launched)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mde_b200.h")


@pytest.fixture(scope="module")
def lib():
    from mono_depth_estimation_b200 import build, _lib
    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mde_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mono_depth_estimation_b200 import _lib
    decl = _declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), "libmde_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert sorted(_lib.SIGNATURES) == decl, "ctypes table and header disagree"


def test_header_enums_match_python(lib):
    from mono_depth_estimation_b200 import _lib
    txt = open(HEADER).read()
    for name, val in (("MDE_METRIC_NQ", _lib.METRIC_NQ), ("MDE_METRIC_NM", _lib.METRIC_NM)):
        assert int(re.search(name + r"\s*=\s*(\d+)", txt).group(1)) == val
    for name, idx in _lib.METRIC_INDEX.items():
        assert int(re.search(r"MDE_M_" + name.upper() + r"\s*=\s*(\d+)", txt).group(1)) == idx
    for k, code in (("L1", 0), ("MSE", 1), ("BERHU", 2), ("LAINA_BERHU", 3), ("SILOG", 4), ("EIGEN", 5)):
        assert int(re.search(r"MDE_LOSS_" + k + r"\s*=\s*(\d+)", txt).group(1)) == code


def test_version_error_and_sizes(lib):
    assert b"sm_100a" in lib.mde_version()
    assert lib.mde_workspace_bytes(1) >= 1536
    assert lib.mde_workspace_bytes(654) >= 1536 + 3 * 654 * 16 * 8
    assert lib.mde_vnl_scratch_bytes(8, 100000, 385, 385) >= 8 * 100000 * 4 * 4 + 8 * 385 * 385 * 8
    # argument validation happens before any CUDA call
    rc = lib.mde_metrics(None, 0, None, 1, 1, 0, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.mde_last_error()
    rc = lib.mde_masked_loss(99, None, 0, None, None, 1, 1, 1, None, 1.0, None, None, None, None, None)
    assert rc == -1


def test_metrics_finalize_host(lib):
    raw = (C.c_double * 12)(*[100, 80, 90, 95, 50.0, 40.0, 3.0, 2.0, 10.0, 8.0, 20.0, 4.0])
    val = (C.c_double * 12)()
    lib.mde_metrics_finalize_host(raw, val)
    v = list(val)
    np.testing.assert_allclose(v[:3], [0.8, 0.9, 0.95])
    np.testing.assert_allclose(v[3:10], [0.5, 0.4, 0.03, 0.02, 0.1, 0.08, 0.2])
    np.testing.assert_allclose(v[10:], [np.sqrt(0.4), np.sqrt(0.04)])


def test_cpu_tensors_are_refused():
    """No CPU fallback: the product path raises for CPU tensors instead of computing."""
    from mono_depth_estimation_b200 import criteria, metrics, dorn
    p, t = torch.rand(1, 1, 4, 4) + 0.5, torch.rand(1, 1, 4, 4) + 0.5
    with pytest.raises(RuntimeError, match="CUDA"):
        criteria.MaskedL1Loss()(p, t)
    with pytest.raises(RuntimeError, match="CUDA"):
        metrics.MetricComputation(["mae"]).compute(p, t)
    with pytest.raises(RuntimeError, match="CUDA"):
        dorn.OrdinalRegressionLayer()(torch.randn(1, 4, 2, 2))


def test_reference_error_contract_on_host():
    from mono_depth_estimation_b200 import criteria, metrics
    with pytest.raises(AssertionError, match="inconsistent dimensions"):
        criteria.berHuLoss()(torch.ones(2, 3, 4), torch.ones(2, 1, 3, 4))
    with pytest.raises(KeyError):
        metrics.MetricComputation(["rmsle"])          # reference test.py:71 names a key METRICS lacks
    mc = metrics.MetricComputation(["delta1", "mae"])
    assert mc.names == ["delta1", "mae"] and mc.count == 0 and mc.sum == [0.0, 0.0]
    for cls in (criteria.MaskedDepthLoss, criteria.MaskedMSELoss, criteria.MaskedL1Loss, criteria.berHuLoss,
                criteria.LainaBerHuLoss):
        assert len(cls().state_dict()) == 0           # checkpoints of the reference load strictly
    assert len(criteria.silog_loss(0.85).state_dict()) == 0
    assert len(criteria.VNL_Loss(519., 519., (385, 385)).state_dict()) == 0


def test_vnl_select_index_contract():
    from mono_depth_estimation_b200 import criteria
    v = criteria.VNL_Loss(519.0, 519.0, (48, 64))
    d = v.select_index()
    assert sorted(d) == ["p1_x", "p1_y", "p2_x", "p2_y", "p3_x", "p3_y"]
    n = int(48 * 64 * 0.15)
    for k, arr in d.items():
        assert len(arr) == n and arr.min() >= 0
        assert arr.max() < (64 if k.endswith("x") else 48)


def test_metric_logger_keys():
    from mono_depth_estimation_b200 import metrics

    class Ctx:
        def __init__(self):
            self.calls = []

        def log(self, name, value, **kw):
            self.calls.append((name, kw))

    ml = metrics.MetricLogger(["delta1", "mae"], Ctx())
    ml.computer.compute = lambda p, t: [1.0, 2.0]
    ml.computer.count = 1
    ml.computer.sum = [1.0, 2.0]
    r = ml.log_train(None, None, 0.5)
    assert r == {"loss": 0.5, "delta1": 1.0, "mae": 2.0}
    assert [c[0] for c in ml.context.calls] == ["loss", "train_delta1", "train_delta1(AVG)", "train_mae", "train_mae(AVG)"]
    ml.context.calls.clear()
    assert ml.log_val(None, None, prefix="front_") == {"front_delta1": 1.0, "front_mae": 2.0}
    assert [c[0] for c in ml.context.calls] == ["val_front_delta1", "val_front_delta1(AVG)", "val_front_mae", "val_front_mae(AVG)"]
    ml.context.calls.clear()
    assert ml.log_test(None, None) == {"delta1": 1.0, "mae": 2.0}
    assert ml.context.calls == [("delta1", {"on_step": True, "on_epoch": True}), ("mae", {"on_step": True, "on_epoch": True})]


def test_decode_rule_matches_softmax_on_cpu():
    """The kernel decides P > 0.5 as fl(b'-a') > 1.5*2^-24 (csrc/dorn.cu header). Check that closed form
    against torch's own softmax on clamp ties, 1..5-ulp near ties at many magnitudes and random pairs."""
    from oracle import dorn as odorn
    g = torch.Generator().manual_seed(3)
    M = 400_000
    base = torch.rand(M, generator=g) * 8 + 1e-3
    base[: M // 8] = torch.rand(M // 8, generator=g) * 0.5          # more samples where ulp is small
    ulps = torch.randint(0, 6, (M,), generator=g)
    b = base.clone()
    for k in range(1, 6):
        b = torch.where(ulps >= k, torch.nextafter(b, torch.tensor(float("inf"))), b)
    swap = torch.rand(M, generator=g) < 0.5
    A, B = torch.where(swap, b, base), torch.where(swap, base, b)
    rnd = torch.randn(2, M, generator=g) * 2
    A, B = torch.cat([A, rnd[0]]), torch.cat([B, rnd[1]])
    x = torch.stack([A, B], 0).view(1, 2, 1, -1)
    decode, P = odorn.ordinal_layer(x)
    ac, bc = torch.clamp(A, 1e-8, 1e4), torch.clamp(B, 1e-8, 1e4)
    rule = (bc - ac) > 8.940696716308594e-08
    assert torch.equal(rule.view(-1), (P > 0.5).view(-1))
    assert int(decode.sum()) == int(rule.sum())


def test_layered_depth_term_selection_on_host():
    """The substring tests of BaseModule.setup_criterion (reference modules/base_module.py:156-194) and what is out of
    scope (SSIM / compositing terms), decided on the host before any launch."""
    import types
    from mono_depth_estimation_b200 import stdepth
    F = stdepth.TERM_FLAGS
    assert stdepth._flags_of("silma") == F["depth_silog"] | F["color_mae"]              # bts default (modules/bts.py:237)
    assert stdepth._flags_of("silms") == F["depth_silog"] | F["color_mse"]
    assert stdepth._flags_of("mse") == F["all_mse"] and stdepth._flags_of("mae") == F["all_mae"]
    assert stdepth._flags_of("mae+mse+fbdivergence") == F["all_mae"] | F["all_mse"] | F["fb_divergence"]
    assert stdepth._flags_of("silma+fbdivergence") == F["depth_silog"] | F["color_mae"] | F["fb_divergence"]
    for name in ("mae+composite", "silma+colorssim", "allssim", "composite+ssim"):      # laina's default is the first
        with pytest.raises(NotImplementedError):
            stdepth._flags_of(name)
    with pytest.raises(RuntimeError):
        stdepth._flags_of("dorn")                                                        # selects no term
    m = types.SimpleNamespace(loss="silma", variance_focus=0.85, depth_loss_weight=1.0, comp_loss_weight=1.0,
                              fbdiv_loss_weight=1.0, ssim_loss_weight=1.0)
    crit = stdepth.setup_criterion(m, single_layer=True)
    with pytest.raises(RuntimeError):                                                    # CPU tensors: no fallback
        crit(torch.zeros(1, 10, 4, 4), torch.zeros(1, 10, 4, 4), torch.zeros(1, 4, 4, 4))


def test_midas_family_host_contract():
    from mono_depth_estimation_b200 import criteria
    for cls in (criteria.MidasLoss, criteria.TrimmedProcrustesLoss):
        with pytest.raises(NotImplementedError):
            cls(reduction="image-based")
        assert len(cls().state_dict()) == 0                                              # checkpoints load strictly
    with pytest.raises(ValueError):
        criteria.MidasLoss(loss="huber")                                                 # criteria.py:316-317
    assert criteria.TrimmedProcrustesLoss().prediction_ssi is None                       # criteria.py:343
    with pytest.raises(RuntimeError):
        criteria.TrimmedProcrustesLoss()(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 4, 4))
