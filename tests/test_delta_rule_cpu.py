"""CPU proof-by-exhaustion-near-the-boundary of the divide-free delta rule of csrc/metric_math.cuh:

    RN(hi/lo) < T   <=>   fmaf(lo, T, -hi) > lo * 2^-24        for T in {1.25, 1.5625, 1.953125}

(reference metrics.py:75-87 counts max(p/t, t/p) < 1.25^k with fp32 divides). The fp32 fma is emulated
exactly in float64: lo*T has <= 31 significant bits and hi is within a factor 2 of it, so the float64
product and difference are exact and one rounding to float32 remains - the definition of fmaf."""
import numpy as np
import pytest

THR = (1.25, 1.5625, 1.953125)


def rule(hi, lo, T):
    s = (lo.astype(np.float64) * T - hi.astype(np.float64)).astype(np.float32)      # == fmaf(lo, T, -hi)
    e = (lo * np.float32(2.0 ** -24)).astype(np.float32)
    return s > e


def reference(hi, lo, T):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.maximum(hi / lo, lo / hi) < np.float32(T)          # fp32 divides, as torch does


@pytest.mark.parametrize("T", THR)
def test_rule_equals_fp32_division_near_thresholds(T):
    rng = np.random.RandomState(5)
    n = 400_000
    lo = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n)).astype(np.float32)
    # hi = the floats around lo*T: every neighbour within +-6 ulp of the rounded product
    base = (lo.astype(np.float64) * T).astype(np.float32)
    for k in range(-6, 7):
        hi = base.copy()
        for _ in range(abs(k)):
            hi = np.nextafter(hi, np.float32(np.inf if k > 0 else 0), dtype=np.float32)
        ok = hi >= lo
        assert np.array_equal(rule(hi[ok], lo[ok], T), reference(hi[ok], lo[ok], T)), (T, k)


def test_rule_far_from_thresholds_and_special_values():
    rng = np.random.RandomState(6)
    a = np.exp(rng.uniform(np.log(1e-7), np.log(1e4), 300_000)).astype(np.float32)
    b = np.exp(rng.uniform(np.log(1e-7), np.log(1e4), 300_000)).astype(np.float32)
    hi, lo = np.maximum(a, b), np.minimum(a, b)
    for T in THR:
        assert np.array_equal(rule(hi, lo, T), reference(hi, lo, T))
    # equal operands, exact-threshold products (strict '<'), inf
    one = np.array([1.0, 3.0, 0.5], np.float32)
    for T in THR:
        assert rule(one, one, T).all()
        assert not rule(one * np.float32(T), one, T).any()
        assert not rule(np.array([np.inf], np.float32), np.array([2.0], np.float32), T).any()
    # a NaN prediction makes hi NaN (max.NaN in the kernel): every comparison is false, as NaN < T is
    assert not rule(np.array([np.nan], np.float32), np.array([2.0], np.float32), 1.25).any()
