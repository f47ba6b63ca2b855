"""Identities the DORN kernels of csrc/dorn.cu rely on (session 4), checked on the CPU in fp32:

* the gradient of torch.clamp(v, 1e-8, 1e4) (network/Dorn.py:308) passes on the closed interval - the kernel tests
  `clamp(v) == v` (one comparison) instead of `v >= 1e-8 and v <= 1e4` (two), NaN included;
* OrdinalRegressionLoss reads plane k where `not (k > label)` (criteria.py:809-810) - the kernel counts those planes once
  per pixel as n_le = clamp(label + 1, 0, K) and compares k < n_le;
* walking K pairs in full groups of U plus one partial group visits every pair exactly once, in order."""
import numpy as np
import pytest


def test_clamp_pass_band_is_where_the_clamp_returns_its_input():
    lo, hi = np.float32(1e-8), np.float32(1e4)
    rng = np.random.default_rng(5)
    v = np.concatenate([
        rng.standard_normal(200000).astype(np.float32) * 4,
        np.array([0.0, -0.0, np.nan, np.inf, -np.inf, lo, hi, np.nextafter(lo, np.float32(0)), np.nextafter(lo, np.float32(1)),
                  np.nextafter(hi, np.float32(0)), np.nextafter(hi, np.float32(np.inf)), 1e-45, -1e-45], dtype=np.float32),
        (rng.random(50000).astype(np.float32) * 2e-8), (hi + (rng.random(50000).astype(np.float32) - 0.5) * 8)])
    with np.errstate(invalid="ignore"):
        clamped = np.where(v < lo, lo, v)                     # the kernel's clamp_logit: NaN propagates
        clamped = np.where(clamped > hi, hi, clamped)
        two_sided = (v >= lo) & (v <= hi)
        one_compare = clamped == v
    assert np.array_equal(two_sided, one_compare)
    assert not one_compare[np.isnan(v)].any()


@pytest.mark.parametrize("K", [1, 10, 68, 71])
def test_selected_plane_count_matches_the_reference_comparison(K):
    labels = np.array([-(2 ** 62), -5, -1, 0, 1, K - 2, K - 1, K, K + 1, 2 ** 62], dtype=np.int64)
    k = np.arange(K, dtype=np.int64)
    for label in labels:
        ref = ~(k > label)                                    # ord_c0 = 1 where not (k > label)
        n_le = 0 if label < 0 else (K if label >= K else int(label) + 1)
        assert np.array_equal(ref, k < n_le), (K, label)


@pytest.mark.parametrize("K,U", [(68, 8), (68, 4), (10, 8), (71, 8), (71, 4), (3, 8), (8, 8), (16, 16), (68, 16)])
def test_group_walk_visits_every_pair_once(K, U):
    seen = []
    k0 = 0
    while k0 + U <= K:
        seen += [k0 + u for u in range(U)]                    # FULL group
        k0 += U
    if k0 < K:
        seen += [k0 + u for u in range(U) if u < K - k0]      # partial group: u < npair
    assert seen == list(range(K))
