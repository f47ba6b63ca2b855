"""Host-side model of the per-image exchange of the register-resident MaskedDepthLoss kernel (csrc/eigen.cu,
eigen_resident_kernel; reference criteria.py:38-41 needs n_b, S1_b, S2_b per image): a CTA covers 4096 consecutive
pixels, publishes one triple for its first image and one for the next, and every CTA gathers image b from the contiguous
CTA range that overlaps it. The index arithmetic is restated here with integers and checked against a direct per-image
sum for shapes on every side of its conditions (H W >= 4096, W % 4 == 0, images not aligned to CTAs, more images than
warps). No GPU, no library call: this pins the arithmetic the kernel's gather relies on."""
import numpy as np
import pytest

CTA_PX = 4096


def gather_per_image(values, n_img, hw):
    """values: one integer per pixel (a stand-in for the mask / d / d^2 terms). Returns the per-image sums formed the
    way the kernel forms them."""
    npx = n_img * hw
    nq = npx // 4
    grid = (nq + 1023) // 1024
    rows = np.zeros((grid, 2), dtype=np.int64)                 # per CTA: {first image, next image}
    for cta in range(grid):
        img_a = (cta * CTA_PX) // hw
        lo, hi = cta * CTA_PX, min((cta + 1) * CTA_PX, npx)
        for q0 in range(lo, hi, 4):                            # a quad never straddles an image: hw % 4 == 0
            img = q0 // hw
            assert img in (img_a, img_a + 1), "a CTA may hold at most two images"
            rows[cta, 0 if img == img_a else 1] += values[q0:q0 + 4].sum()
    out = np.zeros(n_img, dtype=np.int64)
    for b in range(n_img):
        first_px = b * hw
        c_lo = first_px // CTA_PX
        c_hi = min((first_px + hw - 1) // CTA_PX, grid - 1)
        for c in range(c_lo, c_hi + 1):
            ia = (c * CTA_PX) // hw
            assert ia in (b, b - 1)
            out[b] += rows[c, 0 if ia == b else 1]
    return out, rows


@pytest.mark.parametrize("n_img,h,w", [(8, 228, 304), (3, 64, 64), (2, 65, 68), (37, 64, 64), (1, 64, 64), (5, 120, 160),
                                       (7, 64, 68), (4, 100, 44)])
def test_two_triples_per_cta_cover_every_image_exactly_once(n_img, h, w):
    hw = h * w
    assert hw >= CTA_PX and w % 4 == 0
    rng = np.random.default_rng(n_img * 1000 + h)
    values = rng.integers(0, 1000, size=n_img * hw, dtype=np.int64)
    got, rows = gather_per_image(values, n_img, hw)
    want = values.reshape(n_img, hw).sum(axis=1)
    assert np.array_equal(got, want)
    assert rows.sum() == values.sum()                          # nothing is published twice or dropped
