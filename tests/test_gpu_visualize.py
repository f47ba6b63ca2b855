"""GPU parity of colored_depthmap (reference visualize.py:8-17) with the reference's own recipe: numpy float32
arithmetic + cv2.applyColorMap(COLORMAP_INFERNO). Bit-exact (integer / byte work)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def reference(depth, d_min=None, d_max=None, do_mapping=True):
    """visualize.py:8-17, restated line for line."""
    if d_min is None:
        d_min = np.min(depth)
    if d_max is None:
        d_max = np.max(depth)
    depth_relative = (depth - d_min) / (d_max - d_min)
    depth_relative *= 255
    depth_relative = depth_relative.astype(np.uint8)
    if do_mapping:
        return cv2.applyColorMap(depth_relative, cv2.COLORMAP_INFERNO)
    return depth_relative


@pytest.mark.parametrize("shape", [(480, 640), (33, 41), (1, 5), (228, 304)])
def test_colored_depthmap(shape):
    from mono_depth_estimation_b200 import visualize as V
    rs = np.random.RandomState(shape[0])
    depth = (rs.rand(*shape) * 9.5 + 0.5).astype(np.float32)
    depth[0, 0] = 0.0
    out = V.colored_depthmap(depth)
    assert isinstance(out, np.ndarray) and out.dtype == np.uint8 and out.shape == shape + (3,)
    assert np.array_equal(out, reference(depth))
    assert np.array_equal(V.colored_depthmap(depth, do_mapping=False), reference(depth, do_mapping=False))
    # a shared range for two panels, as merge_into_row does (visualize.py:26-29)
    lo, hi = np.float32(0.25), np.float32(11.0)
    assert np.array_equal(V.colored_depthmap(depth, lo, hi), reference(depth, lo, hi))
    assert np.array_equal(V.colored_depthmap(depth, d_max=hi), reference(depth, d_max=hi))
    t = V.colored_depthmap(torch.from_numpy(depth).cuda())
    assert t.is_cuda and t.dtype == torch.uint8 and np.array_equal(t.cpu().numpy(), out)
    # every level of the table
    ramp = np.linspace(0.0, 1.0, 4096, dtype=np.float32).reshape(64, 64)
    assert np.array_equal(V.colored_depthmap(ramp), reference(ramp))
