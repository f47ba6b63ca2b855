"""The PLY writer is host code: byte-for-byte parity with the reference's Python assembly
(depth2pointcloud.py:131-154), restated here because that file is a Blender script that cannot be imported."""
import numpy as np
import pytest


def reference_ply(front_verts, back_verts, color):
    """depth2pointcloud.py:131-154, verbatim logic."""
    points = []
    for v in range(color.shape[0]):
        if not np.isnan(front_verts[v, 0]):
            points.append("%f %f %f %d %d %d 0\n" % (front_verts[v, 0], front_verts[v, 1], front_verts[v, 2], color[v, 2], color[v, 1], color[v, 0]))
        if back_verts is not None and not np.isnan(back_verts[v, 0]):
            points.append("%f %f %f %d %d %d 0\n" % (back_verts[v, 0], back_verts[v, 1], back_verts[v, 2], color[v, 2], color[v, 1], color[v, 0]))
    return '''ply
format ascii 1.0
element vertex %d
property float x
property float y
property float z
property uchar red
property uchar green
property uchar blue
property uchar alpha
end_header
%s
''' % (len(points), "".join(points))


@pytest.mark.parametrize("with_back", [True, False])
def test_ply_bytes_equal_reference(tmp_path, with_back):
    from mono_depth_estimation_b200 import pointcloud as PC
    rs = np.random.RandomState(4)
    n = 48 * 64
    front = rs.randn(n, 3) * np.array([3.0, 2.0, 9.0])
    front[rs.rand(n) < 0.2] = np.nan                      # invalid pixels: every coordinate NaN after the world transform
    front[5] = [1e-7, -1e-7, 123456.789]                  # rounding to six decimals, negative zero text
    front[6] = [-0.0, 0.0000005, 0.0000015]
    back = rs.randn(n, 3) * 5.0
    back[rs.rand(n) < 0.5] = np.nan
    color = rs.randint(0, 256, size=(n, 3)).astype(np.uint8)
    color[0] = [0, 9, 10]; color[1] = [99, 100, 255]
    path = tmp_path / "frame.ply"
    count = PC.write_ply(path, front.reshape(48, 64, 3), color.reshape(48, 64, 3), back.reshape(48, 64, 3) if with_back else None)
    ref = reference_ply(front, back if with_back else None, color)
    assert path.read_bytes() == ref.encode()
    assert count == int(ref.split("element vertex ")[1].split("\n")[0])


def test_ply_empty_and_errors(tmp_path):
    from mono_depth_estimation_b200 import pointcloud as PC
    front = np.full((4, 3), np.nan)
    color = np.zeros((4, 3), np.uint8)
    assert PC.write_ply(tmp_path / "e.ply", front, color) == 0
    assert (tmp_path / "e.ply").read_bytes() == reference_ply(front, None, color).encode()
    with pytest.raises(ValueError):
        PC.write_ply(tmp_path / "x.ply", front, np.zeros((3, 3), np.uint8))
    with pytest.raises(Exception):
        PC.write_ply(tmp_path / "no_such_dir" / "x.ply", front, color)
