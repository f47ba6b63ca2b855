"""GPU parity of the virtual-normal loss and the point-cloud back-projection."""
import math

import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import pointcloud as opc
from oracle import vnl as ovnl
from tests.gpu_util import LOSS_RTOL, T, close, grad_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Cr():
    from mono_depth_estimation_b200 import criteria
    return criteria


def _vnl(Cr, gt, pred, trip, select=True):
    H, W = gt.shape[-2:]
    v = Cr.VNL_Loss(519.0, 519.0, (H, W))
    v.set_triplets(trip.cuda())
    p = pred.cuda().clone().requires_grad_(True)
    loss = v(gt.cuda(), p, select=select)
    loss.backward()
    return loss.detach(), p.grad.detach(), v.last_stats.cpu()


@pytest.mark.parametrize("sel", [True, False])
def test_vnl_small_golden(Cr, golden, sel):
    g = golden("vnl_small.npz")
    gt, pred, trip = T(g["gt"]), T(g["pred"]), T(g["trip"])
    loss, grad, stats = _vnl(Cr, gt, pred, trip, sel)
    close(loss, g[f"loss64_sel{int(sel)}"], LOSS_RTOL)
    grad_close(grad, g[f"grad64_sel{int(sel)}"])
    _, per, mask = ovnl.vnl_loss(gt.double(), pred.double(), trip, 519.0, 519.0, select=sel, return_parts=True)
    assert int(stats[0]) == int(mask.sum())                                    # same valid-triplet set
    if sel:
        assert int(stats[1]) == int(int(mask.sum()) * 0.25)


def test_vnl_vs_oracle_medium(Cr):
    gt, pred, trip = synth.vnl_inputs((4, 1, 97, 129), 51, n_triplets=6000, pad_rows=10, zero_frac=5e-3)
    p64 = pred.double().clone().requires_grad_(True)
    l64 = ovnl.vnl_loss(gt.double(), p64, trip, 519.0, 519.0)
    (g64,) = torch.autograd.grad(l64, p64)
    loss, grad, stats = _vnl(Cr, gt, pred, trip)
    close(loss, l64.detach(), LOSS_RTOL)
    grad_close(grad, g64)
    assert float(stats[4]) == 1.0                                               # no tie at the threshold


def test_vnl_random_sampling_runs(Cr):
    """select_index() default path (random triplets, reference criteria.py:912-932): finite, repeatable shape."""
    gt, pred, _ = synth.vnl_inputs((2, 1, 64, 80), 52, n_triplets=10)
    v = Cr.VNL_Loss(519.0, 519.0, (64, 80))
    p = pred.cuda().requires_grad_(True)
    loss = v(gt.cuda(), p)
    loss.backward()
    assert torch.isfinite(loss) and p.grad.shape == pred.shape and float(p.grad.abs().sum()) > 0


def test_point_cloud(Cr):
    from mono_depth_estimation_b200 import pointcloud as PC
    rs = np.random.RandomState(3)
    depth = (rs.rand(37, 53) * 12).astype(np.float32)
    depth[0, :5] = 0.05; depth[3, 3] = 200.0                                    # outside the clip planes
    cam = PC.Camera(angle_x=0.8575560450553894, clip_start=0.1, clip_end=100.0,
                    matrix_world=[[0.68, -0.32, 0.65, 7.35], [0.73, 0.31, -0.61, -6.92], [-0.01, 0.89, 0.45, 4.95], [0, 0, 0, 1]])
    ref = opc.point_cloud(depth, cam.angle_x, cam.clip_start, cam.clip_end)
    out = PC.point_cloud(depth, cam)
    assert out.dtype == np.float64 and out.shape == (37, 53, 3)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    np.testing.assert_allclose(np.nan_to_num(out), np.nan_to_num(ref), rtol=1e-7, atol=1e-9)
    out32 = PC.point_cloud(torch.from_numpy(depth).cuda(), cam, dtype=torch.float32)
    np.testing.assert_allclose(np.nan_to_num(out32.cpu().numpy()), np.nan_to_num(ref), rtol=1e-6, atol=1e-6)
    refw = opc.to_world(ref, cam.matrix_world)
    outw = PC.point_cloud_world(depth, cam)
    assert np.array_equal(np.isnan(outw), np.isnan(refw))
    np.testing.assert_allclose(np.nan_to_num(outw), np.nan_to_num(refw), rtol=1e-6, atol=1e-6)
    batch = torch.from_numpy(np.stack([depth, depth * 0.5])).cuda()
    ob = PC.point_cloud(batch, cam)
    np.testing.assert_allclose(np.nan_to_num(ob[0].cpu().numpy()), np.nan_to_num(ref), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("shape", [(48, 64), (30, 1028), (480, 640)])
def test_point_cloud_tiled_path(Cr, shape):
    """w % 4 == 0 takes the 128-bit tiled kernel (shared-memory staged stores, reciprocal + residual divide):
    same values as the fp64 oracle to the last bits, partial last tile, batch, world transform."""
    from mono_depth_estimation_b200 import pointcloud as PC
    rs = np.random.RandomState(11)
    depth = (rs.rand(3, *shape) * 12).astype(np.float32)
    depth[:, 0, :7] = 0.05; depth[1, 3, 3] = 200.0
    cam = PC.Camera(angle_x=0.8575560450553894, clip_start=0.1, clip_end=100.0,
                    matrix_world=[[0.68, -0.32, 0.65, 7.35], [0.73, 0.31, -0.61, -6.92], [-0.01, 0.89, 0.45, 4.95], [0, 0, 0, 1]])
    out = PC.point_cloud(torch.from_numpy(depth).cuda(), cam).cpu().numpy()
    assert out.dtype == np.float64 and out.shape == (3, *shape, 3)
    for b in range(3):
        ref = opc.point_cloud(depth[b], cam.angle_x, cam.clip_start, cam.clip_end)
        assert np.array_equal(np.isnan(out[b]), np.isnan(ref))
        np.testing.assert_allclose(np.nan_to_num(out[b]), np.nan_to_num(ref), rtol=1e-15, atol=0)
        assert np.array_equal(np.signbit(out[b][..., 0]), np.signbit(ref[..., 0]))      # -0.0 at invalid pixels
    ref0 = opc.point_cloud(depth[0], cam.angle_x, cam.clip_start, cam.clip_end)
    out32 = PC.point_cloud(torch.from_numpy(depth[0]).cuda(), cam, dtype=torch.float32).cpu().numpy()
    np.testing.assert_array_equal(np.nan_to_num(out32), np.nan_to_num(ref0.astype(np.float32)))
    outw = PC.point_cloud_world(depth[0], cam)
    np.testing.assert_allclose(np.nan_to_num(outw), np.nan_to_num(opc.to_world(ref0, cam.matrix_world)), rtol=1e-6, atol=1e-6)


def test_vnl_config_c4_full_size_properties(Cr):
    """C4 at full size (8 x 385 x 385, 100 000 triplets), size-independent properties: the loss does not depend on
    the order of the triplets nor on the order of the three points' columns being permuted together; the valid-set
    size and the trim count are integers that survive the permutation exactly; gradients agree to fp32 atomics noise."""
    gt, pred, trip = synth.vnl_inputs(synth.SHAPES["C4"], 104)
    loss, grad, stats = _vnl(Cr, gt, pred, trip)
    perm = torch.randperm(trip.shape[1], generator=torch.Generator().manual_seed(1))
    loss_p, grad_p, stats_p = _vnl(Cr, gt, pred, trip[:, perm].contiguous())
    assert int(stats[0]) == int(stats_p[0]) and int(stats[1]) == int(stats_p[1])
    assert int(stats[1]) == int(int(stats[0]) * 0.25)
    close(loss_p, loss, 2e-6)
    close(grad_p, grad, 1e-4, 2e-6 * float(grad.abs().max()))
    assert 0.2 < int(stats[0]) / (8 * trip.shape[1]) < 0.4        # the survey's probe: 25-32 % of the triplets are valid


def test_point_cloud_reference_golden(Cr, golden):
    """The kernel against outputs of the REFERENCE's own point_cloud (depth2pointcloud.py:12-31, golden vectors made by
    oracle/gen_golden.py from the reference source): same NaN pattern, same sign of zero, fp64 values to the last bit
    on the tiled path (w % 4 == 0) and to 1e-15 relative on the scalar path; fp32 output = the rounded fp64 values."""
    from mono_depth_estimation_b200 import pointcloud as PC
    g = golden("pointcloud.npz")
    angle_x, clip_start, clip_end = (float(v) for v in g["camera"])
    cam = PC.Camera(angle_x=angle_x, clip_start=clip_start, clip_end=clip_end)
    for tag in "abc":
        depth, ref = g["depth_" + tag], g["points_" + tag]
        out = PC.point_cloud(torch.from_numpy(depth).cuda(), cam).cpu().numpy()
        assert out.dtype == np.float64 and out.shape == ref.shape
        assert np.array_equal(np.isnan(out), np.isnan(ref)), tag
        assert np.array_equal(np.signbit(out), np.signbit(ref)), tag
        np.testing.assert_allclose(np.nan_to_num(out), np.nan_to_num(ref), rtol=1e-15, atol=0, err_msg=tag)
        if depth.shape[1] % 4 == 0:
            assert np.array_equal(np.nan_to_num(out), np.nan_to_num(ref)), tag
        out32 = PC.point_cloud(torch.from_numpy(depth).cuda(), cam, dtype=torch.float32).cpu().numpy()
        np.testing.assert_allclose(np.nan_to_num(out32), np.nan_to_num(ref.astype(np.float32)), rtol=2e-7, atol=0, err_msg=tag)
