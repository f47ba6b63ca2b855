"""GPU parity of the fused metrics kernel (C ABI mde_metrics) against the pinned CPU oracle."""
import numpy as np
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import metrics as ometrics
from tests.gpu_util import T, close

pytestmark = pytest.mark.gpu
ALL = ["delta1", "delta2", "delta3", "mae", "mse", "log10", "msle", "absrel", "sqrel", "rmse", "rmse_true", "rmse_log"]
METRIC_RTOL = 1e-5   # BASELINE.json: float metrics within 1e-5 relative


@pytest.fixture(scope="module")
def M():
    from mono_depth_estimation_b200 import metrics
    return metrics


def _counts(res):
    raw = res["f64"][24:36].cpu().numpy()
    return [int(round(v)) for v in raw[:4]]


@pytest.mark.parametrize("ref_math", [False, True])
def test_small_golden(M, golden, ref_math):
    g = golden("metrics_small.npz")
    names = [str(n) for n in g["names"]]
    pred, gt = T(g["pred"]).cuda(), T(g["gt"]).cuda()
    mc = M.MetricComputation(names, reference_math=ref_math)
    vals = mc.compute(pred, gt)
    assert all(v.dim() == 0 and v.is_cuda and v.dtype == torch.float32 for v in vals)
    close(torch.stack(vals), g["values64"], METRIC_RTOL)
    res = M.fused_metrics(pred, gt, per_image=True, reference_math=ref_math)
    assert _counts(res) == [int(g["n_valid"])] + [int(c) for c in g["delta_counts"]]      # bit-exact
    idx = [ALL.index(n) for n in names]
    close(res["per_image"][:, idx], g["per_image64"], METRIC_RTOL)
    close(res["image_mean"][idx], g["per_image64"].mean(0), METRIC_RTOL)
    vb = M.MetricComputation(names, reference_math=ref_math).compute_batch(pred, gt)
    close(torch.stack(vb), g["per_image64"].mean(0), METRIC_RTOL)


def test_thresholds_are_strict(M, golden):
    g = golden("metrics_small.npz")
    p, t = T(g["thr_pred"]).cuda(), T(g["thr_gt"]).cuda()
    vals = M.MetricComputation(["delta1", "delta2", "delta3"]).compute(p, t)
    close(torch.stack(vals), g["thr_values"], 1e-7)
    assert _counts(M.fused_metrics(p, t)) == [7, 1, 5, 6]


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_config_vs_oracle(M, golden, name):
    pred, gt = synth.config_inputs(name)
    res = M.fused_metrics(pred.cuda(), gt.cuda())
    v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), ALL)]
    close(res["f64"][:12], v64, METRIC_RTOL)
    assert _counts(res) == list(ometrics.delta_counts(pred, gt))                            # bit-exact
    res_ref = M.fused_metrics(pred.cuda(), gt.cuda(), reference_math=True)
    close(res_ref["f64"][:12], v64, METRIC_RTOL)
    assert _counts(res_ref) == _counts(res)
    if name == "C1":
        g = golden("config_c1.npz")
        names = [str(n) for n in g["metric_names"]]
        close(res["f64"][[ALL.index(n) for n in names]], g["metrics64"], METRIC_RTOL)
        assert _counts(res) == [int(g["n_valid"])] + [int(c) for c in g["delta_counts"]]


def test_default_metric_groups_and_running_avg(M, golden):
    g = golden("metrics_small.npz")
    pred, gt = T(g["pred"]).cuda(), T(g["gt"]).cuda()
    names = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]           # reference evaluate.py:11
    mc = M.MetricComputation(names)
    vals = mc.compute(pred, gt)
    ref = [float(v) for v in ometrics.compute(pred.cpu().double(), gt.cpu().double(), names)]
    close(torch.stack(vals), ref, METRIC_RTOL)
    mc2 = M.MetricComputation(["absrel", "mae"])
    mc2.compute(pred[:2], gt[:2]); mc2.compute(pred[2:], gt[2:])
    close(torch.stack([mc2.avg("absrel"), mc2.avg(1)]), g["running_avg"], METRIC_RTOL)
    assert mc2.count == 2
    mc2.reset()
    assert mc2.count == 0 and mc2.sum == [0.0, 0.0]
    for n in ("absrel", "msle", "delta2"):                                             # METRICS[name](p_1d, t_1d)
        p1, t1 = ometrics.gather_valid(pred.cpu(), gt.cpu())
        close(M.METRICS[n](p1.cuda(), t1.cuda()), float(ometrics.METRIC_FNS[n](p1.double(), t1.double())), METRIC_RTOL)


@pytest.mark.parametrize("shape", [(3, 1, 33, 41), (2, 1, 7, 5), (1, 1, 1, 3), (5, 2, 12, 20)])
def test_odd_shapes_and_repeat_calls(M, shape):
    pred, gt = synth.depth_pair(shape, 21, border=1)
    for _ in range(2):   # second call checks the self-cleaning workspace
        res = M.fused_metrics(pred.cuda(), gt.cuda(), per_image=True)
        v64 = [float(v) for v in ometrics.compute(pred.double(), gt.double(), ALL)]
        close(res["f64"][:12], v64, METRIC_RTOL)
        assert _counts(res) == list(ometrics.delta_counts(pred, gt))
        n_img = pred.numel() // (shape[-1] * shape[-2])
        assert res["per_image"].shape == (n_img, 12)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_pred(M, dtype):
    pred, gt = synth.depth_pair((2, 1, 24, 32), 22, border=1)
    ph = pred.to(dtype)
    res = M.fused_metrics(ph.cuda(), gt.cuda())
    v64 = [float(v) for v in ometrics.compute(ph.double(), gt.double(), ALL)]
    close(res["f64"][:12], v64, METRIC_RTOL)
    assert _counts(res) == list(ometrics.delta_counts(ph.float(), gt))


def test_invalid_target_and_clamp(M):
    pred = torch.ones(1, 1, 8, 8).cuda()
    with pytest.raises(AssertionError, match="invalid target!"):
        M.MetricComputation(["mae"]).compute(pred, torch.zeros(1, 1, 8, 8).cuda())
    vals = M.MetricComputation(["mae"], strict=False).compute(pred, torch.zeros(1, 1, 8, 8).cuda())
    assert torch.isnan(vals[0])
    # negative / zero predictions are clamped to 1e-7 (metrics.py:59)
    p = torch.tensor([[[[-1.0, 0.0, 2.0, 3.0]]]]); t = torch.tensor([[[[1.0, 2.0, 2.0, 0.0]]]])
    res = M.fused_metrics(p.cuda(), t.cuda())
    close(res["f64"][:12], [float(v) for v in ometrics.compute(p.double(), t.double(), ALL)], METRIC_RTOL)
    # NaN prediction on a valid pixel poisons the float metrics but not the counts' validity
    p = torch.tensor([[[[float("nan"), 1.0, 2.0, 3.0]]]]); t = torch.tensor([[[[1.0, 1.0, 2.0, 3.0]]]])
    res = M.fused_metrics(p.cuda(), t.cuda())
    assert torch.isnan(res["values"][3]) and _counts(res) == [4, 3, 3, 3]


def test_batch_linearity_at_scale(M):
    """Size-independent property at a large size: raw sums over a batch = sum over its halves, counts
    exactly, and per-image rows do not depend on which other images share the launch."""
    pred, gt = synth.depth_pair((48, 1, 480, 640), 23, device="cuda")
    full = M.fused_metrics(pred, gt, per_image=True)
    a = M.fused_metrics(pred[:17], gt[:17], per_image=True)
    b = M.fused_metrics(pred[17:], gt[17:], per_image=True)
    raw_f, raw_a, raw_b = full["f64"][24:36], a["f64"][24:36], b["f64"][24:36]
    assert torch.equal(raw_f[:4], raw_a[:4] + raw_b[:4])
    close(raw_f[4:], raw_a[4:] + raw_b[4:], 2e-6)   # different CTA partitions -> different fp32 tile groupings
    pir = torch.cat([a["per_image_raw"], b["per_image_raw"]])
    assert torch.equal(full["per_image_raw"][:, :4], pir[:, :4])
    close(full["per_image_raw"], pir, 2e-6)
    # and against the oracle on a 3-image slice
    sl = slice(5, 8)
    v64 = ometrics.compute_per_image_mean(pred[sl].cpu().double(), gt[sl].cpu().double(), ALL)
    close(full["per_image"][sl].mean(0), v64, METRIC_RTOL)


def test_delta_counts_bit_exact_near_thresholds(M):
    """Adversarial: ratios within a few ulp of 1.25^k in both orientations (p/t and t/p), where the fast
    path must fall back to the exact IEEE divide. Counts must equal the reference's, pixel for pixel."""
    g = torch.Generator().manual_seed(77)
    n = 1 << 20
    t = (torch.rand(n, generator=g, dtype=torch.float64) * 9.5 + 0.5).float()
    k = torch.randint(1, 4, (n,), generator=g)
    thr = torch.tensor([1.25, 1.5625, 1.953125], dtype=torch.float64)[k - 1]
    ulps = torch.randint(-8, 9, (n,), generator=g).double()
    flip = torch.rand(n, generator=g) < 0.5
    ratio = thr * (1.0 + ulps * 2.0 ** -24)
    p = torch.where(flip, t.double() / ratio, t.double() * ratio).float()
    p[::97] = t[::97] * 1.25                       # exact-threshold products
    pred, gt = p.view(4, 1, 512, 512), t.view(4, 1, 512, 512)
    ref = ometrics.delta_counts(pred, gt)
    for ref_math in (False, True):
        res = M.fused_metrics(pred.cuda(), gt.cuda(), per_image=True, reference_math=ref_math)
        assert _counts(res) == list(ref)
        for b in range(4):
            rb = ometrics.delta_counts(pred[b:b + 1], gt[b:b + 1])
            assert [int(round(v)) for v in res["per_image_raw"][b, :4].tolist()] == list(rb)
    # half of the sampled ratios sit within 1e-5 (log_1.25 units) of a threshold: the slow path is exercised
    assert 0.05 < ref[1] / ref[0] < 0.95


@pytest.mark.parametrize("pshape,tshape", [((2, 1, 240, 320), (2, 1, 427, 561)), ((1, 1, 257, 353), (1, 1, 480, 640)),
                                           ((3, 1, 480, 640), (3, 1, 480, 640)), ((2, 1, 109, 147), (2, 1, 55, 74))])
def test_metrics_of_bilinearly_resized_inputs(M, pshape, tshape):
    """the test steps of eigen / dorn / my resize prediction AND target to 480 x 640 with
    F.interpolate(mode='bilinear') before log_test (modules/eigen.py:49-51). Oracle: those two torch ops on the CPU
    followed by the metric oracle. Float metrics within 1e-5; the delta counts may move by the few pixels whose
    interpolated ratio sits within an ulp of a threshold (CPU and GPU blend in a different op order)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(pshape[2] + tshape[3])
    target = torch.rand(tshape, generator=g) * 9.5 + 0.5
    target[torch.rand(tshape, generator=g) < 0.2] = 0.0
    target[:, :, :7, :] = 0.0
    pred = torch.rand(pshape, generator=g) * 9.0 + 0.6
    size = (480, 640)
    y_hat = F.interpolate(pred, size, mode="bilinear")
    y = F.interpolate(target, size, mode="bilinear")
    v64 = [float(v) for v in ometrics.compute(y_hat.double(), y.double(), ALL)]
    ref_counts = list(ometrics.delta_counts(y_hat, y))
    res = M.fused_metrics_resized(pred.cuda(), target.cuda(), size, per_image=True)
    close(res["f64"][:12], v64, 2e-5)
    got = _counts(res)
    assert got[0] == ref_counts[0] or abs(got[0] - ref_counts[0]) <= 4          # y' > 0 at the same pixels
    assert all(abs(a - b) <= 8 for a, b in zip(got, ref_counts)), (got, ref_counts)
    assert res["per_image"].shape == (pshape[0], 12)
    if pshape == tshape and pshape[-2:] == size:   # identity resize: exactly the plain kernel's numbers
        plain = M.fused_metrics(pred.cuda(), target.cuda())
        assert _counts(plain) == got
        close(res["f64"][:12], plain["f64"][:12], 1e-6)
    vals = M.MetricComputation(["delta1", "absrel", "rmse"], strict=False).compute_resized(pred.cuda(), target.cuda(), size)
    close(torch.stack(vals), [v64[ALL.index(n)] for n in ("delta1", "absrel", "rmse")], 2e-5)


@pytest.mark.parametrize("pshape,tshape", [((2, 1, 240, 320), (2, 1, 427, 561)), ((1, 1, 257, 353), (1, 1, 480, 640)),
                                           ((2, 1, 109, 147), (2, 1, 55, 74)), ((2, 1, 228, 304), (2, 1, 480, 640))])
def test_resized_metrics_counts_exact_vs_cuda_interpolate(M, pshape, tshape):
    """Integer outputs of the resize-fused metric pass against what the reference executes on a GPU: the two
    F.interpolate(mode='bilinear') calls ON CUDA (modules/eigen.py:49-51, modules/dorn.py:181-183) followed by the metric
    oracle's op chain on those CUDA tensors. The kernel samples with ATen's CUDA rule and op order, so the interpolated
    values agree bit for bit and the valid / delta counts must be EQUAL, pooled and per image (no slack)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7 * pshape[2] + tshape[3])
    target = torch.rand(tshape, generator=g) * 9.5 + 0.5
    target[torch.rand(tshape, generator=g) < 0.2] = 0.0
    target[:, :, :7, :] = 0.0
    pred = torch.rand(pshape, generator=g) * 9.0 + 0.6
    # ratios parked on the thresholds in the SOURCE images survive the blend only where the taps agree; the random ones matter
    size = (480, 640)
    pc, tc = pred.cuda(), target.cuda()
    y_hat = F.interpolate(pc, size, mode="bilinear")
    y = F.interpolate(tc, size, mode="bilinear")
    ref_counts = list(ometrics.delta_counts(y_hat, y))                       # torch ops on CUDA tensors
    res = M.fused_metrics_resized(pc, tc, size, per_image=True)
    assert _counts(res) == ref_counts
    for b in range(pshape[0]):
        rb = ometrics.delta_counts(y_hat[b:b + 1], y[b:b + 1])
        assert [int(round(v)) for v in res["per_image_raw"][b, :4].tolist()] == list(rb)
    v64 = [float(v) for v in ometrics.compute(y_hat.double().cpu(), y.double().cpu(), ALL)]
    close(res["f64"][:12], v64, 1e-5)


class _EmulatedRank:
    """Stand-in for distributed.PeerComm: rank `rank` of an emulated world whose mailboxes all live on THIS GPU."""

    def __init__(self, handle, world):
        self.handle, self.world, self.seq = handle, world, 0

    def next_seq(self):
        self.seq += 1
        return self.seq


def test_in_kernel_exchange_emulated_on_one_gpu(M):
    """mde_metrics_sharded with three 'ranks' emulated on one GPU: three mailboxes and three communicators built straight
    from the C ABI, one launch per rank on its own stream - the finalisers wait for one another exactly as they do across
    NVLink. All ranks end with bit-identical result vectors; image count, valid count and delta counts equal the single
    launch over all images bit for bit (and the CPU oracle's per-image means to 1e-5); an empty shard takes part; the
    rows are double-buffered, so back-to-back calls do not disturb one another."""
    import ctypes as C
    from mono_depth_estimation_b200 import _lib
    lib = _lib.load()
    world = 3
    boxes = []
    for _ in range(world):
        p = C.c_void_p()
        _lib.check(lib.mde_peer_alloc(_lib.PEER_MAILBOX_BYTES, C.byref(p)))
        boxes.append(p)
    table = (C.c_void_p * world)(*[b.value for b in boxes])
    comms = []
    for r in range(world):
        h = C.c_void_p()
        _lib.check(lib.mde_peer_comm_create(table, r, world, 3000, C.byref(h)))
        comms.append(_EmulatedRank(h, world))
    names = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse", "absrel", "sqrel", "msle"]
    pred, gt = synth.depth_pair((7, 1, 120, 160), 321, border=3)
    gt[2] = 0                                                   # an image without a valid pixel
    pred, gt = pred.cuda(), gt.cuda()
    single = M.fused_metrics(pred, gt, names=names)["f64"].cpu().numpy()
    streams = [torch.cuda.Stream() for _ in range(world)]
    for st in streams:      # per-stream workspaces and allocator blocks exist BEFORE the first exchange: a cudaMalloc between
        with torch.cuda.stream(st):   # the launches would wait for the first rank's kernel, which waits for the others
            M.fused_metrics(pred, gt, names=names)
    # ... and so is the empty-shard kernel's code (CUDA loads a kernel lazily at its first launch, which waits for the device):
    # one launch through a communicator of a world of one, which is also a legal (if pointless) configuration
    solo_tab = (C.c_void_p * 1)(boxes[0].value)
    solo_h = C.c_void_p()
    _lib.check(lib.mde_peer_comm_create(solo_tab, 0, 1, 3000, C.byref(solo_h)))
    solo = _EmulatedRank(solo_h, 2)     # (world > 1 on the Python side selects the sharded entry point)
    alone = M.fused_metrics(pred[:0], gt[:0], names=names, comm=solo)["f64"]
    torch.cuda.synchronize()
    assert float(alone[2 * 12 + 12]) == 0.0 and float(alone[2 * 12]) == 0.0      # no image, no valid pixel
    lib.mde_peer_comm_destroy(solo_h)
    lib.mde_peer_free(boxes[0])
    _lib.check(lib.mde_peer_alloc(_lib.PEER_MAILBOX_BYTES, C.byref(boxes[0])))   # a fresh (zeroed) mailbox for rank 0
    table = (C.c_void_p * world)(*[b.value for b in boxes])
    for c in comms:
        lib.mde_peer_comm_destroy(c.handle)
    comms = []
    for r in range(world):
        h = C.c_void_p()
        _lib.check(lib.mde_peer_comm_create(table, r, world, 3000, C.byref(h)))
        comms.append(_EmulatedRank(h, world))
    torch.cuda.synchronize()
    try:
        for shards in ([(0, 3), (3, 5), (5, 7)], [(0, 7), (7, 7), (7, 7)], [(0, 1), (1, 6), (6, 7)]):   # the middle one: two empty shards
            outs = []
            for r, (a, b) in enumerate(shards):
                streams[r].wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(streams[r]):
                    outs.append(M.fused_metrics(pred[a:b], gt[a:b], names=names, comm=comms[r])["f64"])
            torch.cuda.synchronize()
            outs = [o.cpu().numpy() for o in outs]
            for o in outs[1:]:
                assert np.array_equal(o, outs[0])               # same summation order on every rank
            NM, NQ = 12, 12
            packed, ref = outs[0][2 * NM:], single[2 * NM:]
            assert packed[NQ] == ref[NQ] == 6.0                 # images with a valid pixel
            assert np.array_equal(packed[:4], ref[:4])          # valid count and the three delta counts
            np.testing.assert_allclose(outs[0], single, rtol=2e-6, atol=0)   # fp32 tile sums grouped by other CTA partitions
        # the stand-alone sum of a few doubles through the same mailboxes (mde_peer_allreduce_f64), clearing the source
        srcs = [torch.arange(12, dtype=torch.float64, device="cuda") * (r + 1) + 0.25 for r in range(world)]
        dsts = [torch.empty(12, dtype=torch.float64, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ws = _lib.workspace(torch.device("cuda", torch.cuda.current_device()), 1)
                _lib.check(lib.mde_peer_allreduce_f64(_lib.ptr(srcs[r]), _lib.ptr(dsts[r]), 12, 1, comms[r].handle, comms[r].next_seq(),
                                                      _lib.ptr(ws), _lib.stream_ptr(srcs[r].device)))
        torch.cuda.synchronize()
        for r in range(world):
            assert np.array_equal(dsts[r].cpu().numpy(), np.arange(12) * 6.0 + 0.75) and float(srcs[r].abs().sum()) == 0.0
        keep = [0, 1, 3, 4, 5, 6]                             # the reference's eval loop never sees an image without a valid pixel
        want = ometrics.compute_per_image_mean(pred.cpu()[keep].double(), gt.cpu()[keep].double(), names)
        idx = [ALL.index(n) for n in names]
        close(outs[0][12:24][idx], [float(v) for v in want], METRIC_RTOL)
    finally:
        torch.cuda.synchronize()
        for c in comms:
            lib.mde_peer_comm_destroy(c.handle)
        for b in boxes:
            lib.mde_peer_free(b)


def test_in_kernel_exchange_gives_up_on_a_missing_peer(M):
    """A rank whose peer never launches must not hang the GPU: after the communicator's timeout (here 60 ms) the finaliser
    writes NaN results and raises the workspace's error flag, and the launch completes."""
    import ctypes as C
    import time
    from mono_depth_estimation_b200 import _lib
    lib = _lib.load()
    boxes = []
    for _ in range(2):
        p = C.c_void_p()
        _lib.check(lib.mde_peer_alloc(_lib.PEER_MAILBOX_BYTES, C.byref(p)))
        boxes.append(p)
    table = (C.c_void_p * 2)(*[b.value for b in boxes])
    h = C.c_void_p()
    _lib.check(lib.mde_peer_comm_create(table, 0, 2, 60, C.byref(h)))
    pred, gt = synth.depth_pair((2, 1, 60, 80), 5)
    pred, gt = pred.cuda(), gt.cuda()
    M.fused_metrics(pred, gt, names=["delta1"])
    torch.cuda.synchronize()
    try:
        t0 = time.time()
        out = M.fused_metrics(pred, gt, names=["delta1"], comm=_EmulatedRank(h, 2))["f64"]
        torch.cuda.synchronize()
        dt = time.time() - t0
        assert 0.05 < dt < 2.0, dt
        assert bool(torch.isnan(out[:12]).all()) and bool(torch.isnan(out[24:37]).all())
        ws = _lib.workspace(pred.device, 2)
        assert int(ws[20:24].view(torch.int32)[0]) == 1          # WsHeader.error
        ws[20:24].zero_()
        ok = M.fused_metrics(pred, gt, names=["delta1"])["f64"]   # the workspace is clean: the next plain call is right
        assert abs(float(ok[0]) - float(ometrics.compute(pred.cpu().double(), gt.cpu().double(), ["delta1"])[0])) < 1e-6
    finally:
        torch.cuda.synchronize()
        lib.mde_peer_comm_destroy(h)
        for b in boxes:
            lib.mde_peer_free(b)
