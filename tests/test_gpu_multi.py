"""Two-GPU NCCL tests of the sharded paths: one process per GPU, spawned here. Skipped on boxes with one GPU
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the result is kept under profiles/)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
EVAL = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse", "absrel", "sqrel", "msle"]


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from mono_depth_estimation_b200 import criteria as Cr, distributed as D, metrics as M, synth, _lib
    out = {}
    # ---- sharded evaluation: 13 images 480x640 over two ranks (7 + 6), one image without a valid pixel
    pred, gt = synth.depth_pair((13, 1, 480, 640), 205)
    gt[5] = 0
    a, b = D.shard_range(13, rank, world)
    res = D.sharded_eval(pred[a:b].to(dev), gt[a:b].to(dev), EVAL)
    out["eval"] = {"image_mean": [float(res["image_mean"][n]) for n in EVAL], "pooled": [float(res["pooled"][n]) for n in EVAL],
                   "n_images": float(res["n_images"]), "n_valid": float(res["n_valid"]), "delta_counts": [float(c) for c in res["delta_counts"]]}
    if rank == 0:   # the single-GPU answer for the same 13 images
        full = M.fused_metrics(pred.to(dev), gt.to(dev), names=EVAL)
        f64 = full["f64"].cpu()
        NM, NQ = _lib.METRIC_NM, _lib.METRIC_NQ
        idx = [_lib.METRIC_INDEX[n] for n in EVAL]
        out["eval_single"] = {"image_mean": [float(f64[NM + i]) for i in idx], "pooled": [float(f64[i]) for i in idx],
                              "n_images": float(f64[2 * NM + NQ]), "n_valid": float(f64[2 * NM]), "delta_counts": [float(f64[2 * NM + k]) for k in (1, 2, 3)]}
    # ---- the same evaluation with the exchange INSIDE the launch (NVLink peer mailboxes, no collective), three calls in a
    # row (the rows are double-buffered by call parity), then a 1-image set (rank 1's shard is empty)
    comm = D.PeerComm()
    n0 = _lib.launch_count()
    for rep in range(3):
        res = D.sharded_eval(pred[a:b].to(dev), gt[a:b].to(dev), EVAL, comm=comm)
        torch.cuda.synchronize()
    out["peer_launches"] = (_lib.launch_count() - n0) / 3.0
    out["peer_eval"] = {"image_mean": [float(res["image_mean"][n]) for n in EVAL], "pooled": [float(res["pooled"][n]) for n in EVAL],
                        "n_images": float(res["n_images"]), "n_valid": float(res["n_valid"]), "delta_counts": [float(c) for c in res["delta_counts"]],
                        "packed": res["packed"].cpu().numpy()}
    a1, b1 = D.shard_range(1, rank, world)
    res1 = D.sharded_eval(pred[a1:b1].to(dev), gt[a1:b1].to(dev), EVAL, comm=comm)
    out["peer_one_image"] = {"n_images": float(res1["n_images"]), "delta1": float(res1["image_mean"]["delta1"]), "n_valid": float(res1["n_valid"])}
    if rank == 0:
        one = M.fused_metrics(pred[:1].to(dev), gt[:1].to(dev), names=EVAL)["f64"].cpu()
        out["one_image_single"] = {"n_images": float(one[2 * _lib.METRIC_NM + _lib.METRIC_NQ]), "delta1": float(one[_lib.METRIC_NM + _lib.METRIC_INDEX["delta1"]]),
                                   "n_valid": float(one[2 * _lib.METRIC_NM])}
    # ---- the stand-alone sum of a few doubles through the same mailboxes (what the training steps exchange per logging interval)
    acc = torch.arange(12, dtype=torch.float64, device=dev) * (rank + 1) + 0.5
    tot = torch.empty_like(acc)
    comm.all_reduce_(acc, out=tot, zero_src=True)
    torch.cuda.synchronize()
    out["peer_allreduce"] = {"sum": tot.cpu().numpy(), "src_after": acc.cpu().numpy()}
    comm.close()
    # ---- global-batch losses: C1 batch (8 images) sharded 4 + 4
    pred, gt = synth.config_inputs("C1")
    pred[6, 0, 50, 60] = gt[6, 0, 50, 60] + 25.0        # the berHu / Laina maximum lives on rank 1
    a, b = D.shard_range(8, rank, world)
    for name, make in (("silog", lambda: Cr.silog_loss(0.85)), ("berhu", Cr.berHuLoss), ("laina_berhu", Cr.LainaBerHuLoss), ("l1", Cr.MaskedL1Loss)):
        p = pred[a:b].to(dev).requires_grad_(True)
        loss = D.global_batch_loss(make(), p, gt[a:b].to(dev))
        loss.backward()
        out["gb_" + name] = {"loss": float(loss.detach()), "grad": p.grad.cpu().numpy()}
        if rank == 0:
            pf = pred.to(dev).requires_grad_(True)
            lf = make()(pf, gt.to(dev))
            lf.backward()
            out["full_" + name] = {"loss": float(lf.detach()), "grad": pf.grad.cpu().numpy()}
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def two_rank_results():
    _need_two()
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return results


def test_sharded_eval_equals_single_gpu(two_rank_results):
    """sharded_eval over NCCL (one launch + one 25-double all-reduce per rank) against the single-GPU launch over the same
    images: image count, valid count and the three delta counts bit for bit; float values to fp64 rounding of the sums."""
    single = two_rank_results[0]["eval_single"]
    for rank in (0, 1):
        ev = two_rank_results[rank]["eval"]
        assert ev["n_images"] == single["n_images"] == 12.0
        assert ev["n_valid"] == single["n_valid"] and ev["delta_counts"] == single["delta_counts"]
        np.testing.assert_allclose(ev["image_mean"], single["image_mean"], rtol=1e-6)   # fp32 tile sums grouped by another CTA partition
        np.testing.assert_allclose(ev["pooled"], single["pooled"], rtol=2e-6)   # different CTA partitions of the fp32 tile sums


def test_in_kernel_peer_exchange_equals_collective_and_single_gpu(two_rank_results):
    """sharded_eval(comm=PeerComm): ONE launch per rank, the 25 doubles exchanged by the launches' finalisers over NVLink
    peer memory. Both ranks end with bit-identical vectors (same summation order), the counts equal the single-GPU launch
    bit for bit, the float values equal the NCCL path's; an empty shard still takes part."""
    r0, r1 = two_rank_results[0], two_rank_results[1]
    assert r0["peer_launches"] == 1.0 and r1["peer_launches"] == 1.0
    assert np.array_equal(r0["peer_eval"]["packed"], r1["peer_eval"]["packed"])
    single = r0["eval_single"]
    for r in (r0, r1):
        ev = r["peer_eval"]
        assert ev["n_images"] == single["n_images"] == 12.0
        assert ev["n_valid"] == single["n_valid"] and ev["delta_counts"] == single["delta_counts"]
        np.testing.assert_allclose(ev["image_mean"], r["eval"]["image_mean"], rtol=1e-12)
        np.testing.assert_allclose(ev["pooled"], r["eval"]["pooled"], rtol=1e-12)
        assert r["peer_one_image"] == r0["one_image_single"]
        np.testing.assert_array_equal(r["peer_allreduce"]["sum"], np.arange(12) * 3.0 + 1.0)    # (i + 0.5) + (2 i + 0.5)
        np.testing.assert_array_equal(r["peer_allreduce"]["src_after"], np.zeros(12))


@pytest.mark.parametrize("name", ["silog", "berhu", "laina_berhu", "l1"])
def test_global_batch_loss_equals_single_gpu(two_rank_results, name):
    """global_batch_loss over NCCL: the loss is identical on both ranks and equals the single-GPU full-batch loss, the two
    gradient shards together equal the full-batch gradient (1e-5, north_star's tolerance)."""
    r0, r1 = two_rank_results[0], two_rank_results[1]
    full = r0["full_" + name]
    assert r0["gb_" + name]["loss"] == r1["gb_" + name]["loss"]
    np.testing.assert_allclose(r0["gb_" + name]["loss"], full["loss"], rtol=1e-5)
    grad = np.concatenate([r0["gb_" + name]["grad"], r1["gb_" + name]["grad"]])
    scale = np.abs(full["grad"]).max()
    np.testing.assert_allclose(grad, full["grad"], rtol=1e-5, atol=2e-6 * scale)
