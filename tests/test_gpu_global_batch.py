"""Global-batch (split-phase) losses, SURVEY 8e row 4: the batch is sharded by image, the shards exchange only the scalar
totals, and loss + gradient must equal the single-GPU full-batch call. Here the ranks are EMULATED on one GPU by driving the
stages of distributed.SplitLoss by hand over several shards (the kernels never wait on each other); the NCCL version of
the same exchange runs in tests/test_gpu_multi.py on two GPUs, its host logic over gloo in tests/test_distributed_cpu.py."""
import pytest
import torch

from mono_depth_estimation_b200 import synth
from oracle import losses as olosses
from tests.gpu_util import LOSS_RTOL, close, grad_close

pytestmark = pytest.mark.gpu
NAMES = ["l1", "mse", "berhu", "laina_berhu", "silog"]


def _make(Cr, name):
    return {"l1": Cr.MaskedL1Loss, "mse": Cr.MaskedMSELoss, "berhu": Cr.berHuLoss, "laina_berhu": Cr.LainaBerHuLoss,
            "silog": lambda: Cr.silog_loss(0.85)}[name]()


def _run_sharded(D, crit, pred, gt, cuts):
    """Emulated ranks: shard b holds images cuts[b]:cuts[b+1]. Returns (losses per shard, concatenated gradient)."""
    shards = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        p = pred[a:b].cuda().clone().requires_grad_(True)
        shards.append((p, D.SplitLoss(crit, p, gt[a:b].cuda())))
    if shards[0][1].needs_max:
        gmax = torch.stack([s.stage_max()[4] for _, s in shards]).max()          # all-reduce(MAX)
        for _, s in shards:
            s.partials[4] = gmax
    tot = torch.stack([s.stage_sums()[0:4] for _, s in shards]).sum(0)           # all-reduce(SUM)
    losses, grads = [], []
    for p, s in shards:
        s.partials[0:4] = tot
        loss = s.finish()
        loss.backward()
        losses.append(loss.detach())
        grads.append(p.grad.detach())
    return losses, torch.cat(grads)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("cuts", [[0, 3, 8], [0, 1, 4, 8], [0, 5, 5, 8]])
def test_sharded_equals_full_batch(name, cuts):
    from mono_depth_estimation_b200 import criteria as Cr, distributed as D
    pred, gt = synth.config_inputs("C1")                               # 8 x 1 x 228 x 304
    pred[2, 0, 100, 100] = gt[2, 0, 100, 100] + 30.0                   # the berHu maximum lives in one shard only
    l64, g64 = olosses.loss_and_grad(olosses.LOSSES[name], pred.double(), gt.double())
    losses, grad = _run_sharded(D, _make(Cr, name), pred, gt, cuts)
    for l in losses:                                                   # identical on every (emulated) rank
        assert float(l) == float(losses[0])
    close(losses[0], l64, LOSS_RTOL, msg=name)
    grad_close(grad, g64, msg=name)
    # and against the fused single-launch kernel of the same package on the full batch
    p = pred.cuda().clone().requires_grad_(True)
    full = _make(Cr, name)(p, gt.cuda())
    full.backward()
    close(losses[0], full.detach(), 2e-6, msg=name)
    grad_close(grad, p.grad, msg=name)


def test_local_losses_differ_from_the_global_one():
    """Why the mode exists: the mean of the ranks' LOCAL berHu / SILog losses is not the full-batch loss."""
    from mono_depth_estimation_b200 import criteria as Cr
    pred, gt = synth.config_inputs("C1")
    pred[:4] *= 1.3
    for name in ("berhu", "silog"):
        full = float(_make(Cr, name)(pred.cuda(), gt.cuda()))
        local = 0.5 * (float(_make(Cr, name)(pred[:4].cuda(), gt[:4].cuda())) + float(_make(Cr, name)(pred[4:].cuda(), gt[4:].cuda())))
        assert abs(local - full) > 1e-4 * abs(full), name


def test_global_batch_loss_single_process_and_amp():
    """Without a process group global_batch_loss is the full-batch loss; half-precision predictions keep an fp32 stash."""
    from mono_depth_estimation_b200 import criteria as Cr, distributed as D
    pred, gt = synth.depth_pair((3, 1, 64, 96), 17, border=2)
    l64, g64 = olosses.loss_and_grad(olosses.silog, pred.double(), gt.double(), 0.85)
    p = pred.cuda().requires_grad_(True)
    loss = D.global_batch_loss(Cr.silog_loss(0.85), p, gt.cuda())
    (loss * 3.0).backward()
    close(loss, l64, LOSS_RTOL)
    grad_close(p.grad / 3.0, g64)
    ph = pred.half().cuda().requires_grad_(True)
    lh = D.global_batch_loss("l1", ph, gt.cuda())
    (lh * 65536.0).backward()
    assert ph.grad.dtype == torch.float16 and float((ph.grad != 0).float().mean()) > 0.5
