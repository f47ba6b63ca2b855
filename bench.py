#!/usr/bin/env python
"""bench.py - Mpix/s of the depth supervision + evaluation hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2|C1|C5]

A STEP is one pass of the hot path over one batch: the BTS configuration of BASELINE.json
(configs[1]) = silog_loss forward+backward + the reference's default metric list on a
[16,1,480,640] batch (4 915 200 px), through the package's public, reference-shaped API.
Rank 0 prints ONE JSON line (see the keys below). N > 1 is launched by torchrun, one process per
GPU: every rank runs the same per-GPU batch (weak scaling, images shard with no data-path
collective) plus the tiny all-reduce of the metric raw sums.

  value      whole-job Mpix/s, inputs resident in HBM, steps replayed from a CUDA graph (one graph
             launch per step) over a ring of distinct batches larger than L2
  e2e        same metric through the public API with HOST (pinned) inputs: H2D copy of pred+gt and
             a D2H read of the loss + metric values inside the timed region, every step
  roofline   dominant kernel: algorithmic bytes per launch / its average duration (CUDA events around
             back-to-back launches over the same ring) vs the measured HBM peak
  cpu_baseline  the CPU oracle port of the reference path timed on this box's host cores
--impl reference times that CPU path as its own arm (the reference is pure PyTorch-on-CPU here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TRAIN_METRICS = ["delta1", "delta2", "delta3", "mse", "mae", "log10", "rmse"]  # reference train.py:67 minus ssim
# algorithmic bytes per pixel (loss fwd+bwd fused with metrics: 12 B/px)
BYTES_PER_PX = {"silog_metrics_fused": 12.0, "silog_fwd_bwd": 12.0, "metrics": 8.0}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(shape):
    """The reference's CPU path for one step, as restated by the pinned oracle (oracle/)."""
    from oracle import losses as olosses, metrics as ometrics
    from mono_depth_estimation_b200 import synth
    pred, gt = synth.depth_pair(shape, synth.SEEDS["C2"])

    def step():
        loss, grad = olosses.loss_and_grad(olosses.silog, pred, gt, 0.85)
        vals = ometrics.compute(pred, gt, TRAIN_METRICS)
        return float(loss) + float(vals[0])
    return step, pred.numel()


def time_cpu(shape, steps, warmup, budget_s=25.0):
    torch.set_num_threads(os.cpu_count() or 1)
    step, npx = cpu_step_fn(shape)
    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    return {"value": npx / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of the full %s batch (silog fwd+bwd + %d metrics), fp32, torch CPU ops"
                      % (len(times), "x".join(map(str, shape)), len(TRAIN_METRICS)),
            "ms_per_step": ms, "steps": len(times)}


def run_reference_arm(args, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = time_cpu(shape, args.steps, min(args.warmup, 3))
    line = {"impl": "reference", "metric": "Mpix/s, depth supervision+eval step (silog fwd+bwd + metrics)",
            "value": cb["value"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": cb["steps"], "warmup": min(args.warmup, 3),
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: BTS SILog loss fwd+bwd + metrics, batch %s" % "x".join(map(str, shape)),
                       "metrics": TRAIN_METRICS, "device": "host CPU"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner to
# stdout at init): file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one.
_REAL_STDOUT = None


def _guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--ring", type=int, default=8, help="distinct batches cycled through (ring > L2)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="separate loss and metrics launches (20 B/px)")
    ap.add_argument("--sync-every", type=int, default=50, help="N>1: all-reduce the metric sums every this many steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager-gpu", action="store_true",
                    help="also time the reference's op chain (the oracle port: the same ATen sequence) run EAGERLY on this GPU; "
                         "informative second baseline of SURVEY 8(d), off by default")
    args = ap.parse_args()
    shape = (args.batch, 1, 480, 640)
    if args.impl == "reference":
        run_reference_arm(args, shape)
        return

    from mono_depth_estimation_b200 import _lib, criteria, metrics, synth
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    W = max(args.warmup, 3)
    K = args.steps
    npx = shape[0] * shape[2] * shape[3]

    # ring of distinct device batches (pred+gt = 39 MB each; 8 of them = 315 MB >> 126 MB L2)
    ring = [synth.depth_pair(shape, synth.SEEDS["C2"] + 1000 * rank + i, device=dev) for i in range(args.ring)]
    mcomp = metrics.MetricComputation(TRAIN_METRICS, strict=False)
    # the criterion's launch also produces the metric suite (one read of pred/gt for the whole step)
    crit = criteria.silog_loss(0.85).fuse_metrics(None if args.unfused else mcomp)
    raw_acc = torch.zeros(_lib.METRIC_NQ, dtype=torch.float64, device=dev)

    def step(pred, gt, pool=True):
        p = pred.detach().requires_grad_(True)
        loss = crit(p, gt)                       # reference modules/bts.py:106
        loss.backward()
        vals = mcomp.compute(p.detach(), gt)     # reference metrics.py:16-17 (log_train)
        raw = mcomp.last_f64[2 * _lib.METRIC_NM:2 * _lib.METRIC_NM + _lib.METRIC_NQ]
        if world > 1 and pool:  # pooled raw sums of this rank, accumulated on the device between exchanges
            raw_acc.add_(raw)
        return loss, vals, p.grad, raw

    pending = []

    def exchange():
        # The ONLY inter-GPU traffic of the path: 12 doubles (pooled metric sums and exact counts) summed
        # over the ranks, once per logging interval. The loss itself is local, as under the reference's
        # DDP (pl.Trainer(gpus=N), train.py:137) - which never synchronises metrics at all (no sync_dist,
        # metrics.py:19-39). Issued eagerly, outside the CUDA graphs.
        # Asynchronous: the sums are snapshotted on the step stream and reduced on NCCL's own stream,
        # the steps go on meanwhile; run_steps waits for every outstanding reduction before it returns (inside the
        # timed region).
        if world > 1:
            snap = raw_acc.clone()
            raw_acc.zero_()
            pending.append((dist.all_reduce(snap, async_op=True), snap))

    side = torch.cuda.Stream(device=dev)
    graphs = None
    mega = None          # one graph holding a whole pass over the ring (args.ring steps): no graph boundary between steps
    launches_per_step = None
    with torch.cuda.stream(side):
        for i in range(3):
            out = step(*ring[i % args.ring])
        side.synchronize()
        n0 = _lib.launch_count()
        out = step(*ring[0])
        side.synchronize()
        launches_per_step = _lib.launch_count() - n0
        if not args.no_graph:
            try:
                graphs = []
                for i in range(args.ring):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        o = step(*ring[i])
                    graphs.append((g, o))
                mega = torch.cuda.CUDAGraph()
                with torch.cuda.graph(mega, stream=side):
                    mega_out = [step(*ring[i], pool=False) for i in range(args.ring)]
                    if world > 1:   # the pass's pooled sums in one go (3 small launches per pass instead of one per step)
                        raw_acc.add_(torch.stack([o[3] for o in mega_out]).sum(0))
            except Exception as e:  # capture unsupported (e.g. a collective that cannot be captured) -> eager steps
                sys.stderr.write("graph capture failed (%s); timing eager steps\n" % (str(e).splitlines()[0],))
                graphs = None
                mega = None
                torch.cuda.synchronize()
        if world > 1:  # every rank must take the same path
            flag = torch.tensor([1 if graphs is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                graphs = None
                mega = None

    def run_steps(n, first=0):
        """Exactly n steps. Whole passes over the ring are replayed as one graph, the rest step by step."""
        i = 0
        since = 0
        while i < n:
            j = (first + i) % args.ring
            if mega is not None and j == 0 and n - i >= args.ring:
                mega.replay()
                done = args.ring
            elif graphs is not None:
                graphs[j][0].replay()
                done = 1
            else:
                step(*ring[j])
                done = 1
            i += done
            since += done
            if since >= args.sync_every:
                exchange()
                since = 0
        exchange()
        for work, _ in pending:
            work.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(side):
        run_steps(W)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            barrier()
            e0.record(side)
            run_steps(K, W)
            e1.record(side)
            barrier()
        ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t)
    ms_per_step = ms_total / K
    value = world * npx / (ms_per_step * 1e-3) / 1e6

    # ---- per-kernel timing (C ABI, preallocated outputs, same ring) -> roofline -----------------------
    peak, peak_src = measured_peak_gbs()
    kern = {}
    with torch.cuda.stream(side):
        lib = _lib.load()
        import ctypes as C
        ws = _lib.workspace(dev, shape[0])
        loss_t = torch.empty((), dtype=torch.float32, device=dev)
        grad_t = torch.empty(shape, dtype=torch.float32, device=dev)
        out64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
        out32 = torch.empty(2 * _lib.METRIC_NM, dtype=torch.float32, device=dev)
        lp = _lib.LossParams(0.85, 1e-9, 1, 1)
        sp = _lib.stream_ptr(dev)
        mflags = 0
        for n in TRAIN_METRICS:
            mflags |= _lib.METRIC_GROUP.get(n, 0)

        def k_fused(i):
            pr, g = ring[i % args.ring]
            _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], shape[2],
                                                   shape[3], C.byref(lp), 1.0, mflags, _lib.ptr(ws), _lib.ptr(loss_t), None,
                                                   _lib.ptr(grad_t), _lib.ptr(out64), _lib.ptr(out32), sp))

        def k_silog(i):
            pr, g = ring[i % args.ring]
            _lib.check(lib.mde_masked_loss(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], shape[2], shape[3],
                                           C.byref(lp), 1.0, _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grad_t), sp))

        def k_metrics(i):
            pr, g = ring[i % args.ring]
            _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(g), shape[0], shape[2] * shape[3], mflags, _lib.ptr(ws),
                                       _lib.ptr(out64), _lib.ptr(out32), None, None, sp))

        for name, fn in (("silog_metrics_fused", k_fused), ("silog_fwd_bwd", k_silog), ("metrics", k_metrics)):
            # one CUDA graph = one launch per ring slot, so the measurement has no host launch cost in it
            for i in range(W):
                fn(i)
            side.synchronize()
            kg = None
            if not args.no_graph:
                try:
                    kg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(kg, stream=side):
                        for i in range(args.ring):
                            fn(i)
                except Exception:
                    kg = None
                    torch.cuda.synchronize()
            reps = max(1, K // args.ring)
            if kg is not None:
                kg.replay()
            side.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            for r in range(reps):
                if kg is not None:
                    kg.replay()
                else:
                    for i in range(args.ring):
                        fn(i)
            b.record(side)
            side.synchronize()
            us = 1e3 * a.elapsed_time(b) / (reps * args.ring)
            gbs = BYTES_PER_PX[name] * npx / (us * 1e-6) / 1e9
            kern[name] = {"us_per_launch": us, "algorithmic_bytes": BYTES_PER_PX[name] * npx, "achieved_gbs": gbs,
                          "frac_of_peak": gbs / peak, "launches_timed": reps * args.ring, "graph": kg is not None}
    dom = "silog_fwd_bwd" if args.unfused else "silog_metrics_fused"   # the kernel the timed step launches
    traffic = None   # DRAM bytes per launch of that kernel from the committed ncu --set full capture (C2 batch only)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if dom in tj and args.batch == 16:
            traffic = float(tj[dom]["dram_read_bytes"] + tj[dom]["dram_write_bytes"])
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8000": kern[dom]["achieved_gbs"] / 8000.0, "kernels": kern,
                "note": "launches replayed from a CUDA graph (no host launch cost); per-launch ncu times and the traffic capture are in profiles/r01_ncu_summary_final.md"}

    # ---- e2e: public API, host (pinned) inputs, H2D + D2H inside the timed region -----------------------
    hp, hg = [], []
    for i in range(2):
        pr, g = ring[i]
        hp.append(pr.detach().cpu().pin_memory()); hg.append(g.cpu().pin_memory())
    # Two device buffer pairs: the copy of step i+1 (its own stream, pinned source) overlaps the compute and the result
    # read of step i, as a DataLoader with pin_memory + non_blocking copies does. Every step's inputs cross PCIe once,
    # inside the timed region.
    dbuf = [(torch.empty(shape, device=dev), torch.empty(shape, device=dev)) for _ in range(2)]
    res_host = torch.empty(1 + len(TRAIN_METRICS), dtype=torch.float32).pin_memory()
    mcomp_e2e = mcomp
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_prefetch(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])                     # the step that last used this pair is done with it
            dbuf[b][0].copy_(hp[b], non_blocking=True)
            dbuf[b][1].copy_(hg[b], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_step(i, last):
        b = i % 2
        cur = torch.cuda.current_stream(dev)
        if not last:
            e2e_prefetch(i + 1)
        cur.wait_event(ready[b])
        dp, dg = dbuf[b]
        p = dp.detach().requires_grad_(True)
        loss = crit(p, dg)
        loss.backward()
        vals = mcomp_e2e.compute(p.detach(), dg)
        out = torch.stack([loss.detach()] + vals)
        freed[b].record(cur)
        res_host.copy_(out, non_blocking=False)                  # D2H read of the step's result
        return float(res_host[0])

    Ke = max(5, min(K, 50))
    for b in range(2):
        freed[b].record(torch.cuda.current_stream(dev))
    e2e_prefetch(0)
    for i in range(3):
        e2e_step(i, last=(i == 2))
    barrier()
    t0 = time.perf_counter()
    e2e_prefetch(0)
    for i in range(Ke):
        e2e_step(i, last=(i == Ke - 1))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(dt) / Ke
    e2e = {"value": world * npx / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 2 * 4 * npx, "d2h_bytes_per_step": 4 * (1 + len(TRAIN_METRICS)), "steps": Ke}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = time_cpu(shape, 12, 2)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    eager_gpu = None
    if rank == 0 and world == 1 and args.eager_gpu:
        # the reference's own way of running this step on a GPU: ~40 ATen launches with boolean-mask gathers (each a
        # device->host sync for the output size). Baseline only: nothing of the product path is involved.
        from oracle import losses as olosses, metrics as ometrics
        def eager_step(i):
            pr, g = ring[i % args.ring]
            loss, grad = olosses.loss_and_grad(olosses.silog, pr, g, 0.85)
            vals = ometrics.compute(pr, g, TRAIN_METRICS)
            return float(loss) + float(vals[0])
        for i in range(3):
            eager_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ne = 20
        for i in range(ne):
            eager_step(i)
        torch.cuda.synchronize()
        eg_ms = 1e3 * (time.perf_counter() - t0) / ne
        eager_gpu = {"value": npx / (eg_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": eg_ms, "steps": ne,
                     "kind": "oracle port of the reference's ATen chain, eager on cuda:0, inputs resident"}

    if rank == 0:
        line = {"metric": "Mpix/s, depth supervision+eval step (silog fwd+bwd + metrics)", "value": value, "unit": "Mpix/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: BTS SILog loss fwd+bwd + metrics, batch %s per GPU" % "x".join(map(str, shape)),
                           "metrics": TRAIN_METRICS, "variance_focus": 0.85,
                           "l2_policy": "ring of %d distinct batches (%.0f MB) larger than the 126 MB L2" %
                                        (args.ring, args.ring * 2 * 4 * npx / 1e6),
                           "cuda_graph": graphs is not None, "steps_per_graph": (args.ring if mega is not None else 1), "parallelism": "image-sharded x%d" % world,
                           "collective": None if world == 1 else "all-reduce of 12 doubles every %d steps (NCCL)" % args.sync_every},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches_per_step) * K,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        if eager_gpu is not None:
            line["eager_gpu_baseline"] = eager_gpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
