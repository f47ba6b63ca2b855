#!/usr/bin/env python
"""bench.py - Mpix/s of the depth supervision + evaluation hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A STEP is one pass of the hot path over one batch: the BTS configuration of BASELINE.json
(configs[1], "C2") = silog_loss forward+backward + the reference's default train metrics on a
[16,1,480,640] batch (4 915 200 px), through the package's public, reference-shaped API.
Rank 0 prints ONE JSON line. N > 1 is launched by torchrun, one process per GPU.

  value      whole-job Mpix/s of the C2 step, inputs resident in HBM, steps replayed from CUDA graphs over a ring
             of distinct batches larger than L2. Weak scaling: every rank runs the per-GPU batch, images shard
             with no data-path collective; the metric sums are all-reduced every --sync-every steps.
             The K timed steps are repeated R >= 50 times, each repetition bracketed by a barrier +
             synchronize and timed with CUDA events, max over ranks per repetition, MEDIAN over repetitions.
  e2e        same metric through the public API with HOST (pinned) inputs: H2D copy of pred+gt and a D2H read
             of the loss + metric values inside the timed region, every step
  roofline   the kernel the step launches: algorithmic bytes per launch / its average duration (CUDA events,
             ring of input AND gradient buffers) vs the measured HBM peak
  c5_eval    BASELINE.json configs[4]: the 654-image NYU-test-shaped evaluation SHARDED over the N ranks
             (82/81 images per rank at N=8), one launch + one 25-double all-reduce per evaluation: strong scaling
  configs    (N=1) per-config kernel numbers: C1 berHu+metrics, C3 DORN fused, C4 VNL, C5 metrics
  cpu_baseline  (N=1) the REFERENCE's own criteria.py / metrics.py timed on this box's host cores (subprocess
             with CUDA hidden; oracle/ref_step.py); `kind: "port"` only if the reference copy is absent
--impl reference runs that CPU path as its own arm with the same config / steps / warm-up.
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
EVAL_METRICS = TRAIN_METRICS + ["absrel", "sqrel", "msle"]
METRIC_NAME = "Mpix/s, depth supervision+eval step (silog fwd+bwd + metrics)"
# algorithmic bytes per pixel (loss fwd+bwd fused with metrics: 12 B/px)
BYTES_PER_PX = {"silog_metrics_fused": 12.0, "silog_fwd_bwd": 12.0, "metrics": 8.0}


def workload_name(shape):
    return "C2: BTS SILog loss fwd+bwd + metrics, batch %s per GPU" % "x".join(map(str, shape))


def pin_to_gpu_cpus(index):
    """Bind this process to the CPUs NVML reports as local to GPU `index` (the cores of its NUMA node): the pinned staging
    buffers of the e2e leg are then allocated next to the GPU's PCIe root, and with several ranks per box every rank stays
    on its own side. Returns what was done (part of the e2e object). MDE_BENCH_NO_AFFINITY=1 leaves the process alone."""
    info = {"applied": False}
    if os.environ.get("MDE_BENCH_NO_AFFINITY", "0") != "0" or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        info.update({"gpu_cpus": len(cpus), "allowed_before": len(allowed)})
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["applied"] = True
        info["allowed_now"] = len(os.sched_getaffinity(0))
    except Exception as e:   # NVML missing or the container forbids it: measured as is
        info["error"] = str(e).splitlines()[0][:100]
    return info


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
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
def time_cpu(batch, steps, warmup, budget_s):
    """The reference's CPU path for the C2 step on this box's host cores (oracle/ref_step.py in a subprocess with CUDA
    hidden: the reference moves work to `cuda` whenever it is visible)."""
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "oracle.ref_step", "--batch", str(batch), "--steps", str(steps), "--warmup", str(warmup),
           "--budget", str(budget_s), "--names", ",".join(TRAIN_METRICS)]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=min(budget_s * 6 + 600, 3600))
    if r.returncode != 0:
        raise RuntimeError("oracle.ref_step failed: " + r.stderr[-800:])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    ms = d["ms_per_step"]
    what = "the reference's own criteria.py + metrics.py (unmodified copy, loaded by path)" if d["kind"] == "reference" else \
           "the oracle port of the reference's ATen chain (no reference copy on this box)"
    return {"value": d["npx"] / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "cores": d["cores"], "kind": d["kind"],
            "sample": "%d steps of the full %dx1x480x640 batch (silog fwd+bwd + %d metrics), fp32, torch CPU ops, %s"
                      % (d["steps"], batch, len(TRAIN_METRICS), what),
            "ms_per_step": ms, "steps": d["steps"]}


def run_reference_arm(args, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = max(args.warmup, 3)
    # a reference step takes ~1-2 s on the host cores: up to 50 steps run in full, more are cut by the time budget
    cb = time_cpu(shape[0], args.steps, W, budget_s=120.0 if args.steps > 50 else 1e6)
    line = {"impl": "reference", "metric": METRIC_NAME,
            "value": cb["value"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": W,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(shape), "metrics": TRAIN_METRICS, "variance_focus": 0.85,
                       "device": "host CPU", "steps_timed": cb["steps"]},
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


def graph_timed(fns, dev, reps):
    """us per call of a list of zero-argument callables (one per ring slot), replayed from ONE CUDA graph."""
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        for f in fns:
            f()
        side.synchronize()
        g = None
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for f in fns:
                    f()
        except Exception as e:
            sys.stderr.write("graph capture failed: %s\n" % str(e).splitlines()[0])
            g = None
            torch.cuda.synchronize()

        def once():
            if g is not None:
                g.replay()
            else:
                for f in fns:
                    f()
        for _ in range(3):
            once()
        side.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            once()
            b.record(side)
            side.synchronize()
            ts.append(1e3 * a.elapsed_time(b) / len(fns))
    return statistics.median(ts), g is not None


def config_rows(dev, peak, quick=False):
    """Per-config kernel numbers (N=1): us per launch over inputs larger than L2, algorithmic bytes, fraction of
    the measured HBM peak. C ABI calls with preallocated outputs, replayed from CUDA graphs."""
    import ctypes as C
    from mono_depth_estimation_b200 import _lib, criteria, metrics, synth
    lib = _lib.load()
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731
    out = {}
    reps = 5 if quick else 20

    def row(name, what, us, alg_bytes, extra=None):
        gbs = alg_bytes / (us * 1e-6) / 1e9
        d = {"what": what, "us_per_launch": round(us, 2), "algorithmic_bytes": alg_bytes, "achieved_gbs": round(gbs, 1),
             "frac": round(gbs / peak, 4)}
        if extra:
            d.update(extra)
        out[name] = d

    loss_t = torch.empty((), device=dev)
    # ---- C1: berHu fwd+bwd + 7 metrics in one launch, 8x1x228x304 (6.6 MB per call: a latency case, not a bandwidth case)
    shape = synth.SHAPES["C1"]
    px = shape[0] * shape[2] * shape[3]
    ring = [synth.depth_pair(shape, 101 + i, device=dev) for i in range(48)]      # 48 x 4.4 MB > L2
    grads = [torch.empty(shape, device=dev) for _ in range(48)]
    ws = _lib.workspace(dev, shape[0])
    o64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev); o32 = torch.empty(24, device=dev)
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    mflags = 0
    for n in TRAIN_METRICS:
        mflags |= _lib.METRIC_GROUP.get(n, 0)
    for kname, kind in (("C1", _lib.LOSS_BERHU), ("C1_silog", _lib.LOSS_SILOG), ("C1_l1", _lib.LOSS_L1), ("C1_laina", _lib.LOSS_LAINA_BERHU)):
        fns = [lambda pr=pr, gt=gt, gr=gr, kind=kind: _lib.check(lib.mde_masked_loss_metrics(
            kind, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0, mflags, _lib.ptr(ws),
            _lib.ptr(loss_t), None, _lib.ptr(gr), _lib.ptr(o64), _lib.ptr(o32), sp())) for (pr, gt), gr in zip(ring, grads)]
        us, _ = graph_timed(fns, dev, reps)
        row(kname, {"C1": "berHu", "C1_silog": "SILog", "C1_l1": "L1", "C1_laina": "Laina berHu"}[kname] +
            " fwd+bwd + 7 metrics, one launch (register-resident kernel), 8x1x228x304", us, 12.0 * px)
    # MaskedDepthLoss (the reference's C1 criterion of the eigen module, criteria.py:17-64): fwd+bwd, one launch
    fns = [lambda pr=pr, gt=gt, gr=gr: _lib.check(lib.mde_masked_loss(
        _lib.LOSS_EIGEN, _lib.ptr(pr), 0, _lib.ptr(gt), None, shape[0], shape[2], shape[3], C.byref(lp), 1.0, _lib.ptr(ws),
        _lib.ptr(loss_t), None, _lib.ptr(gr), sp())) for (pr, gt), gr in zip(ring, grads)]
    us, _ = graph_timed(fns, dev, reps)
    row("C1_eigen", "MaskedDepthLoss (Eigen scale-invariant + gradient term) fwd+bwd, one launch (register-resident kernel), "
        "8x1x228x304", us, 12.0 * px)
    del ring, grads

    # ---- C3: DORN fused logits -> decode, depth, ordinal loss, grad (K = 68), 8x136x257x353
    shape = synth.SHAPES["C3"] if not quick else (2, 136, 257, 353)
    N, C2, H, W = shape
    px = N * H * W
    ring = [synth.dorn_inputs(shape, 103 + i, device=dev) for i in range(2)]        # 395 MB of logits each
    dec = torch.empty((N, 1, H, W), dtype=torch.int64, device=dev)
    dep = torch.empty((N, 1, H, W), device=dev)
    gxs = [torch.empty(shape, device=dev) for _ in range(2)]
    ws1 = _lib.workspace(dev, 1)
    fns = [lambda x=x, gt=gt, gx=gx: _lib.check(lib.mde_dorn_fused(_lib.ptr(x), 0, _lib.ptr(gt), N, C2 // 2, H * W, 0.001, 1.0, 0, 1.0,
                                                               _lib.ptr(ws1), _lib.ptr(loss_t), None, _lib.ptr(dec), _lib.ptr(dep),
                                                               _lib.ptr(gx), sp())) for (x, gt), gx in zip(ring, gxs)]
    us, _ = graph_timed(fns, dev, reps)
    row("C3", "DORN fused: logits -> decode + depth + ordLoss + grad, K=68, %s" % "x".join(map(str, shape)), us, (8.0 * C2 + 16.0) * px,
        {"mpix_s": round(px / us, 1)})
    fns = [lambda x=x: _lib.check(lib.mde_ordinal_layer_fwd(_lib.ptr(x), 0, N, C2 // 2, H * W, None, _lib.ptr(dec), sp())) for x, _ in ring]
    us, _ = graph_timed(fns, dev, reps)
    row("C3_decode", "DORN decode only (inference)", us, (4.0 * C2 + 8.0) * px, {"mpix_s": round(px / us, 1)})
    del ring, gxs

    # ---- C4: VNL fwd+bwd, 100k triplets x 8 images at 385x385 (bound by L2 gathers + atomics, not HBM)
    shape = synth.SHAPES["C4"]
    gt, pred, trip = synth.vnl_inputs(shape, 104, device=dev)
    B, _, H, W = shape
    px = B * H * W
    n_trip = trip.shape[1]
    wsb = _lib.workspace(dev, B)
    scratch = torch.empty(int(lib.mde_vnl_scratch_bytes(B, n_trip, H, W)), dtype=torch.uint8, device=dev)
    stats = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.empty_like(pred)
    fns = [lambda: _lib.check(lib.mde_vnl_loss(_lib.ptr(gt), _lib.ptr(pred), 0, _lib.ptr(trip), B, H, W, n_trip, 519.0, 519.0, 1, 1.0,
                                               _lib.ptr(wsb), _lib.ptr(scratch), _lib.ptr(loss_t), _lib.ptr(stats), _lib.ptr(grad), sp()))]
    us, _ = graph_timed(fns, dev, reps)
    row("C4", "VNL fwd+bwd, 100k triplets x 8 images, 385x385", us, 12.0 * px + 24.0 * n_trip,
        {"mtriplets_s": round(B * n_trip / us, 1), "bound": "L2 gathers + atomics + rank select, not HBM"})
    del gt, pred, trip, grad

    # ---- C5: 654 x 480 x 640 evaluation metrics (one launch), default 7 and all 10 keys
    B = 654 if not quick else 64
    px = B * 480 * 640
    pr, gtt = synth.depth_pair((B, 1, 480, 640), 105, device=dev)
    ws5 = _lib.workspace(dev, B)
    for kname, names in (("C5", TRAIN_METRICS), ("C5_10", EVAL_METRICS)):
        fl = 0
        for n in names:
            fl |= _lib.METRIC_GROUP.get(n, 0)
        fns = [lambda fl=fl: _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(gtt), B, 480 * 640, fl, _lib.ptr(ws5), _lib.ptr(o64),
                                                        _lib.ptr(o32), None, None, sp()))]
        us, _ = graph_timed(fns, dev, reps)
        row(kname, "eval metrics over %d images 480x640, %d keys, one launch" % (B, len(names)), us, 8.0 * px, {"mpix_s": round(px / us, 1)})
    del pr, gtt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--ring", type=int, default=8, help="distinct batches cycled through (ring > L2)")
    ap.add_argument("--repeats", type=int, default=0, help="repetitions of the K timed steps (0: at least 50, at most ~0.5 s)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="separate loss and metrics launches (20 B/px)")
    ap.add_argument("--sync-every", type=int, default=50, help="N>1: all-reduce the metric sums every this many steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config kernel table (C1, C3, C4, C5)")
    ap.add_argument("--no-eager-gpu", action="store_true", help="skip the reference's op chain run eagerly on this GPU")
    args = ap.parse_args()
    shape = (args.batch, 1, 480, 640)
    if args.impl == "reference":
        run_reference_arm(args, shape)
        return

    from mono_depth_estimation_b200 import _lib, criteria, distributed as mdist, metrics, synth
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_cpus(local)        # before any pinned host buffer is allocated (first touch decides its NUMA node)
    if world > 1:
        import datetime
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
    W = max(args.warmup, 3)
    K = args.steps
    npx = shape[0] * shape[2] * shape[3]

    # ring of distinct device batches (pred+gt = 39 MB each; 8 of them = 315 MB >> 126 MB L2)
    ring = [synth.depth_pair(shape, synth.SEEDS["C2"] + 1000 * rank + i, device=dev) for i in range(args.ring)]
    mcomp = metrics.MetricComputation(TRAIN_METRICS, strict=False)
    # the criterion's launch also produces the metric suite (one read of pred/gt for the whole step) and books it
    # into the computer's running sums, as the log_train() that follows every criterion call would (metrics.py:16-17)
    crit = criteria.silog_loss(0.85).fuse_metrics(None if args.unfused else mcomp, book=True)
    # N > 1: the pooled raw sums of this rank between two exchanges are accumulated by the same launch (no launch of their own)
    raw_acc = mcomp.raw_accum_buffer(dev) if (world > 1 and not args.unfused) else torch.zeros(_lib.METRIC_NQ, dtype=torch.float64, device=dev)

    def step(pred, gt, pool=True):
        p = pred.detach().requires_grad_(True)
        loss = crit(p, gt)                       # reference modules/bts.py:106
        loss.backward()
        vals = mcomp.compute(p.detach(), gt)     # reference metrics.py:16-17 (log_train)
        raw = mcomp.last_f64[2 * _lib.METRIC_NM:2 * _lib.METRIC_NM + _lib.METRIC_NQ]
        if world > 1 and args.unfused and pool:
            raw_acc.add_(raw)
        return loss, vals, p.grad, raw

    # N > 1: whatever the ranks exchange goes through NVLink peer mailboxes (distributed.PeerComm): the 25 doubles of a
    # sharded evaluation INSIDE its launch (mde_metrics_sharded), the 12 pooled metric sums of the training steps through a
    # one-warp launch (mde_peer_allreduce_f64). If the peer mappings cannot be set up on this box, the process-group path
    # (NCCL) is timed instead
    comm = None
    comm_note = None
    if world > 1 and os.environ.get("MDE_BENCH_NO_PEER", "0") == "0":
        try:
            comm = mdist.PeerComm()
        except Exception as e:
            comm_note = "PeerComm unavailable (%s)" % (str(e).splitlines()[0][:120],)
            comm = None
        flag = torch.tensor([1 if comm is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks take the same path
        if int(flag) == 0:
            comm = None

    pending = []

    def exchange():
        # The ONLY inter-GPU traffic of the C2 path: 12 doubles (pooled metric sums and exact counts) summed over the
        # ranks once per logging interval. The loss itself is local, as under the reference's DDP
        # (pl.Trainer(gpus=N), train.py:137) - which never synchronises metrics at all (no sync_dist, metrics.py:19-39).
        # Asynchronous: the sums are snapshotted on the step stream and reduced on NCCL's own stream while the
        # steps go on; everything outstanding is waited for before the run ends.
        if world > 1 and comm is not None:
            # one one-warp launch on the step stream: read the accumulator, exchange through the mailboxes, write the sum,
            # clear the accumulator (a collective's kernel would hold an SM for tens of microseconds while the
            # cooperative step kernel, which needs every SM, waits behind it: 1.2 us per step at N = 8)
            snap = torch.empty_like(raw_acc)
            comm.all_reduce_(raw_acc, out=snap, zero_src=True)
            pending.append((None, snap))
            if len(pending) > 64:
                del pending[:32]
        elif world > 1:
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
                    if world > 1 and args.unfused:   # the pass's pooled sums in one go
                        raw_acc.add_(torch.stack([o[3] for o in mega_out]).sum(0))
            except Exception as e:  # capture unsupported -> eager steps
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

    since = [0]

    def run_steps(n, first=0):
        """Exactly n steps. Whole passes over the ring are replayed as one graph, the rest step by step."""
        i = 0
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
            since[0] += done
            if since[0] >= args.sync_every:
                exchange()
                since[0] = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(side):
        # clocks: ~150 ms of the same steps before anything is timed (a B200 idles at 120 MHz). A FIXED number of passes:
        # every rank must issue the same sequence of collectives (run_steps exchanges every --sync-every steps)
        for _ in range(750):
            run_steps(args.ring)
        side.synchronize()
        run_steps(W)
        barrier()
        # R repetitions of exactly K steps; a repetition that is short in wall time is launch-jitter territory, so R grows
        # until ~0.5 s of steps are timed (>= 50 repetitions, <= 2000)
        est_ms = K * 0.03
        R = args.repeats if args.repeats > 0 else int(min(2000, max(50, 500.0 / max(est_ms, 1e-3))))
        rep_ms = []
        with ClockSampler(local) as clocks:
            first = W
            for r in range(R):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                run_steps(K, first)
                e1.record(side)
                barrier()
                rep_ms.append(e0.elapsed_time(e1))
                first = (first + K) % args.ring
            exchange()
            for work, _ in pending:
                if work is not None:
                    work.wait()
            pending.clear()
            torch.cuda.synchronize()
    t = torch.tensor(rep_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # per repetition: the slowest rank
    rep_ms = t.cpu().tolist()
    ms_total = statistics.median(rep_ms)
    ms_per_step = ms_total / K
    value = world * npx / (ms_per_step * 1e-3) / 1e6
    timing = {"repetitions": R, "ms_per_step_median": ms_per_step, "ms_per_step_min": min(rep_ms) / K, "ms_per_step_mean": sum(rep_ms) / len(rep_ms) / K,
              "ms_per_step_p90": sorted(rep_ms)[int(0.9 * (len(rep_ms) - 1))] / K}

    # ---- per-kernel timing (C ABI, preallocated outputs, ring of input AND gradient buffers) -> roofline ---------
    peak, peak_src = measured_peak_gbs()
    kern = {}
    import ctypes as C
    lib = _lib.load()
    grads = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(args.ring)]
    with torch.cuda.stream(side):
        ws = _lib.workspace(dev, shape[0])
    loss_t = torch.empty((), dtype=torch.float32, device=dev)
    out64 = torch.empty(_lib.METRICS_OUT_F64, dtype=torch.float64, device=dev)
    out32 = torch.empty(2 * _lib.METRIC_NM, dtype=torch.float32, device=dev)
    lp = _lib.LossParams(0.85, 1e-9, 1, 1)
    sp = lambda: _lib.stream_ptr(dev)  # noqa: E731
    mflags = 0
    for n in TRAIN_METRICS:
        mflags |= _lib.METRIC_GROUP.get(n, 0)

    def k_fused(i):
        pr, g = ring[i]
        _lib.check(lib.mde_masked_loss_metrics(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], shape[2],
                                               shape[3], C.byref(lp), 1.0, mflags, _lib.ptr(ws), _lib.ptr(loss_t), None,
                                               _lib.ptr(grads[i]), _lib.ptr(out64), _lib.ptr(out32), sp()))

    def k_silog(i):
        pr, g = ring[i]
        _lib.check(lib.mde_masked_loss(_lib.LOSS_SILOG, _lib.ptr(pr), 0, _lib.ptr(g), None, shape[0], shape[2], shape[3],
                                       C.byref(lp), 1.0, _lib.ptr(ws), _lib.ptr(loss_t), None, _lib.ptr(grads[i]), sp()))

    def k_metrics(i):
        pr, g = ring[i]
        _lib.check(lib.mde_metrics(_lib.ptr(pr), 0, _lib.ptr(g), shape[0], shape[2] * shape[3], mflags, _lib.ptr(ws),
                                   _lib.ptr(out64), _lib.ptr(out32), None, None, sp()))

    for name, fn in (("silog_metrics_fused", k_fused), ("silog_fwd_bwd", k_silog), ("metrics", k_metrics)):
        us, graphed = graph_timed([lambda i=i, fn=fn: fn(i) for i in range(args.ring)], dev, 30)
        gbs = BYTES_PER_PX[name] * npx / (us * 1e-6) / 1e9
        kern[name] = {"us_per_launch": us, "algorithmic_bytes": BYTES_PER_PX[name] * npx, "achieved_gbs": gbs,
                      "frac_of_peak": gbs / peak, "launches_timed": 30 * args.ring, "graph": graphed}
    del grads
    dom = "silog_fwd_bwd" if args.unfused else "silog_metrics_fused"   # the kernel the timed step launches
    traffic = None   # DRAM bytes per launch of that kernel from the committed ncu --set full capture (C2 batch only)
    traffic_src = None
    for tf in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if dom in tj and args.batch == 16:
                traffic = float(tj[dom]["dram_read_bytes"] + tj[dom]["dram_write_bytes"])
                traffic_src = "profiles/" + tf
                break
        except Exception:
            continue
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac_of_peak"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "frac_of_nominal_8000": kern[dom]["achieved_gbs"] / 8000.0, "kernels": kern,
                "note": "median of 30 replays of one CUDA graph = one launch per ring slot (distinct inputs AND gradient buffers, ring > L2)"}

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
    copy_stream = torch.cuda.Stream(device=dev)
    copy_stream2 = torch.cuda.Stream(device=dev) if os.environ.get("MDE_BENCH_ONE_COPY_STREAM", "0") == "0" else None
    ready2 = [torch.cuda.Event(), torch.cuda.Event()]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_prefetch(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])                     # the step that last used this pair is done with it
            dbuf[b][0].copy_(hp[b], non_blocking=True)
            if copy_stream2 is None:
                dbuf[b][1].copy_(hg[b], non_blocking=True)
            ready[b].record(copy_stream)
        if copy_stream2 is not None:                             # prediction and target travel on two copy engines
            with torch.cuda.stream(copy_stream2):
                copy_stream2.wait_event(freed[b])
                dbuf[b][1].copy_(hg[b], non_blocking=True)
                ready2[b].record(copy_stream2)

    def e2e_step(i, last):
        b = i % 2
        cur = torch.cuda.current_stream(dev)
        if not last:
            e2e_prefetch(i + 1)
        cur.wait_event(ready[b])
        if copy_stream2 is not None:
            cur.wait_event(ready2[b])
        dp, dg = dbuf[b]
        p = dp.detach().requires_grad_(True)
        loss = crit(p, dg)
        loss.backward()
        vals = mcomp.compute(p.detach(), dg)
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
    my_s = time.perf_counter() - t0
    dt = torch.tensor([my_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(dt) / Ke
    h2d = 2 * 4 * npx
    e2e = {"value": world * npx / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * (1 + len(TRAIN_METRICS)), "steps": Ke,
           "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "cpu_affinity": affinity, "h2d_gbs_aggregate": world * h2d / (e2e_ms * 1e-3) / 1e9,
           "note": "pinned host buffers; every rank copies its own 39 MB per step over PCIe - the aggregate figure is the host-side ceiling "
                   "the ranks share (SCALE topology: all GPUs of the box hang off one NUMA node)"}
    del dbuf, hp, hg

    # ---- C5: the 654-image evaluation sharded over the ranks (strong scaling; BASELINE.json configs[4]) -------------
    n_img_total = 654
    a_img, b_img = mdist.shard_range(n_img_total, rank, world)
    n_loc = b_img - a_img
    c5_pred, c5_gt = synth.depth_pair((max(n_loc, 1), 1, 480, 640), synth.SEEDS["C5"] + 7919 * rank, device=dev)
    if n_loc == 0:
        c5_pred, c5_gt = c5_pred[:0], c5_gt[:0]
    c5_px = n_img_total * 480 * 640

    def c5_eval(async_op):
        if comm is not None:
            return mdist.sharded_eval(c5_pred, c5_gt, EVAL_METRICS, comm=comm)
        return mdist.sharded_eval(c5_pred, c5_gt, EVAL_METRICS, all_reduce=True, async_op=async_op)

    Kc = max(3, min(K, 20))            # evaluations per repetition
    Rc = 30
    c5_ms = []
    c5_graph = None
    with torch.cuda.stream(side):
        for _ in range(3):
            c5_eval(False)
        side.synchronize()
        # one launch per evaluation and rank, the exchange inside it: the Kc evaluations of a repetition replay as ONE CUDA
        # graph (at N = 8 an evaluation takes ~55 us on the GPU, less than the Python call that launches it). The
        # process-group path (a collective per evaluation) is timed eagerly.
        if (comm is not None or world == 1) and not args.no_graph:
            try:
                c5_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(c5_graph, stream=side):
                    c5_keep = [c5_eval(False) for _ in range(Kc)]
            except Exception as e:
                sys.stderr.write("C5 graph capture failed (%s); timing eager launches\n" % (str(e).splitlines()[0],))
                c5_graph = None
                torch.cuda.synchronize()
        if world > 1:
            flag = torch.tensor([1 if c5_graph is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                c5_graph = None
        barrier()
        for r in range(Rc):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            if c5_graph is not None:
                c5_graph.replay()
            else:
                works = []
                for _ in range(Kc):
                    works.append(c5_eval(world > 1))      # (NCCL path: the all-reduce of evaluation i overlaps the launch of evaluation i+1)
                for wk in works:
                    if wk.get("work") is not None:
                        wk["work"].wait()                # the step stream waits for every reduction before the region closes
            e1.record(side)
            barrier()
            c5_ms.append(e0.elapsed_time(e1) / Kc)
        res = c5_eval(False)
        c5_check = {"n_images": float(res["n_images"]), "delta1_image_mean": float(res["image_mean"]["delta1"])}
    t = torch.tensor(c5_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c5_med = statistics.median(t.cpu().tolist())
    c5 = {"workload": "C5: NYU-test-shaped eval, 654 x 1x480x640, %d metrics, images sharded %s" %
                      (len(EVAL_METRICS), "/".join(str(mdist.shard_range(n_img_total, r, world)[1] - mdist.shard_range(n_img_total, r, world)[0])
                                                   for r in range(world))),
          "value": c5_px / (c5_med * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_eval": c5_med, "scaling": "strong",
          "evals_per_repetition": Kc, "repetitions": Rc, "cuda_graph": c5_graph is not None, "hbm_frac_per_gpu": 8.0 * c5_px / world / (c5_med * 1e-3) / 1e9 / peak,
          "collective": None if world == 1 else ("none: 25 doubles per rank exchanged inside the launch (finaliser stores into the peers' mailboxes over NVLink, tagged words)"
                                                  if comm is not None else "one all-reduce of 25 doubles per evaluation (NCCL, in place on the kernel's result vector)"),
          "collective_note": comm_note,
          "l2_policy": "per-rank shard %.0f MB > L2" % (n_loc * 2 * 4 * 480 * 640 / 1e6), "check": c5_check}
    del c5_pred, c5_gt
    if comm is not None:
        comm.close()

    cfgs = None
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            with torch.cuda.stream(side):
                pass
            cfgs = config_rows(dev, peak)
        except Exception as e:   # never lose the headline line to a side table
            cfgs = {"error": str(e).splitlines()[0][:200]}

    eager_gpu = None
    if rank == 0 and world == 1 and not args.no_eager_gpu:
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

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = time_cpu(shape[0], 12, 3, budget_s=25.0)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC_NAME, "value": value, "unit": "Mpix/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(shape),
                           "metrics": TRAIN_METRICS, "variance_focus": 0.85,
                           "l2_policy": "ring of %d distinct batches (%.0f MB) larger than the 126 MB L2" %
                                        (args.ring, args.ring * 2 * 4 * npx / 1e6),
                           "cuda_graph": graphs is not None, "steps_per_graph": (args.ring if mega is not None else 1),
                           "parallelism": "image-sharded x%d" % world,
                           "timing": "K steps x %d repetitions, each bracketed by barrier+synchronize, CUDA events, max over ranks per "
                                     "repetition, median over repetitions" % R,
                           "metrics_booking": "in the criterion's launch (fuse_metrics(book=True))",
                           "collective": None if world == 1 else ("sum of 12 doubles every %d steps through NVLink peer mailboxes (one one-warp launch, mde_peer_allreduce_f64)" % args.sync_every
                                                                   if comm is not None else "all-reduce of 12 doubles every %d steps (NCCL, async)" % args.sync_every)},
                "timing": timing, "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches_per_step) * K,
                "roofline": roofline, "c5_eval": c5, "cpu_baseline": cpu_baseline}
        if cfgs is not None:
            line["configs"] = cfgs
        if eager_gpu is not None:
            line["eager_gpu_baseline"] = eager_gpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
