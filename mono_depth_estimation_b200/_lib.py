"""ctypes binding of libmde_b200.so (the C ABI declared in include/mde_b200.h).

There is NO CPU fallback: if the shared library is missing or the tensors are not on a CUDA
device every entry point raises. PyTorch is used only for device memory, streams and autograd.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDE_B200_LIB") or os.path.join(_HERE, "libmde_b200.so")   # the override is for instrumented builds (tools/)

# enums of include/mde_b200.h
F32, F16, BF16 = 0, 1, 2
METRIC_NQ, METRIC_NM = 12, 12
METRICS_OUT_F64 = 3 * METRIC_NM + METRIC_NQ + 1
METRICS_REFERENCE_MATH = 1
METRICS_NEED_LOG, METRICS_NEED_LOG1P, METRICS_NEED_REL, METRICS_NEED_RSQ = 1 << 8, 1 << 9, 1 << 10, 1 << 11
LOSS_L1, LOSS_MSE, LOSS_BERHU, LOSS_LAINA_BERHU, LOSS_SILOG, LOSS_EIGEN = range(6)
LOSS_NTOTALS = 8
DISC_SID, DISC_UD = 0, 1

METRIC_INDEX = {"delta1": 0, "delta2": 1, "delta3": 2, "mae": 3, "mse": 4, "log10": 5, "msle": 6,
                "absrel": 7, "sqrel": 8, "rmse": 9, "rmse_true": 10, "rmse_log": 11}
METRIC_GROUP = {"log10": METRICS_NEED_LOG, "rmse_log": METRICS_NEED_LOG, "msle": METRICS_NEED_LOG1P,
                "absrel": METRICS_NEED_REL, "sqrel": METRICS_NEED_REL, "rmse": METRICS_NEED_RSQ}
MAX_PEERS, PEER_MAILBOX_BYTES, PEER_HANDLE_BYTES = 8, 8192, 64
RAW_INDEX = {"n_valid": 0, "d1": 1, "d2": 2, "d3": 3, "abs": 4, "sq": 5, "log10": 6, "sle": 7,
             "absrel": 8, "sqrel": 9, "rsq": 10, "lnsq": 11}


class LossParams(C.Structure):
    _fields_ = [("variance_focus", C.c_float), ("clamp_val", C.c_float),
                ("use_logs", C.c_int), ("size_average", C.c_int), ("metrics_accum", C.c_void_p),
                ("metrics_raw_accum", C.c_void_p)]


_vp, _i64, _i32, _u32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint, C.c_float

# name -> (restype, argtypes); every symbol include/mde_b200.h declares
SIGNATURES = {
    "mde_metrics": (_i32, [_vp, _i32, _vp, _i64, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mde_metrics_sharded": (_i32, [_vp, _i32, _vp, _i64, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "mde_peer_allreduce_f64": (_i32, [_vp, _vp, _i32, _i32, _vp, _u32, _vp, _vp]),
    "mde_peer_comm_create": (_i32, [C.POINTER(_vp), _i32, _i32, _u32, C.POINTER(_vp)]),
    "mde_peer_comm_destroy": (_i32, [_vp]),
    "mde_peer_alloc": (_i32, [C.c_size_t, C.POINTER(_vp)]),
    "mde_peer_free": (_i32, [_vp]),
    "mde_peer_export": (_i32, [_vp, C.c_char_p]),
    "mde_peer_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "mde_peer_close": (_i32, [_vp]),
    "mde_metrics_resized": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mde_metrics_finalize_host": (None, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mde_masked_loss": (_i32, [_i32, _vp, _i32, _vp, _vp, _i64, _i64, _i64, C.POINTER(LossParams), _f32,
                               _vp, _vp, _vp, _vp, _vp]),
    "mde_masked_loss_metrics": (_i32, [_i32, _vp, _i32, _vp, _vp, _i64, _i64, _i64, C.POINTER(LossParams), _f32,
                                       _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mde_masked_loss_partials": (_i32, [_i32, _i32, _vp, _i32, _vp, _vp, _i64, C.POINTER(LossParams), _vp, _vp, _vp]),
    "mde_masked_loss_from_totals": (_i32, [_i32, _vp, _i32, _vp, _vp, _i64, C.POINTER(LossParams), _vp, _f32, _vp, _vp, _vp]),
    "mde_scale_inplace": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "mde_ordinal_layer_fwd": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp]),
    "mde_ordinal_layer_bwd": (_i32, [_vp, _i32, _vp, _i64, _i64, _i64, _vp, _vp]),
    "mde_label_to_depth_i64": (_i32, [_vp, _i64, _f32, _f32, _i32, _i32, _vp, _vp]),
    "mde_label_to_depth_f32": (_i32, [_vp, _i64, _f32, _f32, _i32, _i32, _vp, _vp]),
    "mde_depth_to_label": (_i32, [_vp, _i64, _f32, _f32, _i32, _i32, _vp, _vp]),
    "mde_ord_loss": (_i32, [_vp, _vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _vp]),
    "mde_dorn_fused": (_i32, [_vp, _i32, _vp, _i64, _i64, _i64, _f32, _f32, _i32, _f32, _vp, _vp, _vp,
                              _vp, _vp, _vp, _vp]),
    "mde_ordinal_regression_loss": (_i32, [_vp, _vp, _i64, _i64, _i64, _f32, _f32, _i32, _f32, _vp, _vp,
                                           _vp, _vp]),
    "mde_vnl_scratch_bytes": (C.c_size_t, [_i64, _i64, _i64, _i64]),
    "mde_vnl_loss": (_i32, [_vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _f32, _f32, _i32, _f32, _vp, _vp,
                            _vp, _vp, _vp, _vp]),
    "mde_wcel_loss": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _vp, _vp, _vp, _vp]),
    "mde_depth_to_bins": (_i32, [_vp, _i64, _f32, _f32, _f32, _f32, _i64, _vp, _vp]),
    "mde_bins_to_depth": (_i32, [_vp, _i32, _vp, _i64, _i64, _i64, _vp, _vp]),
    "mde_bins_to_depth_bwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "mde_scale_and_shift": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mde_apply_scale_shift": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp]),
    "mde_midas_loss": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "mde_midas_ssi_backward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mde_colored_depthmap": (_i32, [_vp, _i64, _f32, _f32, _i32, _i32, _vp, _vp, _vp]),
    "mde_stdepth_loss": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "mde_robust_scratch_bytes": (C.c_size_t, [_i64]),
    "mde_robust_normalize": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mde_midas_loss_masked": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "mde_robust_backward": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mde_point_cloud": (_i32, [_vp, _i64, _i64, _i64, _f32, _f32, _f32, C.POINTER(C.c_float), _i32, _vp, _vp]),
    "mde_write_ply": (_i64, [C.c_char_p, _vp, _vp, _vp, _i64]),
    "mde_workspace_bytes": (C.c_size_t, [_i64]),
    "mde_workspace_init": (_i32, [_vp, _i64, _vp]),
    "mde_last_error": (C.c_char_p, []),
    "mde_version": (C.c_char_p, []),
    "mde_launch_count": (C.c_uint64, []),
    "mde_device_info": (_i32, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mde_debug_set_trace": (_i32, [_vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libmde_b200.so (once). Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmde_b200.so not found at %s. Build it with `python -m mono_depth_estimation_b200.build` "
                "(nvcc, sm_100a). This package has no CPU / PyTorch fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class MdeError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().mde_last_error()
        raise MdeError("libmde_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s (fp32 / fp16 / bf16 only)" % t.dtype)


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mono_depth_estimation_b200 runs on CUDA tensors only (got a %s tensor); "
                               "there is no CPU fallback" % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("tensors are on different devices: %s vs %s" % (dev, t.device))
    return dev


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---- workspaces: one per (device, stream), grown on demand ------------------------------------------
_workspaces = {}


def workspace(device, n_img: int = 1) -> torch.Tensor:
    """Device scratch for the current stream of `device`, able to hold `n_img` per-image rows."""
    lib = load()
    stream = torch.cuda.current_stream(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream.cuda_stream)
    ent = _workspaces.get(key)
    if ent is not None and ent[1] >= n_img:
        return ent[0]
    cap = max(64, int(n_img))
    if ent is not None:
        cap = max(cap, 2 * ent[1])
    nbytes = int(lib.mde_workspace_bytes(cap))
    with torch.cuda.device(device):
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        check(lib.mde_workspace_init(ptr(buf), cap, stream_ptr(device)))
    _workspaces[key] = (buf, cap)
    return buf


def launch_count() -> int:
    return int(load().mde_launch_count())
