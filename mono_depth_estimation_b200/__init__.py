"""B200-native per-pixel depth supervision and evaluation (drop-in for the hot path of
xeTaiz/mono-depth-estimation: criteria.py losses, metrics.py, DORN ordinal decode, VNL,
depth2pointcloud). The CUDA library is loaded lazily on first use; there is no CPU fallback."""

__version__ = "0.1.0"

from . import synth  # noqa: F401  (pure torch, no CUDA needed)


def __getattr__(name):
    import importlib
    if name in ("criteria", "metrics", "dorn", "pointcloud", "distributed", "wcel", "stdepth", "visualize", "_lib", "build"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
