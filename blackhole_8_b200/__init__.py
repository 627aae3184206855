"""B200-native renderer for blackhole_8's per-pixel Schwarzschild null-geodesic hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/bh8.h), abi.py (ctypes mirror of the PODs),
renderer.py (host-side render call), build.py (nvcc, sm_100a, in-tree libbh8.so).
"""
from . import abi  # noqa: F401
from .abi import SceneSnapshot  # noqa: F401

__all__ = ["abi", "SceneSnapshot", "Renderer", "load_library"]


def __getattr__(name):
    if name in ("Renderer", "load_library", "Bh8Error"):
        from . import renderer
        return getattr(renderer, name)
    raise AttributeError(name)
