"""Helpers for the -m gpu tests: render a SceneSnapshot through the C ABI (libbh8.so)."""
import numpy as np

import oracle_lib as O
from blackhole_8_b200 import abi

_RENDERER = None


def renderer():
    global _RENDERER
    if _RENDERER is None:
        from blackhole_8_b200.renderer import Renderer
        _RENDERER = Renderer((0,))
    return _RENDERER


def gpu_render(snap, nstep=None, pixel_format=abi.PIXEL_BGR8, flags=0, stats=True, r=None):
    r = r or renderer()
    r.set_textures(snap, O.load_texture)
    res = r.render(snap, nstep=nstep, pixel_format=pixel_format, want_maps=True, flags=flags, stats=stats)
    px = res["pixels"][0]
    if pixel_format == abi.PIXEL_BGR8:
        bgr = px
    elif pixel_format == abi.PIXEL_BGRA8:
        bgr = px[..., :3]
    else:
        bgr = px[..., 2::-1]
    out = {"bgr": np.ascontiguousarray(bgr), "cls": res["cls"][0], "key": res["key"][0],
           "steps": res["steps"][0], "pixels": px}
    if stats:
        out["stats"] = res["stats"]
    return out
