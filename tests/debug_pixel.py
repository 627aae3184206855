#!/usr/bin/env python3
"""Debugging aid: trace ONE pixel of a random scene through the host-compiled kernel code, printing the lane after
every update.  usage: python tests/debug_pixel.py <seed> <width> <height> <x> <y> [filter_slots]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import oracle_lib as O
    from blackhole_8_b200 import abi
    from random_scenes import random_snapshot
    from test_ray_math_host import HarnessTexture, harness
    seed, w, h, x, y = (int(v) for v in sys.argv[1:6])
    slots = int(sys.argv[6]) if len(sys.argv) > 6 else 4
    snap = random_snapshot(seed, w, h)
    L = harness(())
    texs = [O.load_texture(n) for n in snap.textures]
    tarr = (HarnessTexture * max(1, len(texs)))()
    for i, t in enumerate(texs):
        tarr[i] = HarnessTexture(t.ctypes.data, t.shape[0], t.shape[1])
    prm = snap.params(abi.PIXEL_BGR8, 0, None)
    err = C.create_string_buffer(256)
    rc = L.bh8_harness_trace_pixel(C.byref(snap.scene), C.byref(snap.camera), C.byref(prm), tarr, len(texs), slots, x, y, err)
    assert rc == 0, err.value
    ref = O.render(snap)
    print("oracle: cls", ref["cls"][y, x], "key", ref["key"][y, x], "steps", ref["steps"][y, x])


if __name__ == "__main__":
    main()
