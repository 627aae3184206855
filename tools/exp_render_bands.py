#!/usr/bin/env python3
"""Synchronous bh8_render of one configs[1] frame into pinned host memory, frames/s by the number of row bands
($BH8_RENDER_BANDS is read at context creation).  usage (GPU box): python tools/exp_render_bands.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from blackhole_8_b200 import abi
    from blackhole_8_b200.renderer import Renderer
    seq = bench.frame_sequence("cfg1_spin", 64)
    H, W = seq[0].height, seq[0].width
    ref = None
    for fmt, name, bpp in ((abi.PIXEL_BGR8, "BGR8", 3), (abi.PIXEL_RGBA8, "RGBA8", 4)):
        for bands in (1, 2, 3, 4, 5, 6, 8):
            os.environ["BH8_RENDER_BANDS"] = str(bands)
            r = Renderer((0,))
            r.set_textures(seq[0], bench.load_texture)
            pin = r.pinned((H, W, bpp))
            out = {"pixels": pin.array.reshape(1, H, W, bpp)}
            for i in range(5):
                r.render(seq[i], pixel_format=fmt, out=out)
            best = 0.0
            for rep in range(3):
                t0 = time.perf_counter()
                n = 300
                for i in range(n):
                    r.render(seq[i % len(seq)], pixel_format=fmt, out=out)
                best = max(best, n / (time.perf_counter() - t0))
            r.render(seq[7], pixel_format=fmt, out=out)
            frame = pin.array.copy()
            if bands == 1:
                ref = frame
            same = bool(np.array_equal(frame, ref))
            print("%-5s bands %d: %7.1f frames/s  %7.1f Mrays/s  %.4f ms/frame  same_as_one_band %s"
                  % (name, bands, best, best * H * W / 1e6, 1e3 / best, same), flush=True)
            pin.free()
            r.close()


if __name__ == "__main__":
    main()
