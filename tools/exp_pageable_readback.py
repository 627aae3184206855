#!/usr/bin/env python3
"""Synchronous bh8_render of one 1080p BGR8 frame into pinned, pageable and registered (bh8_host_register) host
memory.  usage (GPU box): python tools/exp_pageable_readback.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from blackhole_8_b200 import abi
from blackhole_8_b200.renderer import Renderer
import ctypes as C
seq = bench.frame_sequence("cfg1_spin", 64)
H, W = seq[0].height, seq[0].width
r = Renderer((0,))
r.set_textures(seq[0], bench.load_texture)
pin = r.pinned((H, W, 3))
page = np.empty((H, W, 3), np.uint8); page[:] = 0
cudart = None
for name, arr in (("pinned", pin.array), ("pageable", page), ("registered", page)):
    if name == "registered":
        t0 = time.perf_counter()
        rc = r.host_register(arr)
        print("cudaHostRegister rc", rc, "%.3f ms" % ((time.perf_counter() - t0) * 1e3))
    out = {"pixels": arr.reshape(1, H, W, 3)}
    for i in range(5):
        r.render(seq[i], pixel_format=abi.PIXEL_BGR8, out=out)
    best = 0
    for rep in range(3):
        t0 = time.perf_counter(); n = 200
        for i in range(n):
            r.render(seq[i % 64], pixel_format=abi.PIXEL_BGR8, out=out)
        best = max(best, n / (time.perf_counter() - t0))
    print("%-10s %7.1f frames/s  %.4f ms/frame" % (name, best, 1e3 / best), flush=True)
t0 = time.perf_counter(); rc = r.host_unregister(page); print("unregister rc", rc, "%.3f ms" % ((time.perf_counter() - t0) * 1e3))
