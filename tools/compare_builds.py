#!/usr/bin/env python3
"""Render the same frames with two builds of libbh8.so and count the pixels that differ.

usage (on a GPU box): python tools/compare_builds.py dump <out.npz> [--lib path]   # one process per build
                      python tools/compare_builds.py diff a.npz b.npz
Frames: configs[1] spin frames 0/10/100 at 1920x1080, the configs[3] fly-through frames 60/180 and
configs[4]'s scene at 1920x1080 with nstep 200."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dump(out, lib):
    if lib:
        os.environ["BH8_LIB_PATH"] = lib
    import bench
    from blackhole_8_b200.renderer import Renderer
    r = Renderer((0,))
    frames = {}
    spin = bench.frame_sequence("cfg1_spin", 240)
    fly = bench.frame_sequence("cfg3_flythrough", 240)
    r.set_textures(spin[0], bench.load_texture)
    for name, snap, nstep in (("spin0", spin[0], None), ("spin10", spin[10], None), ("spin100", spin[100], None),
                              ("fly60", fly[60], None), ("fly180", fly[180], None), ("spin0_nstep200", spin[0], 200)):
        res = r.render(snap, nstep=nstep, want_maps=True)
        frames[name + "_bgr"] = np.asarray(res["pixels"])[0]
        frames[name + "_cls"] = np.asarray(res["cls"])[0]
        frames[name + "_steps"] = np.asarray(res["steps"])[0]
    np.savez(out, **frames)


def diff(a, b):
    A, B = np.load(a), np.load(b)
    for k in A.files:
        x, y = A[k], B[k]
        if k.endswith("_bgr"):
            px = (x != y).reshape(x.shape[0], x.shape[1], -1).any(axis=2)
            big = (np.abs(x.astype(int) - y.astype(int)) > 2).reshape(x.shape[0], x.shape[1], -1).any(axis=2)
            print("%-22s differing pixels %7d of %d (%.5f %%), beyond 2/255: %d" %
                  (k, px.sum(), px.size, 100.0 * px.mean(), big.sum()))
        else:
            print("%-22s differing %7d of %d" % (k, (x != y).sum(), x.size))


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2], sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--lib" else None)
    else:
        diff(sys.argv[2], sys.argv[3])
