#!/usr/bin/env python3
"""The random-scene parity campaign without a GPU: the kernel's per-ray code compiled for the host
(tests/host_harness: one lane at a time, and every fourth scene through the emulated warp schedule) against the C
oracle.  usage: python tools/cpu_random_campaign.py [first_seed] [n_scenes] [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_lib as O
    from random_scenes import random_snapshot
    from test_ray_math_host import harness_render, harness_render_warps
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1500.0
    t0, bad, done, px = time.time(), [], 0, 0
    for seed in range(first, first + n):
        w, h = ((96, 54), (128, 72), (160, 90), (192, 108), (64, 36))[seed % 5]
        snap = random_snapshot(seed, w, h)
        ref = O.render(snap)
        got = harness_render(snap)
        d = int((got["cls"] != ref["cls"]).sum()) + int((got["steps"] != ref["steps"]).sum())
        if seed % 4 == 0:
            gw = harness_render_warps(snap)
            d += int((gw["cls"] != ref["cls"]).sum()) + int((gw["steps"] != ref["steps"]).sum())
        done += 1
        px += w * h
        if d:
            bad.append((seed, w, h, d))
            print("DIFF seed %d %dx%d: %d pixels" % (seed, w, h, d), flush=True)
        if time.time() - t0 > budget:
            break
    print("host-compiled kernel code vs C oracle: random scenes %d..%d (%d scenes, %d pixels), scenes with class or "
          "step-count differences: %d %s, %.0f s" % (first, first + done - 1, done, px, len(bad), bad[:10], time.time() - t0))


if __name__ == "__main__":
    main()
