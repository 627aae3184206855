#!/usr/bin/env python3
"""One-off parity campaign beyond the test suite's 48 GPU seeds: seeded random scenes (tests/random_scenes.py)
rendered by the CUDA path through the C ABI and by the C oracle, differences counted per pixel.
usage (GPU box): python tools/gpu_random_campaign.py [first_seed] [n_scenes] [nstep list, e.g. 64,100,200 or -] [sizes, e.g. 640x360,1920x1080] > profiles/rNN_random_campaign.txt
With an nstep list the scenes' own step counts (7 / 20 / 50) are replaced by its entries in turn: from 64 on the kernel
takes its fine-step instantiation (three updates per round of votes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_lib as O
    import parity
    from gpu_util import gpu_render
    from random_scenes import random_snapshot
    # $BH8_CAMPAIGN_DEVICES=0,0,0: through a multi-device context (an ordinal listed twice is a logical device), i.e.
    # every frame in interleaved 16-row stripes gathered in device 0's buffer
    r = None
    if os.environ.get("BH8_CAMPAIGN_DEVICES"):
        from blackhole_8_b200.renderer import Renderer
        r = Renderer(tuple(int(v) for v in os.environ["BH8_CAMPAIGN_DEVICES"].split(",")))
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    nsteps = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    sizes = [tuple(int(v) for v in t.split("x")) for t in sys.argv[4].split(",")] if len(sys.argv) > 4 else \
        [(192, 108), (256, 144), (333, 187)]
    tot = dict(pixels=0, cls=0, rgb=0, steps=0, scenes_with_any=0)
    classes = np.zeros(4, np.int64)
    worst = []
    t0 = time.time()
    for seed in range(first, first + n):
        W, H = sizes[seed % len(sizes)]
        snap = random_snapshot(seed, W, H)
        ns = nsteps[seed % len(nsteps)] if nsteps else None
        ref = O.render(snap, nstep=ns)
        got = gpu_render(snap, nstep=ns, stats=bool(seed & 1), r=r)  # both kernel instantiations
        cls_bad = int((got["cls"] != ref["cls"]).sum())
        diff = np.abs(got["bgr"].astype(int) - ref["bgr"].astype(int)).max(axis=2)
        rgb_bad = int(((got["cls"] == ref["cls"]) & (diff > parity.RGB_TOL)).sum())
        steps_bad = int((got["steps"] != ref["steps"]).sum())
        tot["pixels"] += ref["cls"].size
        tot["cls"] += cls_bad
        tot["rgb"] += rgb_bad
        tot["steps"] += steps_bad
        classes += np.bincount(ref["cls"].ravel(), minlength=4)[:4]
        if cls_bad or rgb_bad or steps_bad:
            tot["scenes_with_any"] += 1
            worst.append((cls_bad + rgb_bad + steps_bad, seed, cls_bad, rgb_bad, steps_bad, W, H))
    print("random scenes %d..%d (%d scenes, %d pixels, frame sizes %s, nstep %s), "
          "CUDA path%s vs C oracle, %.0f s" % (first, first + n - 1, n, tot["pixels"], " / ".join("%dx%d" % t for t in sizes),
                                            " / ".join(str(v) for v in nsteps) if nsteps else "7 / 20 / 50",
                                            " (devices %s, striped)" % os.environ["BH8_CAMPAIGN_DEVICES"] if r else "",
                                            time.time() - t0))
    print("pixels by class (background, horizon, disc, object):", classes.tolist())
    print("hit class differs: %d pixels; colour differs by more than %d/255 (same class): %d pixels; "
          "step count differs: %d pixels; scenes with any difference: %d"
          % (tot["cls"], parity.RGB_TOL, tot["rgb"], tot["steps"], tot["scenes_with_any"]))
    for w in sorted(worst, reverse=True)[:10]:
        print("  seed %d (%dx%d): class %d, colour %d, steps %d" % (w[1], w[5], w[6], w[2], w[3], w[4]))


if __name__ == "__main__":
    main()
