#!/usr/bin/env python
"""Screen a change of the kernel's warp schedule on the CPU before it costs GPU time.

usage: python tools/schedule_model.py [golden-name] [--width W --height H] [--nstep N] [--define D ...]
Runs the test-only host build of the kernel's per-ray code through the emulated warp schedule
(tests/host_harness: 32 lanes in lockstep, votes, batched exact tests) for a grid of
(updates per vote, batching window) and prints, per setting: update slots per warp, resolve passes per warp
and a crude instruction estimate  slots * 20 + votes * 11 + passes * 550  (the constants come from
profiles/r01g_ncu_function_table.txt; the estimate ranks schedules, it does not predict milliseconds).
--define builds the harness with extra -D flags (an experimental code path behind a macro); the frames of
every run are compared with the default build's and a difference is reported."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", nargs="?", default="cfg1_640x360")
    ap.add_argument("--width", type=int)
    ap.add_argument("--height", type=int)
    ap.add_argument("--nstep", type=int)
    ap.add_argument("--define", action="append", default=[])
    args = ap.parse_args()
    import oracle_lib as O
    from test_ray_math_host import harness_render, harness_render_warps
    snap = O.load_golden(args.name)["snap"]
    if args.width and args.height:
        snap = snap.with_resolution(args.width, args.height)
    base = harness_render(snap, nstep=args.nstep)
    n = base["cls"].size
    print("%s %dx%d nstep %s: %.2f updates per ray, %.3f exact tests per ray (one lane at a time)" %
          (args.name, snap.width, snap.height, args.nstep or snap.nstep, base["updates"] / n, base["exact_tests"] / n))
    print("updates/vote  window   slots/warp  passes/warp  estimate   frames")
    for upv in (1, 2, 3, 4):
        for wait in (0, 1, 2, 4):
            r = harness_render_warps(snap, nstep=args.nstep, updates_per_vote=upv, resolve_wait=wait,
                                     defines=tuple(args.define))
            same = all(np.array_equal(r[k], base[k]) for k in ("bgr", "cls", "key", "steps"))
            s, p = r["update_slots_per_warp"], r["resolve_passes_per_warp"]
            print("%8d %9d %12.1f %12.2f %9.0f   %s" % (upv, wait, s, p, s * 20 + s / upv * 11 + p * 550,
                                                         "identical" if same else "DIFFERENT"))


if __name__ == "__main__":
    main()
