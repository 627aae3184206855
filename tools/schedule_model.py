#!/usr/bin/env python
"""Screen a change of the kernel's warp schedule on the CPU before it costs GPU time.

usage: python tools/schedule_model.py [golden-name] [--width W --height H] [--nstep N] [--define D ...]
Runs the test-only host build of the kernel's per-ray code through the emulated warp schedule
(tests/host_harness: 32 lanes in lockstep, votes, batched exact tests) for a grid of
(updates per vote, batching window) and prints, per setting: update slots per warp, resolve passes per warp,
trips through the rare path per warp (update slots in which ANY lane entered lane_update_rare -- what the warp
pays for) and a crude instruction estimate  slots * 19 + votes * 6 + passes * 490 + trips * 88  (the constants
come from profiles/r02zz_instr_model.json; the estimate ranks schedules, it does not predict milliseconds).
For the kernel's own setting the trips are broken down by what they did (BH8_TRACE in bh8_ray.cuh).
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
    import ctypes as C
    from test_ray_math_host import harness, harness_render, harness_render_warps
    reasons = ("any", "trigger fired", "lease limit", "filter (2) run", "previous side evaluated", "lease granted",
               "lease ran out", "leg change", "parked by filter (2)", "frozen by trigger", "t >= 1", "phi trigger",
               "slow-always ray", "inbound filter")

    def take_trace(lib):
        a, b = (C.c_ulonglong * 16)(), (C.c_ulonglong * 16)()
        lib.bh8_harness_take_trace(a, b)
        return list(a), list(b)
    snap = O.load_golden(args.name)["snap"]
    if args.width and args.height:
        snap = snap.with_resolution(args.width, args.height)
    base = harness_render(snap, nstep=args.nstep)
    n = base["cls"].size
    print("%s %dx%d nstep %s: %.2f updates per ray, %.3f exact tests per ray (one lane at a time)" %
          (args.name, snap.width, snap.height, args.nstep or snap.nstep, base["updates"] / n, base["exact_tests"] / n))
    lib = harness(tuple(args.define))
    n_warps = ((snap.width + 7) // 8) * ((snap.height + 3) // 4)
    take_trace(lib)
    print("updates/vote  window   slots/warp  passes/warp  rare trips/warp  estimate   frames")
    kernel_trace = None
    for upv in (1, 2, 3, 4):
        for wait in (0, 1, 2, 4):
            r = harness_render_warps(snap, nstep=args.nstep, updates_per_vote=upv, resolve_wait=wait,
                                     defines=tuple(args.define))
            same = all(np.array_equal(r[k], base[k]) for k in ("bgr", "cls", "key", "steps"))
            s, p = r["update_slots_per_warp"], r["resolve_passes_per_warp"]
            per_lane, per_slot = take_trace(lib)
            trips = per_slot[0] / n_warps
            if (upv, wait) == (2, 2):
                kernel_trace = (per_lane, per_slot)
            print("%8d %9d %12.1f %12.2f %12.2f %13.0f   %s" %
                  (upv, wait, s, p, trips, s * 19 + s / upv * 6 + p * 490 + trips * 88, "identical" if same else "DIFFERENT"))
    if kernel_trace:
        print("the kernel's setting (2 updates per vote, window 2), rare path by reason: per ray / per warp (slots)")
        for k, name in enumerate(reasons):
            print("  %-26s %7.3f %7.3f" % (name, kernel_trace[0][k] / n, kernel_trace[1][k] / n_warps))


if __name__ == "__main__":
    main()
