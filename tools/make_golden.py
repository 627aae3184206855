#!/usr/bin/env python
"""Generate tests/golden/* from the REFERENCE ITSELF (oracle/_ref/ref_render).

Runs only in the build container, where `make -C oracle ref` has compiled the reference's own
classes from /root/reference.  For every case below it stores
  tests/golden/<name>.json       scene + camera snapshot (bh8 POD, %.17g doubles) and run counters
  tests/golden/<name>.png        the BGR frame (lossless PNG)
  tests/golden/<name>.npz        per-pixel hit key, class and step count (compressed)
and for the full-size BASELINE configurations only digests (tests/golden/digests.json), so the
C port (oracle/bh8_oracle.c) is pinned at full size without committing large frames.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
TEX = os.path.join(ROOT, "build", "textures")
OUT = os.path.join(ROOT, "tests", "golden")

# name: (cfg, width, height, frame, nstep or None)
SMALL = {
    "cfg0_960x540": (0, 960, 540, 0, None),     # == frame of the UNCHANGED blackhole_solution_test
    "cfg0_frame7_320x180": (0, 320, 180, 7, None),  # disc spun 7x RotateZ(pi/180)
    "cfg5_480x270": (5, 480, 270, 0, None),     # "0b": background rectangle only
    "cfg1_640x360": (1, 640, 360, 0, None),
    "cfg2_640x360": (2, 640, 360, 0, None),
    "cfg3_frame60_480x270": (3, 480, 270, 60, None),
    "cfg3_frame180_480x270": (3, 480, 270, 180, None),
    "cfg4_nstep200_320x180": (4, 320, 180, 0, None),
    "cfg1_odd_333x187": (1, 333, 187, 0, None),  # ragged size: not a multiple of any tile
    # stress scenes (apps/scenes.h cfg 6-9): the GPU path's rarely taken branches
    "cfg6_offplane_400x225": (6, 400, 225, 0, None),
    "cfg7_lookaway_400x225": (7, 400, 225, 0, None),
    "cfg8_manyplanes_400x225": (8, 400, 225, 0, None),
    "cfg9_inside_320x180": (9, 320, 180, 0, None),
    "cfg11_fourplanes_400x225": (11, 400, 225, 0, None),
    "cfg1_nstep2_320x180": (1, 320, 180, 0, 2),
    "cfg1_tiny_5x3": (1, 5, 3, 0, None),
    # flat space: the ray_tracer_test.cc scene; frame 0 == the UNCHANGED ray_tracer_test's frame
    "cfg10_flat_800x450": (10, 800, 450, 0, None),
    "cfg10_flat_frame25_640x360": (10, 640, 360, 25, None),
}
DIGEST_ONLY = {
    "cfg1_1920x1080": (1, 1920, 1080, 0, None),
    "cfg2_1920x1080": (2, 1920, 1080, 0, None),
    "cfg0_1920x1080": (0, 1920, 1080, 0, None),
    "cfg3_frame239_1920x1080": (3, 1920, 1080, 239, None),
}


def run(name, case, keep):
    cfg, w, h, frame, nstep = case
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "o")
        cmd = [REF, "--cfg", str(cfg), "--width", str(w), "--height", str(h), "--frame", str(frame),
               "--threads", str(os.cpu_count()), "--texdir", TEX, "--out", prefix]
        if nstep:
            cmd += ["--nstep", str(nstep)]
        subprocess.run(cmd, check=True)
        snap = json.load(open(prefix + ".json"))
        raw = open(prefix + ".bgr", "rb").read()
        bgr = np.frombuffer(raw[8:], dtype=np.uint8).reshape(h, w, 3)
        cls = np.fromfile(prefix + ".cls", dtype=np.uint8).reshape(h, w)
        key = np.fromfile(prefix + ".key", dtype=np.int8).reshape(h, w)
        steps = np.fromfile(prefix + ".steps", dtype=np.uint16).reshape(h, w)
    snap["run"].pop("best_ms", None)
    snap["run"].pop("mean_ms", None)
    snap["run"].pop("threads", None)
    digest = {k: hashlib.md5(a.tobytes()).hexdigest() for k, a in
              (("bgr", bgr), ("cls", cls), ("key", key), ("steps", steps))}
    snap["digest"] = digest
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(snap, f, indent=1)
    if keep:
        cv2.imwrite(os.path.join(OUT, name + ".png"), bgr, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), cls=cls, key=key, steps=steps)
    print(name, snap["run"], digest["bgr"][:12])


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/ref_render missing: run `make -C oracle ref` (needs /root/reference)")
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, case in SMALL.items():
        if not only or name in only:
            run(name, case, True)
    for name, case in DIGEST_ONLY.items():
        if not only or name in only:
            run(name, case, False)


if __name__ == "__main__":
    main()
