#!/usr/bin/env python
"""Snapshot the 240 frames of BASELINE configs[3] (camera fly-through) with the REFERENCE's classes.

oracle/_ref/ref_render --cfg 3 --frame k replays the reference's own Camera::MoveX/RotateZ/MoveY and
Annulus::RotateZ calls (object.h:58-88) k times and dumps the resulting state; this tool stores, per
frame, the camera and the vertices of every object in tests/golden/states/cfg3_flythrough.json (doubles as
%.17g).  Runs only in the build container.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")


def states(cfg, name, n_frames=240, width=1920, height=1080, movers=None):
    frames = []
    base = None
    for k in range(n_frames):
        out = subprocess.run([REF, "--cfg", str(cfg), "--width", str(width), "--height", str(height), "--frame", str(k),
                              "--snapshot-only", "--texdir", os.path.join(ROOT, "build", "textures")],
                             check=True, capture_output=True, text=True).stdout
        d = json.loads(out)
        if base is None:
            base = d
        frames.append({"camera": {k2: d["camera"][k2] for k2 in ("pos", "vx", "vy", "vz")},
                       "v": [o["v"] for o in d["objects"]]})
    base.pop("run", None)
    doc = {"base": base, "frames": frames}
    if movers:  # object indices of the animated rectangles, by texture name
        slot = {n: i for i, n in enumerate(base["textures"])}
        doc["movers"] = [[o["tex_id"] for o in base["objects"]].index(slot[n]) for n in movers]
    with open(os.path.join(ROOT, "tests", "golden", "states", name), "w") as f:
        json.dump(doc, f)
    print("wrote %d frames to %s" % (len(frames), name))


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/ref_render missing: run `make -C oracle ref`")
    states(3, "cfg3_flythrough.json")  # camera script + disc spin
    states(1, "cfg1_spin.json")        # cfg 1 as the reference's frame loop shows it: disc RotateZ(pi/180) per frame
    # the flat-space driver's "Movement test" (ray_tracer_test.cc:237-261): MoveX/Y/Z, RotateX, RotateY
    states(10, "cfg10_movers.json", n_frames=48, width=800, height=450,
           movers=["mooni.jpeg", "karina.jpeg", "winter.jpg"])


if __name__ == "__main__":
    main()
