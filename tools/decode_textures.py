#!/usr/bin/env python
"""Write resource/*.bgr.png as raw '<reference name>.bgr' files (int32 rows, int32 cols, BGR bytes).

That raw layout is what third_party/cvshim's cv::imread and the oracle binaries load
(no image codec in C).  Usage: decode_textures.py [out_dir]   (default build/textures)
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def load_texture(name):
    """Decoded BGR uint8 array (rows, cols, 3) for a reference resource name, md5-checked."""
    with open(os.path.join(ROOT, "resource", "MANIFEST.json")) as f:
        ent = json.load(f)[name]
    img = cv2.imread(os.path.join(ROOT, "resource", ent["file"]), cv2.IMREAD_COLOR)
    if img is None or hashlib.md5(img.tobytes()).hexdigest() != ent["md5_bgr"]:
        raise RuntimeError("texture %s does not match resource/MANIFEST.json" % name)
    return np.ascontiguousarray(img)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "build", "textures")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(ROOT, "resource", "MANIFEST.json")) as f:
        names = sorted(json.load(f))
    for name in names:
        img = load_texture(name)
        with open(os.path.join(out, name + ".bgr"), "wb") as f:
            f.write(np.array(img.shape[:2], dtype=np.int32).tobytes())
            f.write(img.tobytes())
    print("wrote %d textures to %s" % (len(names), out))


if __name__ == "__main__":
    main()
