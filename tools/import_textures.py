#!/usr/bin/env python
"""Decode the reference's resource images once and store the decoded pixels losslessly.

Runs only where /root/reference is mounted.  The reference loads textures with
cv::imread(resource_dir()/name, IMREAD_COLOR) (include/blackhole/config.h:35-37): 8-bit BGR,
alpha dropped.  The decoded BGR pixels are what the hot path reads, so those are
what the fixtures pin: each image is decoded with cv2.imread(IMREAD_COLOR) (same
libpng / libjpeg-turbo family) and re-encoded as a 3-channel PNG (lossless) under
resource/, with the md5 of the raw BGR bytes recorded in resource/MANIFEST.json.
"""
import hashlib
import json
import os
import sys

import cv2

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/resource"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "resource")

manifest = {}
for name in sorted(os.listdir(SRC)):
    img = cv2.imread(os.path.join(SRC, name), cv2.IMREAD_COLOR)
    if img is None:
        continue
    stem = os.path.splitext(name)[0]
    out = stem + ".bgr.png"
    cv2.imwrite(os.path.join(DST, out), img, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    back = cv2.imread(os.path.join(DST, out), cv2.IMREAD_COLOR)
    assert back.tobytes() == img.tobytes(), name
    manifest[name] = {
        "file": out,
        "rows": int(img.shape[0]),
        "cols": int(img.shape[1]),
        "md5_bgr": hashlib.md5(img.tobytes()).hexdigest(),
    }
    print(name, img.shape, manifest[name]["md5_bgr"][:12])
with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
    json.dump(manifest, f, indent=1, sort_keys=True)
