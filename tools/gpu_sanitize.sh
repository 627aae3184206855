#!/bin/bash
# compute-sanitizer over the render kernel (all instantiations the fixtures reach), the device HUD and the JPEG
# encoder with its worst customer (noise at quality 100).  usage (GPU box): bash tools/gpu_sanitize.sh
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import oracle_lib as O, gpu_util, parity
from blackhole_8_b200 import abi
from blackhole_8_b200.renderer import VideoSink
r = gpu_util.renderer()
for name in ("cfg1_odd_333x187", "cfg0_frame7_320x180", "cfg2_640x360", "cfg8_manyplanes_400x225", "cfg11_fourplanes_400x225",
             "cfg9_inside_320x180", "cfg10_flat_frame25_640x360", "cfg1_tiny_5x3"):
    g = O.load_golden(name)
    for stats in (True, False):
        got = gpu_util.gpu_render(g["snap"], stats=stats)
    rep = parity.compare(got, g)
    print(name, rep["class_agreement"], rep["rgb_outlier_share"])
rng = np.random.default_rng(3)
for (h, w, q) in ((187, 333, 100), (64, 96, 100), (360, 640, 95)):
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    import ctypes as C
    buf = r.frame_alloc(h * w * 3)
    import torch
    t = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf), C.c_void_p(t.data_ptr()), C.c_size_t(h * w * 3), 3)
    sink = VideoSink(r, None, w, h, quality=q)
    sink.hud([("noise %dx%d" % (w, h), 0, 12), ("x" * 200, w - 30, h + 5)])
    sink.write_device(buf)
    jpg = sink.last_jpeg()
    sink.close()
    r.frame_free(buf)
    import cv2
    dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
    print("noise", h, w, q, len(jpg), dec.shape)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool python /tmp/san_case.py 2>&1 | grep -v "^$" | tail -14
done
