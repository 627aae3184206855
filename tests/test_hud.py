"""HUD text (bh8_draw_text): the reference draws camera position, basis vectors and field of view into the
frame with cv::putText(FONT_HERSHEY_PLAIN, scale 1, green, thickness 1) after rendering
(blackhole_solution_test.cc:313-325).  The library's blit over a glyph table rasterised from OpenCV must
equal cv2.putText bit for bit for text inside the image."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from blackhole_8_b200.renderer import draw_text  # noqa: E402

import oracle_lib as O  # noqa: E402


def put(img, text, x, y, color=(0, 255, 0)):
    cv2.putText(img, text, (x, y), cv2.FONT_HERSHEY_PLAIN, 1, color, 1)
    return img


def test_the_reference_hud_lines_equal_opencv():
    frame = np.ascontiguousarray(O.load_golden("cfg1_640x360")["bgr"])
    lines = ["Position: [-2000, 0, 400]", "VectorX: [1, 0, 0]", "VectorY: [0, -1, 0]", "VectorZ: [0, 0, -1]",
             "FoV: 90 deg", "Position: [-800.175, -9.99848, 400]", "VectorX: [0.999848, -0.0174524, 0]"]
    want, got = frame.copy(), frame.copy()
    for k, text in enumerate(lines):  # :313-325: {0, 10}, {0, 25}, {0, 40}, ...
        put(want, text, 0, 10 + 15 * k)
        draw_text(got, text, 0, 10 + 15 * k)
    assert (want != frame).any() and np.array_equal(want, got)


def test_random_text_inside_the_image_equals_opencv():
    rng = np.random.default_rng(5)
    for _ in range(300):
        text = "".join(chr(int(rng.integers(32, 127))) for _ in range(int(rng.integers(1, 45))))
        x, y = int(rng.integers(0, 40)), int(rng.integers(13, 100))
        color = tuple(int(v) for v in rng.integers(0, 256, 3))
        base = rng.integers(0, 256, (110, 720, 3)).astype(np.uint8)
        want, got = put(base.copy(), text, x, y, color), draw_text(base.copy(), text, x, y, color)
        assert np.array_equal(want, got), (text, x, y)


def test_row_stride_clipping_and_unprintable_characters():
    wide = np.zeros((40, 200, 3), np.uint8)
    view = wide[:, 20:120]                     # a view: rows 300 bytes wide, stride 600
    draw_text(view, "stride", 2, 20)
    want = put(np.zeros((40, 100, 3), np.uint8), "stride", 2, 20)
    assert np.array_equal(view, want) and not wide[:, :20].any() and not wide[:, 120:].any()
    small = np.zeros((8, 30, 3), np.uint8)     # text larger than the image: nothing outside is touched
    draw_text(small, "WWWWWWWW", -3, 5)
    assert small.any()
    a, b = np.zeros((30, 60, 3), np.uint8), np.zeros((30, 60, 3), np.uint8)
    draw_text(a, "a\x01b\x7f", 1, 20)
    draw_text(b, "a?b?", 1, 20)
    assert np.array_equal(a, b)


def test_cpp_draw_hud_equals_the_reference_block(tmp_path):
    """blackhole::gpu::DrawHud (include/blackhole/gpu/renderer.h) formats and places the five lines as
    blackhole_solution_test.cc:309-326 does; the pixels must be what cv2.putText draws for those strings."""
    import os
    import subprocess
    root = O.ROOT
    exe, raw = str(tmp_path / "hud_check"), str(tmp_path / "hud.bgr")
    subprocess.run(["g++", "-std=gnu++17", "-O2", "-DNDEBUG", "-I" + os.path.join(root, "third_party", "cvshim"),
                    "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "host_harness", "hud_check.cc"),
                    "-o", exe, "-L" + os.path.join(root, "blackhole_8_b200"), "-lbh8",
                    "-Wl,-rpath," + os.path.join(root, "blackhole_8_b200")], check=True)
    lines = subprocess.run([exe, raw], check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(lines) == 5 and lines[0].startswith("Position: [-") and lines[4] == "FoV: 90 deg"
    got = np.fromfile(raw, np.uint8).reshape(360, 640, 3)
    want = np.zeros((360, 640, 3), np.uint8)
    for k, text in enumerate(lines):
        put(want, text, 0, 10 + 15 * k)
    assert want.any() and np.array_equal(want, got)
