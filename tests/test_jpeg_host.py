"""The frame sink's own JPEG encoder (blackhole_8_b200/csrc/bh8_jpeg.cuh), run on the CPU by a TEST-ONLY
harness that calls the very functions the kernels call.  What a reader of the reference's video.avi relies on
 is checked with OpenCV as the judge: the stream decodes, has the frame's size, and is as close
to the original as OpenCV's own encoder gets at the same quality (cv::VideoWriter's MJPG path is libjpeg with
the same IJG tables).  The GPU tests (tests/test_video_sink.py) then require the device's bytes to be equal
to this harness' bytes."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def jpeg_harness():
    global _LIB
    if _LIB is None:
        out = os.path.join(HERE, "host_harness", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libjpeg_harness.so")
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                        "-I" + os.path.join(O.ROOT, "include"), "-I" + os.path.join(O.ROOT, "blackhole_8_b200", "csrc"),
                        os.path.join(HERE, "host_harness", "jpeg_harness.cc"), "-o", so], check=True)
        _LIB = C.CDLL(so)
        _LIB.bh8_jpeg_host_encode.restype = C.c_long
        _LIB.bh8_jpeg_host_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
    return _LIB


def host_encode(bgr, quality=95, ri=2):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    cap = bgr.size * 2 + 4096
    out = np.empty(cap, np.uint8)
    n = jpeg_harness().bh8_jpeg_host_encode(bgr.ctypes.data, bgr.shape[1], bgr.shape[0], quality, ri, out.ctypes.data, cap)
    assert n > 0, n
    return out[:n].tobytes()


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def frames():
    yield "cfg1_640x360", O.load_golden("cfg1_640x360")["bgr"]
    yield "cfg1_odd_333x187", O.load_golden("cfg1_odd_333x187")["bgr"]  # neither size a multiple of 16
    yield "cfg2_640x360", O.load_golden("cfg2_640x360")["bgr"]
    rng = np.random.default_rng(7)
    yield "noise_96x80", rng.integers(0, 256, (80, 96, 3), dtype=np.uint8)  # the entropy coder's worst customer
    yield "flat_48x32", np.full((32, 48, 3), 255, np.uint8)               # runs of zeros, 0xFF bytes in the stream
    yield "tiny_5x3", O.load_golden("cfg1_tiny_5x3")["bgr"]


@pytest.mark.parametrize("quality", [50, 95, 100])
def test_opencv_decodes_it_as_well_as_its_own(quality):
    for name, img in frames():
        jpg = host_encode(img, quality, 2)
        assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9"
        dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
        assert dec is not None and dec.shape == img.shape, name
        ok, ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
        dec_ref = cv2.imdecode(ref, cv2.IMREAD_COLOR)
        p, p_ref = psnr(dec, img), psnr(dec_ref, img)
        print(name, quality, "own %.2f dB (%d B)  OpenCV %.2f dB (%d B)" % (p, len(jpg), p_ref, len(ref)))
        assert p >= p_ref - 1.5, (name, quality, p, p_ref)
        # the restart markers and absolute DC values at interval starts cost bytes, within reason
        assert len(jpg) <= 1.25 * len(ref) + 1200, (name, len(jpg), len(ref))


@pytest.mark.parametrize("ri", [1, 2, 7, 4096])
def test_every_restart_interval_decodes_to_the_same_picture(ri):
    img = O.load_golden("cfg1_odd_333x187")["bgr"]
    base = cv2.imdecode(np.frombuffer(host_encode(img, 90, 1), np.uint8), cv2.IMREAD_COLOR)
    dec = cv2.imdecode(np.frombuffer(host_encode(img, 90, ri), np.uint8), cv2.IMREAD_COLOR)
    assert dec is not None and np.array_equal(dec, base)  # same coefficients, only the framing differs


def test_stuffing_and_markers():
    jpg = host_encode(np.full((32, 48, 3), 255, np.uint8), 100, 1)
    sos = jpg.index(b"\xff\xda")
    scan = jpg[sos + 14:-2]
    k = 0
    seen_rst = 0
    while k < len(scan) - 1:
        if scan[k] == 0xFF:
            nxt = scan[k + 1]
            assert nxt == 0x00 or 0xD0 <= nxt <= 0xD7, hex(nxt)  # only stuffed 0xFF or RSTm inside the scan
            if nxt != 0:
                assert nxt == 0xD0 + (seen_rst & 7)
                seen_rst += 1
            k += 2
        else:
            k += 1
    assert seen_rst == 3 * 2 - 1  # 6 MCUs, restart interval 1
