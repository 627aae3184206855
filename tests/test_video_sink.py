"""Frame sink: Motion-JPEG AVI writer behind bh8_sink_*, the stand-in for the reference's
cv::VideoWriter(fourcc MJPG) of blackhole_solution_test.cc:71-72,334.

A JPEG stream cannot be byte-identical between encoders, so parity here means what a consumer of the
reference's video.avi relies on: the file is a valid MJPG AVI that OpenCV itself reads back with the
right frame count, size and rate, and every decoded frame matches the rendered frame as closely as
OpenCV's own JPEG encoder at the same quality does (PSNR, stated below).
CPU tests cover the container (fed with JPEGs from cv2.imencode); the `gpu` tests the device encoder (own kernels) and the HUD blit."""
import os
import struct

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from blackhole_8_b200 import abi  # noqa: E402
from blackhole_8_b200.renderer import Bh8Error, VideoSink  # noqa: E402

import oracle_lib as O  # noqa: E402

PSNR_MIN_DB = 34.0          # quality 95, 4:2:0: OpenCV's own encoder + AVI reader give 37-38 dB on these frames
PSNR_BELOW_OPENCV_DB = 1.5  # the sink's encoder may be at most this much worse than cv2.imencode at equal quality


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def read_avi(path):
    cap = cv2.VideoCapture(path)
    assert cap.isOpened(), "OpenCV cannot open the AVI"
    info = {"fps": cap.get(cv2.CAP_PROP_FPS), "w": int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
            "h": int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), "fourcc": int(cap.get(cv2.CAP_PROP_FOURCC))}
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    cap.release()
    return info, frames


def test_container_is_a_valid_mjpg_avi(tmp_path):
    g = O.load_golden("cfg1_640x360")
    frame = np.ascontiguousarray(g["bgr"])
    h, w = frame.shape[:2]
    path = str(tmp_path / "video.avi")
    jpegs = []
    with VideoSink(None, path, w, h, fps=29, quality=95) as sink:
        for k in range(5):
            f = np.roll(frame, 7 * k, axis=1)
            ok, enc = cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 95])
            assert ok
            jpegs.append((f, enc.tobytes()))
            sink.append_jpeg(jpegs[-1][1])
            assert sink.last_jpeg() == jpegs[-1][1]
        st = sink.stats()
        assert st["frames"] == 5 and st["jpeg_bytes"] == sum(len(j) for _, j in jpegs)
        size = sink.close()
    assert size == os.path.getsize(path)
    raw = open(path, "rb").read()
    # RIFF header, stream handler, frame count in avih and strh, index
    assert raw[:4] == b"RIFF" and raw[8:12] == b"AVI " and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    assert raw.count(b"00dc") == 10 and b"idx1" in raw and b"vidsMJPG" in raw
    avih = raw.index(b"avih")
    assert struct.unpack("<I", raw[avih + 8 + 16:avih + 8 + 20])[0] == 5
    info, frames = read_avi(path)
    assert (info["w"], info["h"]) == (w, h) and abs(info["fps"] - 29.0) < 1e-6
    assert info["fourcc"] == cv2.VideoWriter_fourcc(*"MJPG")
    assert len(frames) == 5
    for (src, enc), got in zip(jpegs, frames):
        ref = cv2.imdecode(np.frombuffer(enc, np.uint8), cv2.IMREAD_COLOR)
        assert psnr(got, ref) > 40.0        # same bitstream, two decoders (ffmpeg vs libjpeg-turbo colour conversion)
        assert psnr(got, src) > PSNR_MIN_DB


def test_sink_rejects_bad_input(tmp_path):
    with pytest.raises(Bh8Error):
        VideoSink(None, str(tmp_path / "no_such_dir" / "v.avi"), 64, 64)
    with pytest.raises(Bh8Error):
        VideoSink(None, None, 0, 64)
    with pytest.raises(Bh8Error):
        VideoSink(None, None, 64, 64, quality=0)
    sink = VideoSink(None, None, 64, 64)
    with pytest.raises(Bh8Error):
        sink.append_jpeg(b"not a jpeg at all")
    with pytest.raises(Bh8Error):
        sink.write_device(0x1000)  # a host-only sink has no encoder
    sink.close()


@pytest.mark.gpu
def test_gpu_encoded_frames_match_the_rendered_frames(tmp_path):
    from gpu_util import renderer
    g = O.load_golden("cfg1_640x360")
    snap = g["snap"]
    r = renderer()
    r.set_textures(snap, O.load_texture)
    h, w = snap.height, snap.width
    path = str(tmp_path / "video.avi")
    rendered = []
    with VideoSink(r, path, w, h, fps=29, quality=95) as sink:
        for k in range(6):
            d = snap.to_dict()
            d["camera"]["pos"] = [d["camera"]["pos"][0] + 40.0 * k] + list(d["camera"]["pos"][1:])
            s = abi.SceneSnapshot.from_dict(d)
            sink.render(s)
            jpg = sink.last_jpeg()
            assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9"
            frame = r.render(s, pixel_format=abi.PIXEL_BGR8)["pixels"][0]
            rendered.append((frame, jpg))
        st = sink.stats()
        assert st["frames"] == 6 and st["encode_ms"] > 0
        sink.close()
    info, frames = read_avi(path)
    assert (info["w"], info["h"]) == (w, h) and len(frames) == 6
    for (frame, jpg), got in zip(rendered, frames):
        dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
        ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, 95])
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        p_gpu, p_cv = psnr(dec, frame), psnr(ref, frame)
        print("PSNR sink %.2f dB, OpenCV %.2f dB, bytes %d vs %d" % (p_gpu, p_cv, len(jpg), len(enc)))
        assert p_gpu > PSNR_MIN_DB and p_gpu > p_cv - PSNR_BELOW_OPENCV_DB
        assert psnr(got, dec) > 40.0
        assert 0.8 < len(jpg) / len(enc) < 1.25
        # the kernels run the code the CPU harness runs (tests/test_jpeg_host.py): same bytes
        from test_jpeg_host import host_encode
        assert jpg == host_encode(frame, 95, 2)


@pytest.mark.gpu
def test_hud_is_drawn_on_the_device_before_the_frame_is_encoded(tmp_path):
    """blackhole_solution_test.cc:309-334: five cv::putText lines go into the frame, THEN out_capture.write().
    The device blit must equal cv2.putText bit for bit on the raw frame, and the sink must encode the frame
    with the text in it."""
    from gpu_util import renderer
    from test_jpeg_host import host_encode
    g = O.load_golden("cfg1_640x360")
    snap = g["snap"]
    r = renderer()
    r.set_textures(snap, O.load_texture)
    h, w = snap.height, snap.width
    lines = abi.reference_hud(snap.camera) + [("~!@ {|} 0123456789 WMwm_", 7, 300)]
    frame = r.render(snap, pixel_format=abi.PIXEL_BGR8)["pixels"][0]
    want = frame.copy()
    for text, x, y in lines:
        cv2.putText(want, text, (x, y), cv2.FONT_HERSHEY_PLAIN, 1, (0, 255, 0), 1)
    assert (want != frame).any()
    buf = r.frame_alloc(h * w * 3)
    r.render_device(snap, buf, pixel_format=abi.PIXEL_BGR8)
    r.hud_draw_device(buf, w, h, lines)
    got = np.empty((h, w, 3), np.uint8)
    r.memcpy_d2h(got, buf)
    r.frame_free(buf)
    assert np.array_equal(got, want)
    with VideoSink(r, None, w, h) as sink:
        sink.hud(lines)
        sink.render(snap)
        with_hud = sink.last_jpeg()
        sink.submit(snap)       # the pipelined path draws it too
        sink.flush()
        assert sink.last_jpeg() == with_hud
        sink.hud([])
        sink.render(snap)
        without = sink.last_jpeg()
    assert with_hud == host_encode(want, 95, 2) and without == host_encode(frame, 95, 2)


@pytest.mark.gpu
def test_write_device_takes_a_frame_rendered_elsewhere():
    from gpu_util import renderer
    g = O.load_golden("cfg1_odd_333x187")  # odd size: the last MCU row / column is padded by edge replication
    snap = g["snap"]
    r = renderer()
    r.set_textures(snap, O.load_texture)
    h, w = snap.height, snap.width
    buf = r.frame_alloc(h * w * 3)
    r.render_device(snap, buf, pixel_format=abi.PIXEL_BGR8)
    r.sync()
    sink = VideoSink(r, None, w, h)
    sink.write_device(buf)
    dec = cv2.imdecode(np.frombuffer(sink.last_jpeg(), np.uint8), cv2.IMREAD_COLOR)
    sink.close()
    assert dec.shape == (h, w, 3)
    assert psnr(dec, np.ascontiguousarray(g["bgr"])) > PSNR_MIN_DB - 4.0  # small, busy frame


def _jpeg(frame, q=95):
    ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, q])
    assert ok
    return enc.tobytes()


@pytest.mark.parametrize("n_parts,total", [(2, 7), (3, 9), (4, 5), (1, 3)])
def test_merge_puts_the_per_gpu_files_back_into_frame_order(tmp_path, n_parts, total):
    """N ranks each write the frames sharding.frames_of gives them; bh8_sink_merge must return the file a
    single writer would have produced: same frame order, byte-identical bitstreams."""
    from blackhole_8_b200 import sharding
    from blackhole_8_b200.renderer import merge_video_parts
    g = O.load_golden("cfg1_odd_333x187")
    base = np.ascontiguousarray(g["bgr"])
    h, w = base.shape[:2]
    jpegs = [_jpeg(np.roll(base, 11 * k, axis=1)) for k in range(total)]
    parts = []
    for rank in range(n_parts):
        path = str(tmp_path / ("part%d.avi" % rank))
        with VideoSink(None, path, w, h, fps=29) as sink:
            for k in sharding.frames_of(total, rank, n_parts):
                sink.append_jpeg(jpegs[k])
        parts.append(path)
    single = str(tmp_path / "single.avi")
    with VideoSink(None, single, w, h, fps=29) as sink:
        for j in jpegs:
            sink.append_jpeg(j)
    merged = str(tmp_path / "video.avi")
    frames, size = merge_video_parts(parts, merged)
    assert frames == total and size == os.path.getsize(merged)
    assert open(merged, "rb").read() == open(single, "rb").read()
    info, got = read_avi(merged)
    assert len(got) == total and (info["w"], info["h"]) == (w, h)


def test_merge_rejects_parts_that_do_not_belong_together(tmp_path):
    from blackhole_8_b200.renderer import merge_video_parts
    f = np.zeros((48, 64, 3), np.uint8)
    a, b, c = (str(tmp_path / n) for n in ("a.avi", "b.avi", "c.avi"))
    with VideoSink(None, a, 64, 48) as s:
        s.append_jpeg(_jpeg(f))
    with VideoSink(None, b, 64, 48) as s:
        for _ in range(3):
            s.append_jpeg(_jpeg(f))
    with VideoSink(None, c, 32, 48) as s:
        s.append_jpeg(_jpeg(f[:, :32]))
    with pytest.raises(Bh8Error, match="round-robin"):
        merge_video_parts([a, b], str(tmp_path / "o.avi"))
    with pytest.raises(Bh8Error, match="differ"):
        merge_video_parts([a, c], str(tmp_path / "o.avi"))
    with pytest.raises(Bh8Error, match="cannot open"):
        merge_video_parts([a, str(tmp_path / "missing.avi")], str(tmp_path / "o.avi"))
    junk = tmp_path / "junk.avi"
    junk.write_bytes(b"definitely not a RIFF file")
    with pytest.raises(Bh8Error, match="not a RIFF"):
        merge_video_parts([str(junk)], str(tmp_path / "o.avi"))


@pytest.mark.gpu
def test_pipelined_submit_writes_the_same_file_as_the_synchronous_sink(tmp_path):
    from gpu_util import renderer
    g = O.load_golden("cfg1_640x360")
    snap = g["snap"]
    r = renderer()
    r.set_textures(snap, O.load_texture)
    h, w = snap.height, snap.width
    snaps = []
    for k in range(7):
        d = snap.to_dict()
        d["camera"]["pos"] = [d["camera"]["pos"][0] + 25.0 * k] + list(d["camera"]["pos"][1:])
        snaps.append(abi.SceneSnapshot.from_dict(d))
    sync_path, pipe_path = str(tmp_path / "sync.avi"), str(tmp_path / "pipe.avi")
    with VideoSink(r, sync_path, w, h) as sink:
        for s in snaps:
            sink.render(s)
    with VideoSink(r, pipe_path, w, h) as sink:
        for s in snaps[:3]:
            sink.submit(s)
        assert sink.stats()["frames"] == 1      # frame 0 left when its slot was needed again
        sink.flush()
        assert sink.stats()["frames"] == 3
        sink.render(snaps[3])                   # the synchronous call may be mixed in
        for s in snaps[4:]:
            sink.submit(s)
        # close() appends what is still in flight
    assert open(pipe_path, "rb").read() == open(sync_path, "rb").read()


@pytest.mark.gpu
def test_flythrough_video_app_writes_one_readable_file(tmp_path):
    """apps/flythrough_video.py (configs[3] end to end: per-GPU sinks + merge), single GPU, small frames,
    once from host snapshots and once from a device-side script: same frame count, both readable."""
    import json
    import subprocess
    import sys
    app = os.path.join(O.ROOT, "apps", "flythrough_video.py")
    for extra in ([], ["--script"]):
        out = str(tmp_path / ("video%d.avi" % len(extra)))
        res = subprocess.run([sys.executable, app, "--out", out, "--frames", "9", "--width", "320", "--height", "180"] + extra,
                             check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        line = json.loads(res)
        assert line["frames"] == 9 and line["frames_read_back"] == 9
        assert line["min_psnr_db_checked_frames"] > PSNR_MIN_DB - 4.0   # small, busy frames
        info, frames = read_avi(out)
        assert len(frames) == 9 and (info["w"], info["h"]) == (320, 180)
        assert not any(p.endswith(".avi") and ".part" in p for p in os.listdir(tmp_path))


def test_merge_survives_damaged_parts(tmp_path):
    """Truncated, bit-flipped and size-corrupted part files are either merged (if still well-formed) or rejected
    with an error -- never a crash; deeply nested LIST chunks are refused."""
    import struct as st
    from blackhole_8_b200.renderer import merge_video_parts
    f = np.zeros((48, 64, 3), np.uint8)
    f[10:30, 20:40] = 200
    good = str(tmp_path / "good.avi")
    with VideoSink(None, good, 64, 48) as s:
        for _ in range(5):
            s.append_jpeg(_jpeg(f))
    raw = open(good, "rb").read()
    rng = np.random.default_rng(1)
    bad, out = str(tmp_path / "bad.avi"), str(tmp_path / "out.avi")
    outcomes = {"merged": 0, "rejected": 0}
    for it in range(600):
        b = bytearray(raw)
        if it % 3 == 0:
            b = b[:int(rng.integers(0, len(b)))]
        elif it % 3 == 1:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        else:
            i = int(rng.integers(0, len(b) - 4))
            b[i:i + 4] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "little")
        open(bad, "wb").write(bytes(b))
        try:
            merge_video_parts([bad], out)
            outcomes["merged"] += 1
        except Bh8Error:
            outcomes["rejected"] += 1
    assert outcomes["merged"] > 0 and outcomes["rejected"] > 0
    depth = 5000  # RIFF AVI + LIST(LIST(LIST(...)))
    body = b""
    for _ in range(depth):
        body = b"LIST" + st.pack("<I", len(body) + 4) + b"movi" + body
    open(bad, "wb").write(b"RIFF" + st.pack("<I", len(body) + 4) + b"AVI " + body)
    with pytest.raises(Bh8Error, match="nested"):
        merge_video_parts([bad], out)
