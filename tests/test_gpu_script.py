"""Scripted animation on the GPU (bh8_script_*) through the C ABI.

  * the states bh8_animate_kernel leaves in device memory are bit-identical to what the REFERENCE's
    Camera / Annulus / Rectangle classes went through (tests/golden/states, BASELINE configs[3]'s
    240 frames and the flat-space driver's "Movement test");
  * the frame constants bh8_build_frames_kernel derives on the device are byte-identical to the
    host's for the same snapshot;
  * a frame drawn from the script (no per-frame host data) is byte-identical to the frame drawn
    from the host-side snapshot, and matches the reference frames of configs[3].
"""
import numpy as np
import pytest

import oracle_lib as O
import parity
from blackhole_8_b200 import abi
from blackhole_8_b200.renderer import Bh8Error, Script
from test_script_host import disc_index, load_states

pytestmark = pytest.mark.gpu


def snapshot_of(snap0, frame_state):
    d = snap0.to_dict()
    d["camera"].update(frame_state["camera"])
    for o, v in zip(d["objects"], frame_state["v"]):
        o["v"] = v
    return abi.SceneSnapshot.from_dict(d)


def render_script_frame(r, script, k, h, w):
    bufs = {"pix": r.frame_alloc(h * w * 3), "cls": r.frame_alloc(h * w), "key": r.frame_alloc(h * w),
            "steps": r.frame_alloc(h * w * 2)}
    script.render(k, bufs["pix"], bufs["cls"], bufs["key"], bufs["steps"])
    r.sync()
    out = {"bgr": np.empty((h, w, 3), np.uint8), "cls": np.empty((h, w), np.uint8),
           "key": np.empty((h, w), np.int8), "steps": np.empty((h, w), np.uint16)}
    for name, arr in (("pix", out["bgr"]), ("cls", out["cls"]), ("key", out["key"]), ("steps", out["steps"])):
        r.memcpy_d2h(arr, bufs[name])
        r.frame_free(bufs[name])
    return out


@pytest.mark.parametrize("which", ["cfg1_spin", "cfg3_flythrough"])
def test_device_replay_matches_the_reference_classes(which):
    from gpu_util import renderer
    snap0, frames = load_states(which)
    r = renderer()
    r.set_textures(snap0, O.load_texture)
    acts = abi.reference_script(which, len(frames), disc_index(snap0))
    with Script(r, snap0, acts, len(frames)) as sc:
        for k, fr in enumerate(frames):
            cam, objs = sc.state(k)
            for name in ("pos", "vx", "vy", "vz"):
                assert list(getattr(cam, name)) == fr["camera"][name], (which, k, name)
            for j, o in enumerate(objs):
                assert [x for row in o.v for x in row] == fr["v"][j], (which, k, j)
        # frame constants: device-built == host-built, byte for byte
        for k in (0, 1, 59, 120, 121, 239):
            host = r.host_frame_constants(snapshot_of(snap0, frames[k]))
            dev = sc.frame_constants(k)
            if dev != host:
                a, b = np.frombuffer(dev, np.uint8), np.frombuffer(host, np.uint8)
                bad = np.nonzero(a != b)[0]
                raise AssertionError("%s frame %d: %d bytes differ, first at offset %d" % (which, k, bad.size, bad[0]))


@pytest.mark.parametrize("frame,golden", [(60, "cfg3_frame60_480x270"), (180, "cfg3_frame180_480x270")])
def test_script_frames_match_the_reference_and_the_host_path(frame, golden):
    from gpu_util import gpu_render, renderer
    g = O.load_golden(golden)
    h, w = g["snap"].height, g["snap"].width
    snap0, frames = load_states("cfg3_flythrough")
    snap0 = snap0.with_resolution(w, h)
    r = renderer()
    r.set_textures(snap0, O.load_texture)
    acts = abi.reference_script("cfg3_flythrough", frame + 1, disc_index(snap0))
    with Script(r, snap0, acts, frame + 1, pixel_format=abi.PIXEL_BGR8) as sc:
        got = render_script_frame(r, sc, frame, h, w)
    rep = parity.assert_parity(got, g, golden)
    print(golden, rep)
    host = gpu_render(g["snap"], stats=False)
    for name in ("bgr", "cls", "key", "steps"):
        assert np.array_equal(got[name], host[name]), name


def test_script_on_the_flat_space_tracer():
    """The "Movement test" of ray_tracer_test.cc:237-261 drawn by the linear tracer from a script."""
    import json
    import os
    from gpu_util import gpu_render, renderer
    from test_script_host import STATES
    snap0, frames = load_states("cfg10_movers")
    with open(os.path.join(STATES, "cfg10_movers.json")) as f:
        movers = json.load(f)["movers"]
    pi = 3.14159265358979323846
    acts = []
    for k in range(25):
        acts += [abi.Action(k, movers[0], abi.OP_MOVE_Y, 0, 60.0), abi.Action(k, movers[0], abi.OP_MOVE_Z, 0, 50.0),
                 abi.Action(k, movers[0], abi.OP_ROTATE_X, 0, pi / 180),
                 abi.Action(k, movers[0], abi.OP_MOVE_Y, 0, -60.0), abi.Action(k, movers[0], abi.OP_MOVE_Z, 0, -50.0)]
    r = renderer()
    r.set_textures(snap0, O.load_texture)
    h, w = snap0.height, snap0.width
    with Script(r, snap0, acts, 26, pixel_format=abi.PIXEL_BGR8) as sc:
        got = render_script_frame(r, sc, 25, h, w)
        cam, objs = sc.state(25)
    d = snap0.to_dict()
    for o, q in zip(d["objects"], objs):
        o["v"] = [x for row in q.v for x in row]
    host = gpu_render(abi.SceneSnapshot.from_dict(d), stats=False)
    for name in ("bgr", "cls", "key", "steps"):
        assert np.array_equal(got[name], host[name]), name
    assert (got["cls"] == abi.CLASS_OBJECT).any()


def test_script_errors():
    from gpu_util import renderer
    snap0, _ = load_states("cfg1_spin")
    r = renderer()
    r.set_textures(snap0, O.load_texture)
    with pytest.raises(Bh8Error, match="target"):
        Script(r, snap0, [abi.Action(0, 99, abi.OP_MOVE_X, 0, 1.0)], 4)
    with pytest.raises(Bh8Error, match="sorted"):
        Script(r, snap0, [abi.Action(2, 0, abi.OP_MOVE_X, 0, 1.0), abi.Action(1, 0, abi.OP_MOVE_X, 0, 1.0)], 4)
    with Script(r, snap0, [], 2) as sc:
        buf = r.frame_alloc(snap0.height * snap0.width * 4)
        with pytest.raises(Bh8Error, match="out of range"):
            sc.render(2, buf)
        sc.render(1, buf)
        r.sync()
        r.frame_free(buf)


def test_graph_captured_range_draws_the_same_frames():
    """bh8_script_render_range: 12 frames with one host launch (CUDA graph) == 12 bh8_script_render calls;
    a second launch re-uses the graph, a different range rebuilds it."""
    from gpu_util import renderer
    snap0, frames = load_states("cfg3_flythrough")
    snap0 = snap0.with_resolution(320, 180)
    h, w = 180, 320
    r = renderer()
    r.set_textures(snap0, O.load_texture)
    acts = abi.reference_script("cfg3_flythrough", 140, disc_index(snap0))
    fb = h * w * 4
    with Script(r, snap0, acts, 140) as sc:
        one = r.frame_alloc(fb)
        many = r.frame_alloc(12 * fb)
        for first in (0, 118, 118, 3):
            launches = r.launches
            r.memset_d(many, 0, 12 * fb)
            sc.render_range(first, 12, many, fb)
            r.sync()
            assert r.launches - launches == 12
            got = np.empty((12, h, w, 4), np.uint8)
            r.memcpy_d2h(got, many)
            for k in range(12):
                sc.render(first + k, one)
                r.sync()
                ref = np.empty((h, w, 4), np.uint8)
                r.memcpy_d2h(ref, one)
                assert np.array_equal(got[k], ref), (first, k)
        with pytest.raises(Bh8Error, match="out of range"):
            sc.render_range(130, 12, many, fb)
        with pytest.raises(Bh8Error, match="stride"):
            sc.render_range(0, 2, many, fb - 4)
        r.frame_free(one)
        r.frame_free(many)
