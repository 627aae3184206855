"""Scripted animation on the CPU: the per-entity replay code bh8_animate_kernel runs
(blackhole_8_b200/csrc/bh8_anim.cuh), compiled for the host by the test-only harness, against the
states the REFERENCE's own Camera / Annulus / Rectangle classes went through
(tests/golden/states/*.json, written by tools/make_flythrough.py from oracle/_ref/ref_render).
Bit-exact: every double must be identical."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from blackhole_8_b200 import abi
from test_ray_math_host import harness

STATES = os.path.join(O.ROOT, "tests", "golden", "states")


def load_states(name):
    with open(os.path.join(STATES, name + ".json")) as f:
        d = json.load(f)
    return abi.SceneSnapshot.from_dict(d["base"]), d["frames"]


def host_replay(snap0, actions, n_frames, basis=None):
    L = harness()
    n_obj = snap0.scene.n_obj
    cams = (abi.Camera * n_frames)()
    objs = (abi.Object * (n_frames * n_obj))()
    acts = (abi.Action * max(1, len(actions)))(*actions)
    bas = (abi.Basis * n_obj)(*basis) if basis is not None else None
    err = C.create_string_buffer(256)
    rc = L.bh8_harness_replay(C.byref(snap0.scene), bas, C.byref(snap0.camera), acts, len(actions), n_frames,
                              cams, objs, err)
    return rc, err.value.decode(), cams, objs


def assert_states_equal(frames, cams, objs, n_obj, label):
    for k, fr in enumerate(frames):
        for name in ("pos", "vx", "vy", "vz"):
            got = list(getattr(cams[k], name))
            assert got == fr["camera"][name], "%s frame %d camera.%s: %r != %r" % (label, k, name, got, fr["camera"][name])
        for j in range(n_obj):
            got = [x for row in objs[k * n_obj + j].v for x in row]
            assert got == fr["v"][j], "%s frame %d object %d vertices differ" % (label, k, j)


def disc_index(snap):
    return [o.kind for o in snap.objects].index(abi.KIND_ANNULUS)


@pytest.mark.parametrize("which", ["cfg1_spin", "cfg3_flythrough"])
def test_replay_reproduces_the_reference_classes_bit_for_bit(which):
    snap0, frames = load_states(which)
    n = len(frames)
    assert n == 240
    acts = abi.reference_script(which, n, disc_index(snap0))
    rc, err, cams, objs = host_replay(snap0, acts, n)
    assert rc == 0, err
    assert_states_equal(frames, cams, objs, snap0.scene.n_obj, which)
    # everything Snapshot() copies besides the vertices is carried along unchanged
    for k in (0, n - 1):
        for j, o0 in enumerate(snap0.objects):
            o = objs[k * snap0.scene.n_obj + j]
            assert (o.kind, o.key, o.tex_id, o.pattern, o.mass, o.r_in, o.r_out) == \
                   (o0.kind, o0.key, o0.tex_id, o0.pattern, o0.mass, o0.r_in, o0.r_out)
            assert list(o.n) == list(o0.n)  # Annulus::norm_ is NOT rotated (vector_object.h:325)
        assert cams[k].focus_len == snap0.camera.focus_len and cams[k].width == snap0.camera.width


def test_all_ops_against_the_reference_movement_test():
    """ray_tracer_test.cc:237-261 ("Movement test"): MoveX/Y/Z, RotateX, RotateY on three rectangles,
    states from the reference's classes (tests/golden/states/cfg10_movers.json)."""
    snap0, frames = load_states("cfg10_movers")
    with open(os.path.join(STATES, "cfg10_movers.json")) as f:
        movers = json.load(f)["movers"]  # object indices of mooni, karina, winter
    pi = 3.14159265358979323846
    acts, direction = [], 1
    for k in range(len(frames) - 1):
        A = lambda t, op, amt: acts.append(abi.Action(k, t, op, 0, float(amt)))  # noqa: E731
        A(movers[1], abi.OP_MOVE_X, 3 * direction)
        x = frames[k + 1]["v"][movers[1]][0]  # position()[0] after the move
        if x > 100:
            direction = -1
        elif x < 0:
            direction = 1
        y, z = -120.0, -100.0
        A(movers[0], abi.OP_MOVE_Y, -y / 2)
        A(movers[0], abi.OP_MOVE_Z, -z / 2)
        A(movers[0], abi.OP_ROTATE_X, pi / 180)
        A(movers[0], abi.OP_MOVE_Y, y / 2)
        A(movers[0], abi.OP_MOVE_Z, z / 2)
        x, z = 100.0, -100.0
        A(movers[2], abi.OP_MOVE_X, -x)
        A(movers[2], abi.OP_MOVE_Z, -z / 2)
        A(movers[2], abi.OP_ROTATE_Y, pi / 120)
        A(movers[2], abi.OP_MOVE_X, x)
        A(movers[2], abi.OP_MOVE_Z, z / 2)
    rc, err, cams, objs = host_replay(snap0, acts, len(frames))
    assert rc == 0, err
    assert_states_equal(frames, cams, objs, snap0.scene.n_obj, "cfg10_movers")


def test_move_to_and_plane_basis():
    """MoveTo moves vertex()[0] only (object.h:82-83); an InfinitePlane's (ex, ey, n) follow its basis."""
    snap0 = O.load_golden("cfg2_640x360")["snap"]
    plane = [o.kind for o in snap0.objects].index(abi.KIND_INFINITE_PLANE)
    rect = [o.kind for o in snap0.objects].index(abi.KIND_RECTANGLE)
    acts = [abi.Action(0, rect, abi.OP_MOVE_TO, 0, 0.0, abi.Vec3(1.0, 2.0, 3.0)),
            abi.Action(0, plane, abi.OP_ROTATE_X, 0, 0.25),
            abi.Action(1, abi.TARGET_CAMERA, abi.OP_MOVE_TO, 0, 0.0, abi.Vec3(-5.0, 6.0, 7.0))]
    rc, err, cams, objs = host_replay(snap0, acts, 3)
    assert rc == 0, err
    n = snap0.scene.n_obj
    r0, r1 = objs[rect], objs[n + rect]
    assert list(r1.v[0]) == [1.0, 2.0, 3.0] and [list(x) for x in r1.v][1:] == [list(x) for x in r0.v][1:]
    p0, p1 = objs[plane], objs[n + plane]
    assert list(p1.ex) == list(p0.ex)                      # the rotation axis
    c, s = np.cos(0.25), np.sin(0.25)
    ey0, n0 = np.array(list(p0.ey)), np.array(list(p0.n))
    ex0 = np.array(list(p0.ex))
    assert np.allclose(list(p1.ey), c * ey0 + s * np.cross(ex0, ey0), atol=1e-15)
    assert np.allclose(list(p1.n), c * n0 + s * np.cross(ex0, n0), atol=1e-15)
    assert list(cams[1].pos) == list(snap0.camera.pos) and list(cams[2].pos) == [-5.0, 6.0, 7.0]


def test_malformed_scripts_are_rejected():
    snap0, _ = load_states("cfg1_spin")
    for bad in ([abi.Action(0, 99, abi.OP_MOVE_X, 0, 1.0)],
                [abi.Action(0, 0, 17, 0, 1.0)],
                [abi.Action(3, 0, abi.OP_MOVE_X, 0, 1.0), abi.Action(2, 0, abi.OP_MOVE_X, 0, 1.0)],
                [abi.Action(-1, 0, abi.OP_MOVE_X, 0, 1.0)]):
        rc, err, _, _ = host_replay(snap0, bad, 4)
        assert rc == abi.EINVAL and err


def test_cbrt_literal_is_what_libm_returns():
    """bh8_frame.h hard-codes cbrt(DBL_EPSILON) so host and device start SolveG's interval from the same bits."""
    libm = C.CDLL("libm.so.6")
    libm.cbrt.restype = C.c_double
    libm.cbrt.argtypes = [C.c_double]
    assert libm.cbrt(2.220446049250313e-16) == float.fromhex("0x1.965fea53d6e3dp-18")
