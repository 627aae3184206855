"""The per-ray source the CUDA kernel inlines (blackhole_8_b200/csrc/bh8_ray.cuh), compiled for the
host by a TEST-ONLY harness and compared with the reference's frames (tests/golden, produced by the
reference's own classes).  This checks the restructured algorithm -- closed-form ray setup,
phi-crossing / distance / horizon filters, acos-free texture index -- on the CPU box; the same
comparison runs against the real kernel in tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import parity
from blackhole_8_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class HarnessTexture(C.Structure):
    _fields_ = [("bgr", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32)]


_VARIANTS = {}


def harness(defines=()):
    """The test-only host build of the kernel's per-ray code; `defines` selects a build variant
    (e.g. ("BH8_ANNULUS_GATE=1",)), each in its own library."""
    global _LIB
    key = tuple(defines)
    if key not in _VARIANTS:
        out = os.path.join(HERE, "host_harness", "_build")
        os.makedirs(out, exist_ok=True)
        tag = "".join(c if c.isalnum() else "_" for c in "_".join(key))
        so = os.path.join(out, "libharness%s.so" % (("_" + tag) if tag else ""))
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off"] +
                       ["-D" + d for d in key] +
                       ["-I" + os.path.join(O.ROOT, "include"),
                        "-I" + os.path.join(O.ROOT, "blackhole_8_b200", "csrc"),
                        os.path.join(HERE, "host_harness", "harness.cc"), "-o", so], check=True)
        _VARIANTS[key] = C.CDLL(so)
    if not key:
        _LIB = _VARIANTS[key]
    return _VARIANTS[key]


def harness_render(snap, nstep=None, filter_slots=4, defines=()):
    L = harness(defines)
    h, w = snap.height, snap.width
    texs = [O.load_texture(n) for n in snap.textures]
    tarr = (HarnessTexture * max(1, len(texs)))()
    for i, t in enumerate(texs):
        tarr[i] = HarnessTexture(t.ctypes.data, t.shape[0], t.shape[1])
    prm = snap.params(abi.PIXEL_BGR8, 0, nstep)
    out = {"bgr": np.zeros((h, w, 3), np.uint8), "cls": np.zeros((h, w), np.uint8),
           "key": np.zeros((h, w), np.int8), "steps": np.zeros((h, w), np.uint16)}
    err = C.create_string_buffer(256)
    counters = (C.c_uint64 * 2)()
    rc = L.bh8_harness_render(C.byref(snap.scene), C.byref(snap.camera), C.byref(prm), tarr, len(texs),
                              C.c_int(filter_slots),
                              out["bgr"].ctypes.data_as(C.c_void_p), out["cls"].ctypes.data_as(C.c_void_p),
                              out["key"].ctypes.data_as(C.c_void_p), out["steps"].ctypes.data_as(C.c_void_p),
                              counters, err)
    assert rc == 0, err.value
    out["updates"], out["exact_tests"] = int(counters[0]), int(counters[1])
    L.bh8_harness_take_filter_evaluations.restype = C.c_ulonglong
    out["filter_evaluations"] = int(L.bh8_harness_take_filter_evaluations())
    return out


def harness_render_warps(snap, nstep=None, filter_slots=4, updates_per_vote=2, resolve_wait=2, defines=()):
    """The kernel's WARP schedule emulated on the CPU (32 lanes in lockstep, votes, batched exact tests with
    register parking: bh8_warp.cuh, the code render_tile inlines)."""
    L = harness(defines)
    h, w = snap.height, snap.width
    texs = [O.load_texture(n) for n in snap.textures]
    tarr = (HarnessTexture * max(1, len(texs)))()
    for i, t in enumerate(texs):
        tarr[i] = HarnessTexture(t.ctypes.data, t.shape[0], t.shape[1])
    prm = snap.params(abi.PIXEL_BGR8, 0, nstep)
    out = {"bgr": np.zeros((h, w, 3), np.uint8), "cls": np.zeros((h, w), np.uint8),
           "key": np.zeros((h, w), np.int8), "steps": np.zeros((h, w), np.uint16)}
    err = C.create_string_buffer(256)
    counters = (C.c_uint64 * 2)()
    rc = L.bh8_harness_render_warps(C.byref(snap.scene), C.byref(snap.camera), C.byref(prm), tarr, len(texs),
                                    C.c_int(filter_slots), C.c_int(updates_per_vote), C.c_int(resolve_wait),
                                    out["bgr"].ctypes.data_as(C.c_void_p), out["cls"].ctypes.data_as(C.c_void_p),
                                    out["key"].ctypes.data_as(C.c_void_p), out["steps"].ctypes.data_as(C.c_void_p),
                                    counters, err)
    assert rc == 0, err.value
    n_warps = ((h + 3) // 4) * ((w + 7) // 8)
    out["update_slots_per_warp"] = counters[0] / n_warps
    out["resolve_passes_per_warp"] = counters[1] / n_warps
    return out


@pytest.mark.parametrize("name", O.golden_names(full=False))
def test_warp_schedule_draws_the_same_frames(name):
    """Lanes that wait frozen while their neighbours travel, exact tests run in batches, stepping values
    parked in the mailbox around the test: none of it may change a pixel, a hit key or a step count
    against the one-lane-at-a-time run -- for the kernel's schedule and for other batching windows."""
    g = O.load_golden(name)
    if g["snap"].linear_steps > 0:
        pytest.skip("flat-space scene: the linear kernel has no warp schedule")
    one = harness_render(g["snap"])
    for upv, wait in ((2, 2), (1, 0), (3, 5), (2, -1)):
        got = harness_render_warps(g["snap"], updates_per_vote=upv, resolve_wait=wait)
        for k in ("bgr", "cls", "key", "steps"):
            assert np.array_equal(got[k], one[k]), (name, upv, wait, k)
    got = harness_render_warps(g["snap"])
    print(name, "update slots per warp %.1f (mean steps per ray %.1f), resolve passes per warp %.2f" %
          (got["update_slots_per_warp"], one["steps"].mean(), got["resolve_passes_per_warp"]))


@pytest.mark.parametrize("name", O.golden_names(full=False))
def test_kernel_ray_math_matches_reference_frames(name):
    g = O.load_golden(name)
    got = harness_render(g["snap"])
    rep = parity.assert_parity(got, g, name)
    print(name, rep, "exact tests per ray: %.3f" % (got["exact_tests"] / got["cls"].size))
    assert rep["steps_agreement"] > 0.999


@pytest.mark.parametrize("name", ["cfg1_640x360", "cfg2_640x360", "cfg5_480x270"])
def test_generic_instantiation_matches_too(name):
    """Scenes with more non-central planes than FP32 filter slots take the generic code path
    (exact test on every gated step); force it here and require the same picture."""
    g = O.load_golden(name)
    a = harness_render(g["snap"])
    b = harness_render(g["snap"], filter_slots=-1)
    parity.assert_parity(b, g, name)
    assert np.array_equal(a["cls"], b["cls"]) and np.array_equal(a["steps"], b["steps"])
    assert b["exact_tests"] > a["exact_tests"]


def test_fast_turning_point_equals_bisection_and_reference():
    """solve_turning_point (closed form + Newton + grid tests) must return what the 20-step bisection
    returns -- the kernel's literal copy of it AND the reference's StaticBlackhole::SolveG (via the C
    oracle) -- for impact parameters from b_c up to far field, including values next to b_c."""
    L = harness()
    rng = np.random.default_rng(8)
    for M in (10.0, 20.0, 1.0, 3.7):
        b_c = 3.0 * np.sqrt(3.0) * M
        b = np.concatenate([
            b_c * (1.0 + np.logspace(-13, 2.5, 4000)),
            b_c * (1.0 + rng.random(4000) * 0.01),
            b_c * (1.0 + rng.random(20000) * 60.0),
            [b_c, np.nextafter(b_c, np.inf), 1e6 * M],
        ])
        fast = np.empty_like(b)
        bis = np.empty_like(b)
        L.bh8_harness_solve(C.c_double(M), b.ctypes.data_as(C.c_void_p), C.c_int(len(b)),
                            fast.ctypes.data_as(C.c_void_p), bis.ctypes.data_as(C.c_void_p))
        ref = np.array([O.lib().bh8_oracle_solve_g(M, float(x)) for x in b])
        grid = (1.0 / (3 * M) - np.cbrt(np.finfo(float).eps)) / 2 ** 20
        assert np.all(np.abs(bis - ref) <= 1e-3 * grid)   # the kernel's bisection == the reference's
        bad = np.abs(fast - bis) > 1e-3 * grid
        assert bad.sum() == 0, (M, b[bad][:5], fast[bad][:5], bis[bad][:5])


def test_sincos_primitive_is_within_one_ulp():
    L = harness()
    rng = np.random.default_rng(3)
    x = np.concatenate([(rng.random(2_000_000) - 0.3) * 200.0, (rng.random(200_000) - 0.5) * 2.0,
                        np.arange(-60, 61) * (np.pi / 4), [0.0, 1e-300, -1e-9]])
    s = np.empty_like(x)
    c = np.empty_like(x)
    L.bh8_harness_sincos(x.ctypes.data_as(C.c_void_p), C.c_int(len(x)), s.ctypes.data_as(C.c_void_p),
                         c.ctypes.data_as(C.c_void_p))
    assert np.abs(s - np.sin(x)).max() <= 2.3e-16
    assert np.abs(c - np.cos(x)).max() <= 2.3e-16


def test_fast_atan2f_error_is_inside_the_arming_margin():
    """arm_central's FP32 atan2 must stay well inside kArmMargin (8e-6 rad)."""
    L = harness()
    rng = np.random.default_rng(11)
    ang = np.concatenate([rng.random(2_000_000) * 2 * np.pi - np.pi,
                          np.arange(-8, 9) * (np.pi / 8), np.arange(-8, 9) * (np.pi / 8) + 1e-6])
    rad = 10.0 ** (rng.random(len(ang)) * 12 - 6)
    x = (rad * np.cos(ang)).astype(np.float32)
    y = (rad * np.sin(ang)).astype(np.float32)
    out = np.empty(len(ang), np.float32)
    L.bh8_harness_atan2f(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int(len(ang)),
                         out.ctypes.data_as(C.c_void_p))
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    err = np.abs(out.astype(np.float64) - ref)
    err = np.minimum(err, 2 * np.pi - err)  # +pi and -pi are the same direction
    assert err.max() < 1e-6, err.max()


@pytest.mark.parametrize("name", ["cfg1_640x360", "cfg2_640x360", "cfg9_inside_320x180", "cfg6_offplane_400x225"])
def test_frozen_lanes_are_inert(name):
    """The kernel's warps run the straight-line update for every lane, frozen (parked / ended) ones
    included.  Emulate that: apply 5 extra updates to each frozen lane before its exact test and after
    its end -- the frame must not change in any byte."""
    g = O.load_golden(name)
    L = harness()
    a = harness_render(g["snap"])
    L.bh8_harness_set_frozen_updates(5)
    try:
        b = harness_render(g["snap"])
    finally:
        L.bh8_harness_set_frozen_updates(0)
    for k in ("bgr", "cls", "key", "steps"):
        assert np.array_equal(a[k], b[k]), k


def test_flat_space_scene_may_contain_a_hole_as_a_black_sphere():
    """ray_tracer_test.cc's scene with a StaticBlackhole dropped in: FindCollision treats it as a sphere of
    radius 2M (blackhole_solution.h:65-88), the linear tracer bends nothing, so bh_index stays -1.  The frame
    builder used to reject the scene ('more than one black hole')."""
    g = O.load_golden("cfg10_flat_800x450")
    d = g["snap"].to_dict()
    d["width"], d["height"] = 200, 112
    d["camera"].update(width=200, height=112, focus_len=d["camera"]["focus_len"] * 200 / 800)
    hole = dict(d["objects"][0], kind=abi.KIND_BLACKHOLE, key=9, tex_id=-1, pattern=0, mass=30.0,
                v=[-150.0, 60.0, 60.0] + [0.0] * 12)
    d["objects"] = [hole] + d["objects"]
    assert d["bh_index"] == -1
    snap = abi.SceneSnapshot.from_dict(d)
    got = harness_render(snap)
    ref = O.render(snap)
    assert (ref["cls"] == abi.CLASS_HORIZON).sum() > 50  # the sphere is in view
    for k in ("bgr", "cls", "key", "steps"):
        assert np.array_equal(got[k], ref[k]), k
    # two spheres: the kernels hold one radius, so this must be refused, with the right message
    d["objects"] = [dict(hole, key=10, v=[0.0, 300.0, 0.0] + [0.0] * 12)] + d["objects"]
    with pytest.raises(AssertionError, match="more than one black hole"):
        harness_render(abi.SceneSnapshot.from_dict(d))


@pytest.mark.parametrize("name", ["cfg1_odd_333x187", "cfg2_640x360"])
def test_step_counts_other_than_the_references(name):
    """nstep is a parameter of the path (blackhole_solution_test.cc:202): the leg changes sit at nstep - 1,
    nstep and 2 nstep - 1, adjacent for small nstep, and the turn's short step is taken inside its event
    (lane_event, fuse).  Hit class and step count equal the oracle's for odd, tiny and large step counts,
    one lane at a time and under the warp schedule."""
    snap = O.load_golden(name)["snap"]
    for n in (2, 3, 7, 33):
        ref = O.render(snap, nstep=n)
        one = harness_render(snap, nstep=n)
        got = harness_render_warps(snap, nstep=n)
        for k in ("bgr", "cls", "key", "steps"):
            assert np.array_equal(got[k], one[k]), (name, n, k)
        assert np.array_equal(one["cls"], ref["cls"]), (name, n)
        assert np.array_equal(one["steps"], ref["steps"]), (name, n)
        off = np.abs(one["bgr"].astype(int) - ref["bgr"].astype(int)).max(axis=2) > 2
        assert off.mean() <= 1e-3, (name, n, off.mean())  # north_star: RGB within 2/255 on >= 99.9 % of pixels
