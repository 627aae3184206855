"""Parity on seeded random scenes (tests/random_scenes.py): the kernel's per-ray code against the C
oracle, which follows the reference's formulas literally and is pinned byte-for-byte to the reference
on every fixture (tests/test_oracle.py).  CPU: the code compiled for the host (tests/host_harness);
-m gpu: the real kernel through the C ABI.  Tolerance as everywhere: north_star's."""
import numpy as np
import pytest

import oracle_lib as O
import parity
from random_scenes import random_snapshot

SEEDS = list(range(100))
GPU_SEEDS = SEEDS[:48] + list(range(100, 250))  # (the campaign of tools/gpu_random_campaign.py goes far beyond)


def check(got, ref, what):
    n = ref["cls"].size
    cls_bad = int((got["cls"] != ref["cls"]).sum())
    diff = np.abs(got["bgr"].astype(int) - ref["bgr"].astype(int)).max(axis=2)
    rgb_bad = int(((got["cls"] == ref["cls"]) & (diff > parity.RGB_TOL)).sum())
    # small frames: the 99.9 % bar of north_star would allow only 5 pixels of 5184; state it per pixel count
    assert cls_bad <= max(2, int(0.001 * n)), (what, "class mismatches", cls_bad)
    assert rgb_bad <= max(3, int(0.001 * n)), (what, "rgb outliers", rgb_bad)
    steps_bad = int((got["steps"] != ref["steps"]).sum())
    assert steps_bad <= max(2, int(0.001 * n)), (what, "step count mismatches", steps_bad)
    return cls_bad, rgb_bad, steps_bad


@pytest.mark.parametrize("seed", SEEDS)
def test_host_compiled_kernel_code_matches_oracle_on_random_scenes(seed):
    from test_ray_math_host import harness_render
    snap = random_snapshot(seed)
    ref = O.render(snap)
    got = harness_render(snap)
    print(seed, check(got, ref, "seed %d" % seed), "classes", np.bincount(ref["cls"].ravel(), minlength=4).tolist())


@pytest.mark.parametrize("seed", SEEDS[:40])
def test_emulated_warp_schedule_matches_oracle_on_random_scenes(seed):
    """The same scenes through the kernel's warp schedule emulated on the CPU (tests/host_harness): lanes in
    lockstep, batched exact tests, register parking."""
    from test_ray_math_host import harness_render_warps
    snap = random_snapshot(seed)
    check(harness_render_warps(snap), O.render(snap), "seed %d (warp schedule)" % seed)


def test_random_scenes_cover_every_class_and_both_plane_kinds():
    seen = np.zeros(4, int)
    central = noncentral = 0
    for seed in SEEDS:
        snap = random_snapshot(seed)
        seen += np.bincount(O.render(snap)["cls"].ravel(), minlength=4)
        bh = np.array(list(snap.objects[snap.scene.bh_index].v[0]))
        for o in snap.objects:
            if o.kind == 1:
                c = abs(np.dot(np.array(list(o.n)), bh - np.array(list(o.v[0]))))
                central += c == 0.0
                noncentral += c != 0.0
    assert (seen > 500).all(), seen
    assert central >= 10 and noncentral >= 10


@pytest.mark.gpu
@pytest.mark.parametrize("seed", GPU_SEEDS)
def test_gpu_matches_oracle_on_random_scenes(seed):
    from gpu_util import gpu_render
    snap = random_snapshot(seed, 192, 108)
    ref = O.render(snap)
    got = gpu_render(snap, stats=False)
    print(seed, check(got, ref, "seed %d" % seed))


# Scenes a 1 500-seed GPU campaign (tools/gpu_random_campaign.py, profiles/r02zzz_random_campaign.txt) found: rays
# that leave the setup under a lease from their start point, freeze for an exact test inside the inbound gated
# range and travel on.  They came back with the lease's lo = 0 as their base range and skipped filter (2) on the
# gated steps that were left -- a rectangle met there was missed (up to 0.26 % of a frame's pixels).
LEASE_THEN_FREEZE = [(495, 192, 108), (823, 256, 144), (1364, 333, 187), (1470, 192, 108), (1013, 333, 187),
                     (1036, 256, 144)]


@pytest.mark.parametrize("seed,width,height", LEASE_THEN_FREEZE)
def test_ray_frozen_after_a_start_lease_keeps_its_base_range(seed, width, height):
    from test_ray_math_host import harness_render, harness_render_warps
    snap = random_snapshot(seed, width, height)
    ref = O.render(snap)
    for got in (harness_render(snap), harness_render_warps(snap)):
        assert int((got["cls"] != ref["cls"]).sum()) == 0
        assert int((got["steps"] != ref["steps"]).sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed,width,height", LEASE_THEN_FREEZE)
def test_gpu_ray_frozen_after_a_start_lease_keeps_its_base_range(seed, width, height):
    from gpu_util import gpu_render
    snap = random_snapshot(seed, width, height)
    ref = O.render(snap)
    for stats in (False, True):
        got = gpu_render(snap, stats=stats)
        assert int((got["cls"] != ref["cls"]).sum()) == 0, stats
        assert int((got["steps"] != ref["steps"]).sum()) <= 1, stats


def test_cpu_campaign_of_random_scenes_at_mixed_sizes():
    """600 more scenes at five frame sizes through the host-compiled kernel code (every fourth one through the
    emulated warp schedule too) -- tools/cpu_random_campaign.py in small.  Before the fix above this stretch of
    seeds held five failing scenes (237, 247, 287 at 160x90; 495 and 500 at 96x54): the campaign needs no GPU
    to find such a bug."""
    from test_ray_math_host import harness_render, harness_render_warps
    bad = []
    for seed in range(100, 700):
        w, h = ((96, 54), (128, 72), (160, 90), (192, 108), (64, 36))[seed % 5]
        snap = random_snapshot(seed, w, h)
        ref = O.render(snap)
        got = [harness_render(snap)] + ([harness_render_warps(snap)] if seed % 4 == 0 else [])
        for g in got:
            d = int((g["cls"] != ref["cls"]).sum()) + int((g["steps"] != ref["steps"]).sum())
            if d:
                bad.append((seed, w, h, d))
    assert not bad, bad


def test_damaged_snapshots_never_hang_the_ray_code():
    """NaN, infinities, zeros and huge values anywhere in a snapshot: the per-ray state machine (the code the
    kernel's lanes run) must still terminate for every pixel -- on the GPU a lane that never ends would hang
    the frame.  The frames themselves are meaningless; only termination (and a clean error for snapshots the
    frame builder rejects) is asserted."""
    import signal
    from blackhole_8_b200 import abi
    from test_ray_math_host import harness_render
    special = [float("nan"), float("inf"), -float("inf"), 0.0, -0.0, 1e300, -1e300, 1e-300, 5e-324, 1e18, -1e18]
    rng = np.random.default_rng(7)

    class Hang(Exception):
        pass

    def on_alarm(sig, frame):
        raise Hang()

    old = signal.signal(signal.SIGALRM, on_alarm)
    rendered = rejected = 0
    try:
        for it in range(300):
            d = random_snapshot(int(rng.integers(0, 100)), 16, 9).to_dict()
            for _ in range(int(rng.integers(1, 4))):
                v = special[int(rng.integers(0, len(special)))]
                where = int(rng.integers(0, 3))
                if where == 0:
                    d["camera"][["pos", "vx", "vy", "vz"][int(rng.integers(0, 4))]][int(rng.integers(0, 3))] = v
                elif where == 1:
                    d["camera"]["focus_len"] = v
                else:
                    o = d["objects"][int(rng.integers(0, len(d["objects"])))]
                    f = ["v", "n", "ex", "ey", "r_in", "r_out", "mass", "pattern_size"][int(rng.integers(0, 8))]
                    if isinstance(o[f], list):
                        o[f][int(rng.integers(0, len(o[f])))] = v
                    else:
                        o[f] = v
            signal.alarm(30)
            try:
                harness_render(abi.SceneSnapshot.from_dict(d))
                rendered += 1
            except AssertionError:  # bh8_build_frame said no (mass <= 0, zero normal, ...)
                rejected += 1
            finally:
                signal.alarm(0)
    finally:
        signal.signal(signal.SIGALRM, old)
    assert rendered > 200 and rejected > 0
