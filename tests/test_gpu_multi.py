"""Several devices in ONE process (bh8_create with n_dev > 1): frames are dealt round-robin, a single
frame is split into interleaved row stripes that every device stores straight into device 0's frame
buffer over NVLink (peer access).  Results must equal the single-GPU render bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
from blackhole_8_b200 import abi

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def pair():
    from blackhole_8_b200.renderer import Renderer
    n = min(_device_count(), 8)
    # On a one-GPU box the context gets GPU 0 three times: three logical devices (own streams, buffers,
    # textures) run the same sharding code -- stripe ownership, round-robin, gather into device 0's frame --
    # with plain stores where a multi-GPU box has NVLink stores (profiles/r02zz_gpu_multi_tests.log is the
    # 2-GPU run).
    many_devs = tuple(range(n)) if n >= 2 else (0, 0, 0)
    one, many = Renderer((0,)), Renderer(many_devs)
    yield one, many
    one.close()
    many.close()


def test_single_frame_striped_over_devices(pair):
    one, many = pair
    for name in ("cfg1_640x360", "cfg2_640x360", "cfg1_odd_333x187"):
        snap = O.load_golden(name)["snap"]
        for r in (one, many):
            r.set_textures(snap, O.load_texture)
        a = one.render(snap, pixel_format=abi.PIXEL_RGBA8, want_maps=True, stats=True)
        b = many.render(snap, pixel_format=abi.PIXEL_RGBA8, want_maps=True, stats=True)
        for k in ("pixels", "cls", "key", "steps"):
            assert np.array_equal(a[k], b[k]), (name, k)
        assert a["stats"].steps == b["stats"].steps and a["stats"].rays == b["stats"].rays


def test_frames_dealt_round_robin(pair):
    one, many = pair
    snaps = [O.load_golden(n)["snap"] for n in ("cfg3_frame60_480x270", "cfg3_frame180_480x270")] * 5
    for r in (one, many):
        r.set_textures(snaps[0], O.load_texture)
    a = one.render(snaps, pixel_format=abi.PIXEL_BGR8)
    b = many.render(snaps, pixel_format=abi.PIXEL_BGR8)
    assert np.array_equal(a["pixels"], b["pixels"])


def test_streaming_over_devices(pair):
    one, many = pair
    snaps = [O.load_golden(n)["snap"] for n in ("cfg3_frame60_480x270", "cfg3_frame180_480x270")]
    for r in (one, many):
        r.set_textures(snaps[0], O.load_texture)
    want = [one.render(s)["pixels"][0].copy() for s in snaps]
    depth = 2 * len(many.devices)
    bufs = [many.pinned((270, 480, 4)) for _ in range(depth)]
    tickets = []
    for k in range(3 * depth):
        if len(tickets) == depth:
            t, slot, which = tickets.pop(0)
            many.wait(t)
            assert np.array_equal(bufs[slot].array, want[which])
        tickets.append((many.submit(snaps[k & 1], bufs[k % depth].array), k % depth, k & 1))
    for t, slot, which in tickets:
        many.wait(t)
        assert np.array_equal(bufs[slot].array, want[which])
