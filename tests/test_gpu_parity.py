"""GPU parity: the CUDA path, called through the C ABI (bh8_render), against the reference.

Small sizes compare with the committed reference frames (tests/golden, made by the reference's own
classes); BASELINE.json's full sizes compare with the C oracle run on the box's host cores (itself
pinned byte-for-byte to the reference at those sizes by tests/test_oracle.py) and through
size-independent properties.  Tolerance is north_star's: hit class identical on >= 99.9 % of
pixels, BGR within 2/255 on matching pixels (texel flips on exact texel boundaries bounded at
0.1 % of pixels), residual reported.
"""
import numpy as np
import pytest

import oracle_lib as O
import parity
from blackhole_8_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("name", O.golden_names(full=False))
def test_small_frames_match_reference(G, name):
    g = O.load_golden(name)
    got = G.gpu_render(g["snap"])
    rep = parity.assert_parity(got, g, name)
    print(name, rep)
    assert rep["key_agreement"] >= 0.999
    assert rep["steps_agreement"] >= 0.999
    # no ray may run past the end of its last leg (a missed leg-change event shows up here first)
    if not g["snap"].linear_steps:
        assert int(got["steps"].max()) <= 2 * g["snap"].nstep - 1
    # the launch without device counters is a different instantiation of the kernel: same frame, byte for byte
    plain = G.gpu_render(g["snap"], stats=False)
    for k in ("bgr", "cls", "key", "steps"):
        assert np.array_equal(plain[k], got[k]), k
    st = got["stats"]
    assert st.rays == g["snap"].width * g["snap"].height
    assert st.steps == int(got["steps"].sum())
    assert list(st.class_count) == [int((got["cls"] == c).sum()) for c in range(4)]
    assert st.tex_oob == 0
    assert abs(st.steps - g["run"]["steps"]) <= 1e-4 * g["run"]["steps"]


@pytest.mark.parametrize("name", ["cfg1_1920x1080", "cfg2_1920x1080", "cfg0_1920x1080", "cfg3_frame239_1920x1080"])
def test_full_size_frames_match_oracle(G, name):
    g = O.load_golden(name)
    ref = O.render(g["snap"])
    assert O.digest(ref["bgr"]) == g["digest"]["bgr"]  # oracle == reference at this size
    got = G.gpu_render(g["snap"])
    rep = parity.assert_parity(got, ref, name)
    print(name, rep)
    assert rep["steps_agreement"] >= 0.999


def test_8k_fine_steps_tile_rows_match_oracle(G):
    """cfg 4: 7680x4320, nstep 200.  The oracle renders bands of rows (incl. the hole's silhouette);
    the GPU renders the whole frame."""
    g = O.load_golden("cfg1_1920x1080")
    snap = g["snap"].with_resolution(7680, 4320)
    got = G.gpu_render(snap, nstep=200)
    assert got["stats"].rays == 7680 * 4320
    for rows in ((300, 316), (2900, 2916), (2940, 2948), (4200, 4216)):
        ref = O.render(snap, nstep=200, rows=rows)
        sl = slice(*rows)
        rep = parity.assert_parity({k: got[k][sl] for k in ("bgr", "cls", "key", "steps")},
                                   {k: ref[k][sl] for k in ("bgr", "cls", "key", "steps")}, "8k rows %s" % (rows,))
        print(rows, rep)


@pytest.mark.parametrize("name", ["cfg1_odd_333x187", "cfg2_640x360"])
def test_step_counts_other_than_the_references(G, name):
    """nstep 3, 7, 33 (UPV 2) and 64, 100 (UPV 3): leg changes on adjacent steps, the fused turn, both
    instantiations of the stepping loop -- against the oracle on the host cores."""
    snap = O.load_golden(name)["snap"]
    for n in (3, 7, 33, 64, 100):
        ref = O.render(snap, nstep=n)
        got = G.gpu_render(snap, nstep=n)
        rep = parity.assert_parity(got, ref, "%s nstep %d" % (name, n))
        assert rep["steps_agreement"] >= 0.999, (n, rep)
        assert int(got["steps"].max()) <= 2 * n - 1
        plain = G.gpu_render(snap, nstep=n, stats=False)
        for k in ("bgr", "cls", "key", "steps"):
            assert np.array_equal(plain[k], got[k]), (n, k)


def test_pixel_formats_agree(G):
    g = O.load_golden("cfg1_odd_333x187")  # odd width: scalar store path
    a = G.gpu_render(g["snap"], pixel_format=abi.PIXEL_BGR8)
    b = G.gpu_render(g["snap"], pixel_format=abi.PIXEL_BGRA8)
    c = G.gpu_render(g["snap"], pixel_format=abi.PIXEL_RGBA8)
    assert np.array_equal(a["bgr"], b["bgr"]) and np.array_equal(a["bgr"], c["bgr"])
    assert (b["pixels"][..., 3] == 255).all() and (c["pixels"][..., 3] == 255).all()
    g2 = O.load_golden("cfg2_640x360")  # width % 4 == 0: 16-byte vector store path
    a = G.gpu_render(g2["snap"], pixel_format=abi.PIXEL_BGR8)
    c = G.gpu_render(g2["snap"], pixel_format=abi.PIXEL_RGBA8)
    assert np.array_equal(a["bgr"], c["bgr"])


def test_batching_does_not_change_any_pixel(G):
    for name in ("cfg1_640x360", "cfg2_640x360"):
        g = O.load_golden(name)
        a = G.gpu_render(g["snap"])
        b = G.gpu_render(g["snap"], flags=abi.FLAG_NO_BATCHING)
        for k in ("bgr", "cls", "key", "steps"):
            assert np.array_equal(a[k], b[k]), (name, k)


def test_deterministic_and_batch_equals_single(G):
    r = G.renderer()
    snaps = [O.load_golden(n)["snap"] for n in ("cfg3_frame60_480x270", "cfg3_frame180_480x270")]
    r.set_textures(snaps[0], O.load_texture)
    batch = r.render(snaps + snaps, pixel_format=abi.PIXEL_RGBA8, want_maps=True)
    for i, s in enumerate(snaps + snaps):
        one = r.render(s, pixel_format=abi.PIXEL_RGBA8, want_maps=True)
        assert np.array_equal(batch["pixels"][i], one["pixels"][0])
        assert np.array_equal(batch["steps"][i], one["steps"][0])


def test_device_path_stripes_tile_the_frame(G):
    """Row-stripe shards rendered one after another into one device buffer == whole-frame render
    (this is what each rank does into GPU 0's peer-mapped buffer)."""
    r = G.renderer()
    g = O.load_golden("cfg1_640x360")
    snap = g["snap"]
    r.set_textures(snap, O.load_texture)
    whole = r.render(snap, pixel_format=abi.PIXEL_RGBA8)["pixels"][0]
    h, w = snap.height, snap.width
    d = r.frame_alloc(h * w * 4)
    try:
        for shards, stripe in ((2, 16), (3, 8), (8, 24)):
            r.memset_d(d, 0x5A, h * w * 4)
            for i in range(shards):
                r.render_device(snap, d, pixel_format=abi.PIXEL_RGBA8, stripe_rows=stripe, shard_index=i,
                                shard_count=shards)
            r.sync()
            host = np.empty((h, w, 4), np.uint8)
            r.memcpy_d2h(host, d)
            assert np.array_equal(host, whole), (shards, stripe)
        # a single shard leaves the other stripes untouched
        r.memset_d(d, 0x5A, h * w * 4)
        r.render_device(snap, d, pixel_format=abi.PIXEL_RGBA8, stripe_rows=16, shard_index=1, shard_count=2)
        r.sync()
        host = np.empty((h, w, 4), np.uint8)
        r.memcpy_d2h(host, d)
        assert (host[0:16] == 0x5A).all() and np.array_equal(host[16:32], whole[16:32])
    finally:
        r.frame_free(d)


def test_errors_are_loud(G):
    from blackhole_8_b200.renderer import Bh8Error
    r = G.renderer()
    g = O.load_golden("cfg0_frame7_320x180")
    d = g["snap"].to_dict()
    d["objects"][1]["kind"] = 7  # Triangle/Sphere/...: not on the GPU path, and never a silent fallback
    with pytest.raises(Bh8Error) as e:
        r.render(abi.SceneSnapshot.from_dict(d))
    assert e.value.code == abi.EUNSUPPORTED
    d = g["snap"].to_dict()
    d["objects"][1]["tex_id"] = 9  # texture slot never set
    with pytest.raises(Bh8Error) as e:
        r.render(abi.SceneSnapshot.from_dict(d))
    assert e.value.code == abi.EINVAL
    d = g["snap"].to_dict()
    d["bh_index"] = 1
    with pytest.raises(Bh8Error):
        r.render(abi.SceneSnapshot.from_dict(d))


def test_streaming_submit_wait_equals_render(G):
    r = G.renderer()
    snaps = [O.load_golden(n)["snap"] for n in ("cfg3_frame60_480x270", "cfg3_frame180_480x270")]
    r.set_textures(snaps[0], O.load_texture)
    bufs = [r.pinned((270, 480, 4)) for _ in range(2)]
    want = [r.render(s, pixel_format=abi.PIXEL_RGBA8)["pixels"][0].copy() for s in snaps]
    prev = None
    for k in range(7):  # more submissions than buffers: slots are recycled
        t = r.submit(snaps[k & 1], bufs[k & 1].array)
        if prev is not None:
            r.wait(prev[0])
            assert np.array_equal(bufs[prev[1]].array, want[prev[1]])
        prev = (t, k & 1)
    r.wait(prev[0])
    assert np.array_equal(bufs[prev[1]].array, want[prev[1]])
    for b in bufs:
        b.free()


def test_single_frame_render_in_row_bands_equals_whole_frame(monkeypatch):
    """bh8_render of ONE frame draws it as row bands (band k is read back while band k+1 is drawn, bands on two
    streams); pixels, maps and counters must equal the unbanded call's ($BH8_RENDER_BANDS=1), also when the
    height is not a multiple of the band height."""
    from blackhole_8_b200.renderer import Renderer
    results = {}
    for bands in ("1", "3", "4", "5"):
        monkeypatch.setenv("BH8_RENDER_BANDS", bands)
        r = Renderer((0,))
        try:
            for name in ("cfg1_640x360", "cfg2_640x360"):
                snap = O.load_golden(name)["snap"]
                r.set_textures(snap, O.load_texture)
                for fmt in (abi.PIXEL_BGR8, abi.PIXEL_RGBA8):
                    res = r.render(snap, pixel_format=fmt, want_maps=True, stats=True)
                    results[(bands, name, fmt)] = ({k: np.array(res[k]) for k in ("pixels", "cls", "key", "steps")},
                                                   (res["stats"].rays, res["stats"].steps))
        finally:
            r.close()
    for (bands, name, fmt), (maps, counts) in results.items():
        want_maps, want_counts = results[("1", name, fmt)]
        for k in maps:
            assert np.array_equal(maps[k], want_maps[k]), (bands, name, fmt, k)
        assert counts == want_counts, (bands, name, fmt)
