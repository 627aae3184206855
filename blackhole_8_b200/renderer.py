"""Host-side mirror of the render call: Python over the C ABI of include/bh8.h (ctypes).

`Renderer.render(snapshot)` is the batched replacement of the reference's inline pixel loop
(include/blackhole/blackhole_solution_test.cc:161-308).  All pixels come from the CUDA kernels in
libbh8.so; if the library or a CUDA device is missing this module raises -- there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .build import LIB, is_stale

_lib = None


class Bh8Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bh8 error %d: %s" % (code, msg))
        self.code = code


def load_library(path=None):
    """Load libbh8.so and declare the signatures of include/bh8.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("BH8_LIB_PATH") or LIB
    if not os.path.exists(path):
        raise RuntimeError(
            "%s is missing: build it with `python -m blackhole_8_b200.build` (nvcc, sm_100a). "
            "The renderer has no CPU fallback." % path)
    L = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    sig = {
        "bh8_abi_version": (i32, []),
        "bh8_pixel_bytes": (C.c_size_t, [i32]),
        "bh8_launch_param_bytes": (C.c_size_t, []),
        "bh8_create": (i32, [C.POINTER(vp), C.POINTER(C.c_int), i32]),
        "bh8_destroy": (None, [vp]),
        "bh8_last_error": (C.c_char_p, [vp]),
        "bh8_launch_count": (u64, [vp]),
        "bh8_set_texture": (i32, [vp, i32, vp, i32, i32, C.c_size_t]),
        "bh8_render": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), i32, C.POINTER(abi.Params),
                             vp, vp, vp, vp, C.POINTER(abi.Stats)]),
        "bh8_submit": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), C.POINTER(abi.Params), vp,
                             C.POINTER(u64)]),
        "bh8_wait": (i32, [vp, u64]),
        "bh8_render_device": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), C.POINTER(abi.Params),
                                    vp, vp, vp, vp]),
        "bh8_sync": (i32, [vp]),
        "bh8_timer_begin": (i32, [vp]),
        "bh8_timer_end_ms": (i32, [vp, C.POINTER(C.c_double)]),
        "bh8_read_stats": (i32, [vp, C.POINTER(abi.Stats)]),
        "bh8_frame_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "bh8_frame_free": (i32, [vp, vp]),
        "bh8_ipc_export": (i32, [vp, vp, vp]),
        "bh8_ipc_import": (i32, [vp, vp, C.POINTER(vp)]),
        "bh8_ipc_close": (i32, [vp, vp]),
        "bh8_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "bh8_memset_d": (i32, [vp, vp, i32, C.c_size_t]),
        "bh8_host_alloc": (i32, [C.POINTER(vp), C.c_size_t]),
        "bh8_host_alloc_flags": (i32, [C.POINTER(vp), C.c_size_t, C.c_uint]),
        "bh8_measure_d2h": (i32, [vp, vp, vp, C.c_size_t, i32, C.POINTER(C.c_double)]),
        "bh8_host_free": (i32, [vp]),
        "bh8_host_register": (i32, [vp, C.c_size_t]),
        "bh8_host_unregister": (i32, [vp]),
        "bh8_measure_fp64_peak": (i32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "bh8_measure_stepping": (i32, [vp, i32, C.POINTER(C.c_double)]),
        "bh8_script_create": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Basis), C.POINTER(abi.Camera),
                                    C.POINTER(abi.Params), C.POINTER(abi.Action), i32, i32, C.POINTER(vp)]),
        "bh8_script_frames": (i32, [vp]),
        "bh8_script_render": (i32, [vp, i32, vp, vp, vp, vp]),
        "bh8_script_render_range": (i32, [vp, i32, i32, vp, C.c_size_t]),
        "bh8_script_state": (i32, [vp, i32, C.POINTER(abi.Camera), C.POINTER(abi.Object)]),
        "bh8_script_frame_constants": (i32, [vp, i32, vp, C.c_size_t]),
        "bh8_host_frame_constants": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), C.POINTER(abi.Params),
                                           vp, C.c_size_t]),
        "bh8_frame_bytes": (C.c_size_t, []),
        "bh8_script_destroy": (None, [vp]),
        "bh8_draw_text": (i32, [vp, i32, i32, C.c_size_t, i32, i32, C.c_char_p, i32, i32, i32]),
        "bh8_sink_open": (i32, [vp, C.c_char_p, i32, i32, C.c_double, i32, C.POINTER(vp)]),
        "bh8_sink_render": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), C.POINTER(abi.Params)]),
        "bh8_sink_submit": (i32, [vp, C.POINTER(abi.Scene), C.POINTER(abi.Camera), C.POINTER(abi.Params)]),
        "bh8_sink_flush": (i32, [vp]),
        "bh8_sink_hud": (i32, [vp, C.POINTER(abi.HudLine), i32]),
        "bh8_hud_draw_device": (i32, [vp, vp, i32, i32, C.POINTER(abi.HudLine), i32]),
        "bh8_sink_merge": (i32, [C.POINTER(C.c_char_p), i32, C.c_char_p, C.POINTER(u64), C.POINTER(u64)]),
        "bh8_sink_write_device": (i32, [vp, vp]),
        "bh8_sink_append_jpeg": (i32, [vp, vp, C.c_size_t]),
        "bh8_sink_last_jpeg": (i32, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "bh8_sink_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double)]),
        "bh8_sink_last_error": (C.c_char_p, [vp]),
        "bh8_sink_close": (i32, [vp, C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.bh8_abi_version() != abi.ABI_VERSION:
        raise RuntimeError("libbh8.so ABI %d != python mirror %d" % (L.bh8_abi_version(), abi.ABI_VERSION))
    if path == LIB:
        _lib = L
    return L


class PinnedBuffer:
    """cudaHostAlloc'ed host memory viewed as a numpy array."""

    def __init__(self, lib, shape, dtype, write_combined=False):
        self._lib = lib
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = lib.bh8_host_alloc_flags(C.byref(p), max(1, self.nbytes), 1 if write_combined else 0)
        if rc != 0:
            raise Bh8Error(rc, "cudaHostAlloc failed")
        self.ptr = p.value
        buf = (C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.bh8_host_free(C.c_void_p(self.ptr))
            self.ptr = None


class Renderer:
    """One bh8 context (one or more CUDA devices of this process)."""

    def __init__(self, devices=(0,)):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = self.lib.bh8_create(C.byref(self._ctx), devs, len(devices))
        if rc != 0:
            raise Bh8Error(rc, (self.lib.bh8_last_error(None) or b"").decode())
        self.devices = tuple(devices)
        self._textures = {}

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise Bh8Error(rc, (self.lib.bh8_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            self.lib.bh8_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launches(self):
        return int(self.lib.bh8_launch_count(self._ctx))

    # -- scene inputs ----------------------------------------------------------------------------
    def set_texture(self, slot, bgr):
        """bgr: (rows, cols, 3) uint8, the decoded Material::texture_ (BGR)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        assert bgr.ndim == 3 and bgr.shape[2] == 3
        self._check(self.lib.bh8_set_texture(self._ctx, slot, bgr.ctypes.data_as(C.c_void_p), bgr.shape[0],
                                             bgr.shape[1], bgr.strides[0]))
        self._textures[slot] = bgr.shape[:2]

    def set_textures(self, snapshot, loader):
        for slot, name in enumerate(snapshot.textures):
            self.set_texture(slot, loader(name))

    # -- host-buffer path (what a user calls) ------------------------------------------------------
    def render(self, snapshots, nstep=None, pixel_format=abi.PIXEL_RGBA8, want_maps=False, flags=0,
               stats=False, out=None, stripe_rows=0):
        """Render one snapshot or a list of them (same size) -> dict(pixels[, cls, key, steps, stats]).

        pixels: (n, H, W, bpp) uint8 in host memory.  Pass `out` (a dict of pre-allocated, ideally
        pinned arrays) to avoid allocations in a timed loop."""
        single = isinstance(snapshots, abi.SceneSnapshot)
        snaps = [snapshots] if single else list(snapshots)
        n = len(snaps)
        h, w = snaps[0].height, snaps[0].width
        bpp = abi.pixel_bytes(pixel_format)
        scenes = (abi.Scene * n)(*[s.scene for s in snaps])
        cams = (abi.Camera * n)(*[s.camera for s in snaps])
        prm = snaps[0].params(pixel_format, flags, nstep, stripe_rows)
        res = out if out is not None else {}
        if "pixels" not in res:
            res["pixels"] = np.empty((n, h, w, bpp), np.uint8)
        if want_maps:
            res.setdefault("cls", np.empty((n, h, w), np.uint8))
            res.setdefault("key", np.empty((n, h, w), np.int8))
            res.setdefault("steps", np.empty((n, h, w), np.uint16))
        st = abi.Stats() if stats else None

        def ptr(name):
            a = res.get(name)
            return a.ctypes.data_as(C.c_void_p) if a is not None else None

        self._check(self.lib.bh8_render(self._ctx, scenes, cams, n, C.byref(prm), ptr("pixels"), ptr("cls"),
                                        ptr("key"), ptr("steps"), C.byref(st) if stats else None))
        if stats:
            res["stats"] = st
        return res

    # -- streaming host-buffer path ------------------------------------------------------------------
    def submit(self, snap, out_array, nstep=None, pixel_format=abi.PIXEL_RGBA8, flags=0):
        """Launch one frame and queue its read-back into out_array (pinned host memory); returns a
        ticket for wait().  Two frames per device may be in flight."""
        prm = snap.params(pixel_format, flags, nstep)
        t = C.c_uint64()
        self._check(self.lib.bh8_submit(self._ctx, C.byref(snap.scene), C.byref(snap.camera), C.byref(prm),
                                        out_array.ctypes.data_as(C.c_void_p), C.byref(t)))
        return t.value

    def wait(self, ticket):
        self._check(self.lib.bh8_wait(self._ctx, C.c_uint64(ticket)))

    # -- device-resident path --------------------------------------------------------------------------
    def frame_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.bh8_frame_alloc(self._ctx, nbytes, C.byref(p)))
        return p.value

    def frame_free(self, ptr):
        self._check(self.lib.bh8_frame_free(self._ctx, C.c_void_p(ptr)))

    def render_device(self, snap, d_pixels, d_cls=None, d_key=None, d_steps=None, nstep=None,
                      pixel_format=abi.PIXEL_RGBA8, flags=0, stripe_rows=0, shard_index=0, shard_count=0):
        prm = snap.params(pixel_format, flags, nstep, stripe_rows, shard_index, shard_count)
        self._check(self.lib.bh8_render_device(self._ctx, C.byref(snap.scene), C.byref(snap.camera), C.byref(prm),
                                               C.c_void_p(d_pixels), C.c_void_p(d_cls), C.c_void_p(d_key),
                                               C.c_void_p(d_steps)))

    def hud_draw_device(self, d_bgr_frame, width, height, lines, color=(0, 255, 0)):
        """cv::putText semantics on a BGR8 frame in device memory (bh8_hud_draw_device)."""
        arr = abi.hud_lines(lines, color)
        self._check(self.lib.bh8_hud_draw_device(self._ctx, C.c_void_p(d_bgr_frame), width, height, arr, len(lines)))

    def sync(self):
        self._check(self.lib.bh8_sync(self._ctx))

    def timer_begin(self):
        self._check(self.lib.bh8_timer_begin(self._ctx))

    def timer_end_ms(self):
        ms = C.c_double()
        self._check(self.lib.bh8_timer_end_ms(self._ctx, C.byref(ms)))
        return ms.value

    def read_stats(self):
        st = abi.Stats()
        self._check(self.lib.bh8_read_stats(self._ctx, C.byref(st)))
        return st

    def memcpy_d2h(self, host_array, d_ptr):
        self._check(self.lib.bh8_memcpy_d2h(self._ctx, host_array.ctypes.data_as(C.c_void_p), C.c_void_p(d_ptr),
                                            host_array.nbytes))

    def memset_d(self, d_ptr, value, nbytes):
        self._check(self.lib.bh8_memset_d(self._ctx, C.c_void_p(d_ptr), value, nbytes))

    def ipc_export(self, d_ptr):
        h = (C.c_uint8 * 64)()
        self._check(self.lib.bh8_ipc_export(self._ctx, C.c_void_p(d_ptr), h))
        return bytes(h)

    def ipc_import(self, handle):
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self.lib.bh8_ipc_import(self._ctx, h, C.byref(p)))
        return p.value

    def ipc_close(self, d_ptr):
        self._check(self.lib.bh8_ipc_close(self._ctx, C.c_void_p(d_ptr)))

    def pinned(self, shape, dtype=np.uint8, write_combined=False):
        return PinnedBuffer(self.lib, shape, dtype, write_combined)

    def host_register(self, array):
        """Pin a numpy array the caller owns (bh8_host_register): read-backs into it become plain DMA.  Costs
        ~19 ms; call host_unregister(array) before the array is freed."""
        self._check(self.lib.bh8_host_register(array.ctypes.data_as(C.c_void_p), array.nbytes))

    def host_unregister(self, array):
        self._check(self.lib.bh8_host_unregister(array.ctypes.data_as(C.c_void_p)))

    def measure_d2h(self, buf_a, buf_b, nbytes, reps):
        """Seconds for `reps` read-backs of nbytes each into the two pinned buffers, no kernels (bh8_measure_d2h)."""
        s = C.c_double()
        self._check(self.lib.bh8_measure_d2h(self._ctx, C.c_void_p(buf_a.ptr), C.c_void_p(buf_b.ptr), nbytes, reps,
                                             C.byref(s)))
        return s.value

    def host_frame_constants(self, snap, nstep=None, pixel_format=abi.PIXEL_RGBA8, flags=0):
        """The frame constants the host derives from a snapshot (what every non-scripted launch passes)."""
        prm = snap.params(pixel_format, flags, nstep)
        buf = (C.c_uint8 * self.lib.bh8_frame_bytes())()
        self._check(self.lib.bh8_host_frame_constants(self._ctx, C.byref(snap.scene), C.byref(snap.camera),
                                                      C.byref(prm), buf, len(buf)))
        return bytes(buf)

    def measure_stepping(self, fp32):
        v = C.c_double()
        self._check(self.lib.bh8_measure_stepping(self._ctx, 1 if fp32 else 0, C.byref(v)))
        return v.value

    def measure_fp64_peak(self):
        f, s = C.c_double(), C.c_double()
        self._check(self.lib.bh8_measure_fp64_peak(self._ctx, C.byref(f), C.byref(s)))
        return f.value, s.value


class Script:
    """A scripted animation replayed and rendered on the GPU (bh8_script_*): the reference's
    per-frame Camera::Move*/Rotate* and Annulus::RotateZ calls (blackhole_solution_test.cc:346-407) as a
    list of abi.Action.  After construction no per-frame data comes from the host."""

    def __init__(self, renderer, snap0, actions, n_frames, basis=None, nstep=None, pixel_format=abi.PIXEL_RGBA8,
                 flags=0, stripe_rows=0, shard_index=0, shard_count=0):
        self._r = renderer
        self.lib = renderer.lib
        self._h = C.c_void_p()
        self.n_obj = snap0.scene.n_obj
        self.snap0 = snap0
        prm = snap0.params(pixel_format, flags, nstep, stripe_rows, shard_index, shard_count)
        acts = (abi.Action * max(1, len(actions)))(*actions)
        bas = (abi.Basis * self.n_obj)(*basis) if basis is not None else None
        renderer._check(self.lib.bh8_script_create(renderer._ctx, C.byref(snap0.scene), bas, C.byref(snap0.camera),
                                                   C.byref(prm), acts, len(actions), n_frames, C.byref(self._h)))
        self.n_frames = n_frames

    def render(self, frame, d_pixels, d_cls=None, d_key=None, d_steps=None):
        self._r._check(self.lib.bh8_script_render(self._h, frame, C.c_void_p(d_pixels), C.c_void_p(d_cls),
                                                  C.c_void_p(d_key), C.c_void_p(d_steps)))

    def render_range(self, first, count, d_base, frame_stride_bytes):
        """Frames first .. first+count-1 into d_base + i * stride with ONE host launch (a CUDA graph)."""
        self._r._check(self.lib.bh8_script_render_range(self._h, first, count, C.c_void_p(d_base), frame_stride_bytes))

    def state(self, frame):
        """(abi.Camera, [abi.Object]) of frame `frame`, read back from the device."""
        cam = abi.Camera()
        objs = (abi.Object * self.n_obj)()
        self._r._check(self.lib.bh8_script_state(self._h, frame, C.byref(cam), objs))
        return cam, list(objs)

    def frame_constants(self, frame):
        buf = (C.c_uint8 * self.lib.bh8_frame_bytes())()
        self._r._check(self.lib.bh8_script_frame_constants(self._h, frame, buf, len(buf)))
        return bytes(buf)

    def close(self):
        if self._h:
            self.lib.bh8_script_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class VideoSink:
    """Motion-JPEG AVI writer, the mirror of include/bh8.h's bh8_sink_* (SURVEY 8f-2): stands where the
    reference's cv::VideoWriter(path, fourcc('M','J','P','G'), fps, size) stands
    (blackhole_solution_test.cc:71-72,334).  With a Renderer the frames are encoded on the GPU
    (own JPEG kernels) from the device-resident frame; with renderer=None it is a plain container for ready
    JPEGs (no GPU needed)."""

    def __init__(self, renderer, path, width, height, fps=29.0, quality=95):
        self.lib = renderer.lib if renderer is not None else load_library()
        self._renderer = renderer
        self._h = C.c_void_p()
        ctx = renderer._ctx if renderer is not None else None
        rc = self.lib.bh8_sink_open(ctx, path.encode() if path else None, width, height, float(fps), int(quality),
                                    C.byref(self._h))
        if rc != 0:
            raise Bh8Error(rc, (self.lib.bh8_last_error(ctx) or b"").decode())
        self.width, self.height = width, height

    def _check(self, rc):
        if rc != 0:
            raise Bh8Error(rc, (self.lib.bh8_sink_last_error(self._h) or b"").decode())

    def render(self, snap, nstep=None, flags=0):
        """Trace one frame on the GPU and append it (nothing but the bitstream leaves the device)."""
        prm = snap.params(abi.PIXEL_BGR8, flags, nstep)
        self._check(self.lib.bh8_sink_render(self._h, C.byref(snap.scene), C.byref(snap.camera), C.byref(prm)))

    def submit(self, snap, nstep=None, flags=0):
        """Pipelined render(): returns once the frame's kernel and encode are queued; frame k+1 is
        traced while frame k is encoded.  flush() / close() append what is still in flight."""
        prm = snap.params(abi.PIXEL_BGR8, flags, nstep)
        self._check(self.lib.bh8_sink_submit(self._h, C.byref(snap.scene), C.byref(snap.camera), C.byref(prm)))

    def flush(self):
        self._check(self.lib.bh8_sink_flush(self._h))

    def hud(self, lines, color=(0, 255, 0)):
        """Text drawn on the device into every frame rendered from now on, before it is encoded: the
        reference's cv::putText block (blackhole_solution_test.cc:309-326).  lines: [(text, x, y), ...]; [] = off."""
        arr = abi.hud_lines(lines, color)
        self._check(self.lib.bh8_sink_hud(self._h, arr, len(lines)))

    def write_device(self, d_bgr_frame):
        self._check(self.lib.bh8_sink_write_device(self._h, C.c_void_p(d_bgr_frame)))

    def append_jpeg(self, data):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        self._check(self.lib.bh8_sink_append_jpeg(self._h, buf, len(data)))

    def last_jpeg(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.bh8_sink_last_jpeg(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value) if n.value else b""

    def stats(self):
        f, b, ms = C.c_uint64(), C.c_uint64(), C.c_double()
        self._check(self.lib.bh8_sink_stats(self._h, C.byref(f), C.byref(b), C.byref(ms)))
        return {"frames": f.value, "jpeg_bytes": b.value, "encode_ms": ms.value}

    def close(self):
        """Finish the file; returns its size in bytes (0 for a sink without a file)."""
        if not self._h:
            return 0
        n = C.c_uint64()
        rc = self.lib.bh8_sink_close(self._h, C.byref(n))
        self._h = C.c_void_p()
        if rc != 0:
            raise Bh8Error(rc, "bh8_sink_close failed (write error on the AVI file)")
        return n.value

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def merge_video_parts(part_paths, out_path):
    """One MJPG AVI from the per-GPU files of a frame-sharded job (bh8_sink_merge): frame k of the result
    is frame k // n of part k % n.  Host only.  Returns (frames, file_bytes)."""
    lib = load_library()
    arr = (C.c_char_p * len(part_paths))(*[p.encode() for p in part_paths])
    frames, size = C.c_uint64(), C.c_uint64()
    rc = lib.bh8_sink_merge(arr, len(part_paths), out_path.encode(), C.byref(frames), C.byref(size))
    if rc != 0:
        raise Bh8Error(rc, (lib.bh8_last_error(None) or b"").decode())
    return frames.value, size.value


def draw_text(frame_bgr, text, x, y, color=(0, 255, 0)):
    """cv::putText(frame, text, {x, y}, FONT_HERSHEY_PLAIN, 1, color, 1) -- the reference's HUD text
    (blackhole_solution_test.cc:313-325) -- on an (H, W, 3) uint8 BGR array, in place (bh8_draw_text)."""
    lib = load_library()
    assert frame_bgr.dtype == np.uint8 and frame_bgr.ndim == 3 and frame_bgr.shape[2] == 3
    rc = lib.bh8_draw_text(frame_bgr.ctypes.data_as(C.c_void_p), frame_bgr.shape[0], frame_bgr.shape[1],
                           frame_bgr.strides[0], int(x), int(y), text.encode("latin-1", "replace"),
                           int(color[0]), int(color[1]), int(color[2]))
    if rc != 0:
        raise Bh8Error(rc, "bh8_draw_text: bad arguments")
    return frame_bgr
