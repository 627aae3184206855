"""ctypes mirror of include/bh8.h (the C ABI of the B200 geodesic renderer).

Field order and types must stay in lock-step with the header; tests/test_abi.py checks the struct
sizes against the values the compiled library reports.
"""
import ctypes as C
import json

import numpy as np

ABI_VERSION = 5
MAX_OBJECTS = 16
MAX_TEXTURES = 16

KIND_BLACKHOLE, KIND_ANNULUS, KIND_RECTANGLE, KIND_INFINITE_PLANE = 0, 1, 2, 3
CLASS_BACKGROUND, CLASS_HORIZON, CLASS_DISC, CLASS_OBJECT = 0, 1, 2, 3
PATTERN_BLACK, PATTERN_CHESS = 0, 1
PIXEL_RGBA8, PIXEL_BGRA8, PIXEL_BGR8 = 0, 1, 2
FLAG_STATS, FLAG_NO_BATCHING = 1, 2
TRACER_GEODESIC, TRACER_LINEAR = 0, 1

EINVAL, EUNSUPPORTED, ECUDA, ENOMEM, ENODEVICE = -1, -2, -3, -4, -5

Vec3 = C.c_double * 3


class Camera(C.Structure):
    _fields_ = [
        ("pos", Vec3),
        ("vx", Vec3),
        ("vy", Vec3),
        ("vz", Vec3),
        ("focus_len", C.c_double),
        ("width", C.c_int32),
        ("height", C.c_int32),
    ]


class Object(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("key", C.c_int32),
        ("tex_id", C.c_int32),
        ("pattern", C.c_int32),
        ("v", Vec3 * 5),
        ("n", Vec3),
        ("ex", Vec3),
        ("ey", Vec3),
        ("r_in", C.c_double),
        ("r_out", C.c_double),
        ("mass", C.c_double),
        ("pattern_size", C.c_double),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("n_obj", C.c_int32),
        ("bh_index", C.c_int32),
        ("obj", C.POINTER(Object)),
    ]


class Params(C.Structure):
    _fields_ = [
        ("nstep", C.c_int32),
        ("pixel_format", C.c_int32),
        ("flags", C.c_uint32),
        ("stripe_rows", C.c_int32),
        ("shard_index", C.c_int32),
        ("shard_count", C.c_int32),
        ("tracer", C.c_int32),
        ("linear_steps", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("steps", C.c_uint64),
        ("class_count", C.c_uint64 * 4),
        ("tex_oob", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("total_ms", C.c_double),
        ("warps", C.c_uint64),
        ("update_slots", C.c_uint64),
        ("resolve_passes", C.c_uint64),
        ("exact_tests", C.c_uint64),
    ]

    def schedule(self):
        """The warp schedule in the units DESIGN.md 4.2 uses (geodesic kernel, BH8_FLAG_STATS launches)."""
        if not self.warps or not self.rays:
            return None
        return {"update_slots_per_warp": self.update_slots / self.warps,
                "updates_per_ray": self.steps / self.rays,
                "slot_utilisation": self.steps / (32.0 * self.update_slots) if self.update_slots else None,
                "resolve_passes_per_warp": self.resolve_passes / self.warps,
                "exact_tests_per_ray": self.exact_tests / self.rays}


OP_MOVE_X, OP_MOVE_Y, OP_MOVE_Z, OP_ROTATE_X, OP_ROTATE_Y, OP_ROTATE_Z, OP_MOVE_TO = range(7)
TARGET_CAMERA = -1


class Action(C.Structure):
    """bh8_action: one Object::Move* / Rotate* call of a scripted animation."""
    _fields_ = [
        ("frame", C.c_int32),
        ("target", C.c_int32),
        ("op", C.c_int32),
        ("reserved", C.c_int32),
        ("amount", C.c_double),
        ("to", Vec3),
    ]


HUD_MAX_LINES, HUD_MAX_TEXT = 16, 128


class HudLine(C.Structure):
    """bh8_hud_line: one line of HUD text, cv::putText(FONT_HERSHEY_PLAIN, 1, colour, 1) semantics."""
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("b", C.c_uint8), ("g", C.c_uint8), ("r", C.c_uint8),
                ("reserved", C.c_uint8), ("text", C.c_char * HUD_MAX_TEXT)]


def hud_lines(lines, color=(0, 255, 0)):
    """[(text, x, y) ...] -> array of HudLine."""
    arr = (HudLine * max(1, len(lines)))()
    for k, (text, x, y) in enumerate(lines):
        arr[k] = HudLine(int(x), int(y), color[0], color[1], color[2], 0, text.encode("latin-1", "replace")[:HUD_MAX_TEXT - 1])
    return arr


def reference_hud(camera, fov_deg=90):
    """The five lines blackhole_solution_test.cc:309-326 draws, for an abi.Camera (operator<< of cv::Vec3d)."""
    def vec(v):
        return "[" + ", ".join("%g" % x for x in v) + "]"
    return [("Position: " + vec(camera.pos), 0, 10), ("VectorX: " + vec(camera.vx), 0, 25),
            ("VectorY: " + vec(camera.vy), 0, 40), ("VectorZ: " + vec(camera.vz), 0, 55),
            ("FoV: %g deg" % fov_deg, 0, 70)]


class Basis(C.Structure):
    _fields_ = [("vx", Vec3), ("vy", Vec3), ("vz", Vec3)]


def reference_script(which, n_frames, disc_index):
    """The reference's frame loop as a list of Actions (blackhole_solution_test.cc:346-407):
      "cfg1_spin"        fixed camera; the disc spins RotateZ(pi/180) after every frame (:407)
      "cfg3_flythrough"  BASELINE configs[3]: 'w' MoveX(+10) after frames 0-119, then 'L'
                         RotateZ(+pi/1800*10) and 'd' MoveY(+10); plus the disc spin.
    Must stay in step with oracle/ref_render.cc's replay (tools/make_flythrough.py made the golden states)."""
    pi = 3.14159265358979323846
    acts = []
    for k in range(n_frames - 1):
        if which == "cfg3_flythrough":
            if k < 120:
                acts.append(Action(k, TARGET_CAMERA, OP_MOVE_X, 0, 10.0))
            else:
                acts.append(Action(k, TARGET_CAMERA, OP_ROTATE_Z, 0, pi / 1800 * 10))
                acts.append(Action(k, TARGET_CAMERA, OP_MOVE_Y, 0, 10.0))
        acts.append(Action(k, disc_index, OP_ROTATE_Z, 0, pi / 180))
    return acts


def pixel_bytes(pixel_format):
    return 3 if pixel_format == PIXEL_BGR8 else 4


class SceneSnapshot:
    """A scene + camera snapshot held as ctypes PODs (what bh8_render consumes).

    Built from the JSON that oracle/_ref/ref_render (reference classes) or the repo's own C++
    snapshot helper writes; `textures` lists the reference resource names by texture slot.
    """

    def __init__(self, camera, objects, bh_index, textures=(), nstep=20, meta=None, linear_steps=0):
        self.camera = camera
        self.objects = (Object * len(objects))(*objects)
        self.scene = Scene(len(objects), bh_index, C.cast(self.objects, C.POINTER(Object)))
        self.textures = list(textures)
        self.nstep = nstep
        self.linear_steps = linear_steps  # > 0: flat-space scene (RayTracer::Prograde), no black hole needed
        self.meta = meta or {}

    def params(self, pixel_format=PIXEL_RGBA8, flags=0, nstep=None, stripe_rows=0, shard_index=0, shard_count=0):
        """bh8_params for this snapshot: geodesic tracer, or the linear one for a flat-space scene."""
        return Params(nstep or self.nstep, pixel_format, flags, stripe_rows, shard_index, shard_count,
                      TRACER_LINEAR if self.linear_steps > 0 else TRACER_GEODESIC, self.linear_steps)

    @property
    def width(self):
        return self.camera.width

    @property
    def height(self):
        return self.camera.height

    @classmethod
    def from_dict(cls, d):
        c = d["camera"]
        cam = Camera(Vec3(*c["pos"]), Vec3(*c["vx"]), Vec3(*c["vy"]), Vec3(*c["vz"]),
                     c["focus_len"], c["width"], c["height"])
        objs = []
        for o in d["objects"]:
            v = o["v"]
            rows = (Vec3 * 5)(*[Vec3(*v[3 * i:3 * i + 3]) for i in range(5)])
            objs.append(Object(o["kind"], o["key"], o["tex_id"], o["pattern"], rows, Vec3(*o["n"]),
                               Vec3(*o["ex"]), Vec3(*o["ey"]), o["r_in"], o["r_out"], o["mass"],
                               o["pattern_size"]))
        meta = {k: d[k] for k in ("cfg", "frame", "run") if k in d}
        return cls(cam, objs, d["bh_index"], d.get("textures", ()), d.get("nstep", 20), meta,
                   d.get("linear_steps", 0))

    @classmethod
    def from_json(cls, path):
        with open(path) as f:
            return cls.from_dict(json.load(f))

    def to_dict(self):
        c = self.camera
        objs = []
        for o in self.objects:
            objs.append({
                "kind": o.kind, "key": o.key, "tex_id": o.tex_id, "pattern": o.pattern,
                "v": [x for row in o.v for x in row], "n": list(o.n), "ex": list(o.ex),
                "ey": list(o.ey), "r_in": o.r_in, "r_out": o.r_out, "mass": o.mass,
                "pattern_size": o.pattern_size,
            })
        return {
            "nstep": self.nstep, "linear_steps": self.linear_steps, "width": c.width, "height": c.height,
            "camera": {"pos": list(c.pos), "vx": list(c.vx), "vy": list(c.vy), "vz": list(c.vz),
                       "focus_len": c.focus_len, "width": c.width, "height": c.height},
            "textures": list(self.textures), "bh_index": self.scene.bh_index, "objects": objs,
        }

    def with_resolution(self, width, height, fov=None):
        """Same scene seen by Camera(width, height, fov): focus_len = width/(2 tan(fov/2)),
        camera.h:37-41.  The reference picture depends on the resolution, so this is a different
        workload, not a resampling."""
        d = self.to_dict()
        old = d["camera"]
        if fov is None:
            fov = 2.0 * np.arctan(old["width"] / (2.0 * old["focus_len"]))
        d["camera"]["focus_len"] = float(width / (2.0 * np.tan(fov / 2.0)))
        d["camera"]["width"] = d["width"] = int(width)
        d["camera"]["height"] = d["height"] = int(height)
        out = SceneSnapshot.from_dict(d)
        out.meta = dict(self.meta)
        return out
