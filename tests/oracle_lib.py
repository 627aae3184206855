"""ctypes driver for the ORACLE (oracle/libbh8_oracle.so).  Test infrastructure only.

Nothing under blackhole_8_b200/ may import this module.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess

import cv2
import numpy as np

from blackhole_8_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_LIB = None


class OracleTexture(C.Structure):
    _fields_ = [("bgr", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32)]


class OracleResult(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("steps", C.c_uint64), ("class_count", C.c_uint64 * 4),
                ("tex_oob", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "libbh8_oracle.so")
        src = os.path.join(ROOT, "oracle", "bh8_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True,
                           stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(so)
        _LIB.bh8_oracle_G.restype = C.c_double
        _LIB.bh8_oracle_G.argtypes = [C.c_double] * 3
        _LIB.bh8_oracle_solve_g.restype = C.c_double
        _LIB.bh8_oracle_solve_g.argtypes = [C.c_double] * 2
        _LIB.bh8_oracle_render.restype = C.c_int
        _LIB.bh8_oracle_render_linear.restype = C.c_int
    return _LIB


_TEX_CACHE = {}


def load_texture(name):
    """Decoded BGR pixels of a reference resource (resource/*.bgr.png, md5-checked)."""
    if name not in _TEX_CACHE:
        with open(os.path.join(ROOT, "resource", "MANIFEST.json")) as f:
            ent = json.load(f)[name]
        img = cv2.imread(os.path.join(ROOT, "resource", ent["file"]), cv2.IMREAD_COLOR)
        assert img is not None and hashlib.md5(img.tobytes()).hexdigest() == ent["md5_bgr"], name
        _TEX_CACHE[name] = np.ascontiguousarray(img)
    return _TEX_CACHE[name]


def render(snap, nstep=None, threads=None, rows=None):
    """Run the C oracle on a SceneSnapshot -> dict(bgr, cls, key, steps, result)."""
    L = lib()
    h, w = snap.height, snap.width
    texs = [load_texture(n) for n in snap.textures]
    tarr = (OracleTexture * max(1, len(texs)))()
    for i, t in enumerate(texs):
        tarr[i] = OracleTexture(t.ctypes.data, t.shape[0], t.shape[1])
    bgr = np.zeros((h, w, 3), np.uint8)
    cls = np.zeros((h, w), np.uint8)
    key = np.full((h, w), -1, np.int8)
    steps = np.zeros((h, w), np.uint16)
    res = OracleResult()
    r0, r1 = rows if rows else (0, h)
    if snap.linear_steps > 0:
        fn, count = L.bh8_oracle_render_linear, snap.linear_steps
    else:
        fn, count = L.bh8_oracle_render, nstep or snap.nstep
    rc = fn(C.byref(snap.scene), C.byref(snap.camera), C.c_int(count),
                             tarr, C.c_int(len(texs)), C.c_int(r0), C.c_int(r1),
                             C.c_int(threads or os.cpu_count() or 1),
                             bgr.ctypes.data_as(C.c_void_p), cls.ctypes.data_as(C.c_void_p),
                             key.ctypes.data_as(C.c_void_p), steps.ctypes.data_as(C.c_void_p),
                             C.byref(res))
    assert rc == 0, rc
    return {"bgr": bgr, "cls": cls, "key": key, "steps": steps, "result": res}


def golden_names(full=True):
    return sorted(n[:-5] for n in os.listdir(GOLDEN)
                  if n.endswith(".json") and (full or os.path.exists(os.path.join(GOLDEN, n[:-5] + ".npz"))))


def load_golden(name):
    snap = abi.SceneSnapshot.from_json(os.path.join(GOLDEN, name + ".json"))
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    out = {"snap": snap, "digest": meta["digest"], "run": meta["run"]}
    if os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        out.update(cls=z["cls"], key=z["key"], steps=z["steps"],
                   bgr=cv2.imread(os.path.join(GOLDEN, name + ".png"), cv2.IMREAD_COLOR))
    return out


def digest(arr):
    return hashlib.md5(np.ascontiguousarray(arr).tobytes()).hexdigest()
