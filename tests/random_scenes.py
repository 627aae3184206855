"""Seeded random scenes for the parity tests: what a user of the reference's API could build with
StaticBlackhole / Annulus / Rectangle / InfinitePlane(ChessPattern2D) and a Camera that has been moved
and rotated, expressed as the snapshot PODs both the oracle and the renderer consume.

The fixed BASELINE scenes put the hole at the origin, the disc through its centre and the camera on an
axis; these do not: the hole sits anywhere near the origin, discs may miss the hole's centre, planes
tilt, the camera looks from any direction (also from inside the disc's radius, also away from the
hole), masses and fields of view vary -- the conservative filters of the kernel (bh8_ray.cuh) must
stay conservative everywhere."""
import numpy as np

from blackhole_8_b200 import abi

TEXTURES = ["winter.jpg", "acc_disc.png", "karina.jpeg", "mooni.jpeg", "disc_4color.png"]


def _basis(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q[:, 0], q[:, 1], q[:, 2]


def _normalize(v):  # cv::normalize
    n = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return v * ((1.0 / n) if n else 0.0)


def _obj(kind, key, tex_id=-1, pattern=0, v=None, n=(0, 0, 0), ex=(0, 0, 0), ey=(0, 0, 0), r_in=0.0, r_out=0.0,
         mass=0.0, pattern_size=0.0):
    vv = np.zeros((5, 3)) if v is None else np.asarray(v, float)
    return {"kind": kind, "key": key, "tex_id": tex_id, "pattern": pattern, "v": [float(x) for x in vv.reshape(-1)],
            "n": [float(x) for x in n], "ex": [float(x) for x in ex], "ey": [float(x) for x in ey],
            "r_in": float(r_in), "r_out": float(r_out), "mass": float(mass), "pattern_size": float(pattern_size)}


def _quad(center, a, b, half_a, half_b):
    """Corner points p1..p4 in the reference's order (vector_object.h:94-105) around `center`."""
    return [center + a * half_a + b * half_b, center - a * half_a + b * half_b,
            center - a * half_a - b * half_b, center + a * half_a - b * half_b]


def random_snapshot(seed, width=96, height=54):
    rng = np.random.default_rng(seed)
    # The reference's impact parameter is b = |F x d| / |F - d| with the UN-normalised pixel vector d
    # (|d| ~ focus_len ~ width): the hole's apparent size is set by b_c = 5.2 M against the width, not by
    # the camera's distance.  Masses are therefore drawn relative to the width.
    mass = float(np.exp(rng.uniform(np.log(0.004), np.log(0.08))) * width)
    bh = rng.uniform(-40.0, 40.0, 3) if rng.random() < 0.7 else np.zeros(3)
    objs, key = [], 0
    objs.append(_obj(abi.KIND_BLACKHOLE, key, v=np.vstack([bh] + [np.zeros(3)] * 4), mass=mass))
    key += 1
    textures = []

    def tex(name):
        if name not in textures:
            textures.append(name)
        return textures.index(name)

    # accretion disc(s): through the hole's centre (the usual case) or offset from it
    for _ in range(int(rng.integers(0, 3))):
        a, b, _n = _basis(rng)
        r_out = float(rng.uniform(15.0, 60.0) * mass)
        r_in = float(rng.uniform(2.5, 8.0) * mass)
        center = bh.copy() if rng.random() < 0.6 else bh + rng.normal(size=3) * rng.uniform(0.0, 6.0) * mass
        p = _quad(center, a, b, r_out, r_out)
        norm = _normalize(np.cross(p[2] - p[0], p[1] - p[0]))  # vector_object.h:325
        objs.append(_obj(abi.KIND_ANNULUS, key, tex(TEXTURES[int(rng.integers(0, 5))]), v=np.vstack([center] + p), n=norm,
                         r_in=r_in, r_out=r_out))
        key += 1
    # textured rectangles anywhere within the rays' reach
    for _ in range(int(rng.integers(0, 4))):
        a, b, _n = _basis(rng)
        center = bh + _normalize(rng.normal(size=3)) * rng.uniform(10.0, 120.0) * mass
        p = _quad(center, a, b, float(rng.uniform(5.0, 80.0) * mass), float(rng.uniform(5.0, 80.0) * mass))
        objs.append(_obj(abi.KIND_RECTANGLE, key, tex(TEXTURES[int(rng.integers(0, 5))]), v=np.vstack([np.zeros(3)] + p)))
        key += 1
    # a chess floor
    if rng.random() < 0.4:
        ex, ey, n = _basis(rng)
        pos = bh + n * rng.uniform(-40.0, 40.0) * mass * (rng.random() < 0.8)
        objs.append(_obj(abi.KIND_INFINITE_PLANE, key, pattern=abi.PATTERN_CHESS, v=np.vstack([pos] + [np.zeros(3)] * 4),
                         n=n, ex=ex, ey=ey, pattern_size=float(rng.uniform(2.0, 40.0))))
        key += 1
    order = rng.permutation(len(objs))  # unordered_map iteration order is arbitrary
    objs = [objs[i] for i in order]
    bh_index = [o["kind"] for o in objs].index(abi.KIND_BLACKHOLE)

    # camera: anywhere from just outside the photon sphere to far away, looking roughly at the hole or not
    dist = float(np.exp(rng.uniform(np.log(3.5), np.log(250.0))) * mass)
    pos = bh + _normalize(rng.normal(size=3)) * dist
    to_hole = _normalize(bh - pos)
    if rng.random() < 0.8:
        vx = _normalize(to_hole + rng.normal(size=3) * rng.uniform(0.0, 0.5))
    else:
        vx = _normalize(rng.normal(size=3))
    helper = _normalize(rng.normal(size=3))
    vy = _normalize(np.cross(helper, vx))
    vz = _normalize(np.cross(vx, vy))
    fov = float(rng.uniform(0.3, 2.2))
    cam = {"pos": [float(x) for x in pos], "vx": [float(x) for x in vx], "vy": [float(x) for x in vy],
           "vz": [float(x) for x in vz], "focus_len": float(width / (2.0 * np.tan(fov / 2.0))), "width": width,
           "height": height}
    nstep = int(rng.choice([20, 20, 20, 7, 50]))
    d = {"nstep": nstep, "linear_steps": 0, "width": width, "height": height, "camera": cam, "textures": textures,
         "bh_index": bh_index, "objects": objs}
    return abi.SceneSnapshot.from_dict(d)
