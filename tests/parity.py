"""Parity metric of BASELINE.json's north_star, shared by the CPU harness test and the GPU tests.

  * hit class (background / horizon / disc / object) must agree on >= 99.9 % of pixels;
  * on pixels whose class agrees, every BGR channel must be within 2/255 -- a nearest-texel
    flip on a detailed texture can exceed that, so the share of such pixels is bounded too;
  * the residual is reported: how many mismatching pixels are near-critical rays (within 2 px of
    the horizon silhouette) or sit on an object edge.
"""
import numpy as np

CLASS_AGREEMENT_MIN = 0.999
RGB_TOL = 2
RGB_OUTLIER_MAX = 0.001


def compare(got, ref):
    """got/ref: dicts with bgr (H,W,3) u8, cls (H,W) u8, optional key, steps."""
    cls_ok = got["cls"] == ref["cls"]
    n = cls_ok.size
    diff = np.abs(got["bgr"].astype(np.int16) - ref["bgr"].astype(np.int16)).max(axis=2)
    rgb_bad = cls_ok & (diff > RGB_TOL)
    rep = {
        "pixels": int(n),
        "class_agreement": float(cls_ok.mean()),
        "class_mismatch": int((~cls_ok).sum()),
        "rgb_outliers": int(rgb_bad.sum()),
        "rgb_outlier_share": float(rgb_bad.sum() / n),
        "rgb_max_diff_matching": int(diff[cls_ok].max()) if cls_ok.any() else 0,
        "exact_pixels": float((cls_ok & (diff == 0)).mean()),
    }
    if "key" in got and "key" in ref:
        rep["key_agreement"] = float((got["key"] == ref["key"]).mean())
    if "steps" in got and "steps" in ref:
        rep["steps_agreement"] = float((got["steps"] == ref["steps"]).mean())
        rep["steps_total"] = (int(got["steps"].sum()), int(ref["steps"].sum()))
    # where is the residual?  near the horizon silhouette (near-critical rays) or elsewhere
    bad = ~cls_ok | rgb_bad
    if bad.any():
        hz = ref["cls"] == 1
        near = np.zeros_like(hz)
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                near |= np.roll(np.roll(hz, dy, 0), dx, 1)
        edge = near & ~(hz & np.roll(hz, 1, 0) & np.roll(hz, -1, 0) & np.roll(hz, 1, 1) & np.roll(hz, -1, 1))
        rep["residual_near_horizon_edge"] = int((bad & edge).sum())
        rep["residual_elsewhere"] = int((bad & ~edge).sum())
    return rep


def assert_parity(got, ref, what=""):
    rep = compare(got, ref)
    assert rep["class_agreement"] >= CLASS_AGREEMENT_MIN, (what, rep)
    assert rep["rgb_outlier_share"] <= RGB_OUTLIER_MAX, (what, rep)
    return rep
