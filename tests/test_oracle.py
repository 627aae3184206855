"""The oracle (oracle/bh8_oracle.c, plain C) against the REFERENCE's own output.

tests/golden/* was produced by oracle/_ref/ref_render -- the reference's classes compiled from
/root/reference (tools/make_golden.py) -- and cfg0_960x540 is byte-identical to the frame the
UNCHANGED blackhole_solution_test binary writes (oracle/Makefile: check-ref).  The C port must
reproduce every fixture byte for byte: pixels, hit keys, classes and step counts.
"""
import math
import os

import numpy as np
import pytest

import oracle_lib as O

SMALL = [n for n in O.golden_names(full=False)]
FULL = [n for n in O.golden_names() if n not in SMALL]


@pytest.mark.parametrize("name", SMALL)
def test_port_reproduces_reference_frames(name):
    g = O.load_golden(name)
    r = O.render(g["snap"])
    assert O.digest(g["bgr"]) == g["digest"]["bgr"], "fixture PNG does not decode to the recorded frame"
    assert np.array_equal(r["cls"], g["cls"])
    assert np.array_equal(r["key"], g["key"])
    assert np.array_equal(r["steps"], g["steps"])
    assert np.array_equal(r["bgr"], g["bgr"])
    assert r["result"].steps == g["run"]["steps"]
    assert list(r["result"].class_count) == g["run"]["class_count"]
    assert r["result"].tex_oob == 0


@pytest.mark.parametrize("name", FULL)
def test_port_reproduces_reference_digests_full_size(name):
    g = O.load_golden(name)
    r = O.render(g["snap"])
    for k in ("bgr", "cls", "key", "steps"):
        assert O.digest(r[k]) == g["digest"][k], k
    assert r["result"].steps == g["run"]["steps"]


def test_threads_and_row_ranges_do_not_change_results():
    g = O.load_golden("cfg1_odd_333x187")
    a = O.render(g["snap"], threads=1)
    b = O.render(g["snap"], threads=5)
    assert np.array_equal(a["bgr"], b["bgr"]) and np.array_equal(a["steps"], b["steps"])
    top = O.render(g["snap"], rows=(0, 90))
    bot = O.render(g["snap"], rows=(90, 187))
    assert np.array_equal(top["bgr"][:90], a["bgr"][:90])
    assert np.array_equal(bot["bgr"][90:], a["bgr"][90:])
    assert not top["bgr"][90:].any()
    assert top["result"].rays + bot["result"].rays == a["result"].rays


def test_solve_g_known_answers():
    """StaticBlackhole::SolveG (blackhole_solution.h:35-53): 20 bisections from
    [cbrt(DBL_EPSILON), 1/(3M)], returns the LEFT end, so G(sol) > 0 and the true root lies
    within one grid step (1/(3M) - eps) / 2^20 to the right."""
    L = O.lib()
    M = 10.0
    eps = float(np.cbrt(np.finfo(np.float64).eps))
    h = (1.0 / (3 * M) - eps) / 2 ** 20
    for b in (3 * math.sqrt(3) * M, 52.0, 60.0, 100.0, 333.3, 1000.0, 2039.6):
        sol = L.bh8_oracle_solve_g(M, b)
        assert L.bh8_oracle_G(M, sol, b) > 0
        assert eps <= sol < 1 / (3 * M)
        if b > 52:
            assert L.bh8_oracle_G(M, sol + 1.0001 * h, b) <= 0
        k = (sol - eps) / h
        assert abs(k - round(k)) < 1e-6  # sol sits on the bisection grid
    # weak field: root -> 1/b
    assert abs(L.bh8_oracle_solve_g(M, 2000.0) * 2000.0 - 1) < 0.02


def test_reference_binary_agrees_when_present(tmp_path):
    """If the reference-built ref_render travelled with the repo, run it here and compare the C
    port against it at a size not in the fixtures (it never reads /root/reference at run time)."""
    ref = os.path.join(O.ROOT, "oracle", "_ref", "ref_render")
    tex = os.path.join(O.ROOT, "build", "textures")
    if not (os.path.exists(ref) and os.path.isdir(tex)):
        pytest.skip("oracle/_ref/ref_render or build/textures not present")
    import subprocess
    from blackhole_8_b200 import abi
    prefix = str(tmp_path / "r")
    subprocess.run([ref, "--cfg", "2", "--width", "411", "--height", "233", "--threads", "4",
                    "--texdir", tex, "--out", prefix], check=True)
    snap = abi.SceneSnapshot.from_json(prefix + ".json")
    r = O.render(snap)
    raw = np.fromfile(prefix + ".bgr", dtype=np.uint8)[8:].reshape(233, 411, 3)
    assert np.array_equal(r["bgr"], raw)
    assert np.array_equal(r["steps"], np.fromfile(prefix + ".steps", dtype=np.uint16).reshape(233, 411))
    assert np.array_equal(r["key"], np.fromfile(prefix + ".key", dtype=np.int8).reshape(233, 411))
