"""The C-ABI library loads on a CPU-only box and exports every symbol include/bh8.h declares."""
import ctypes as C
import os
import re

import pytest

from blackhole_8_b200 import abi
from blackhole_8_b200.build import LIB, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    build()
    from blackhole_8_b200.renderer import load_library
    return load_library()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bh8.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bh8_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version_and_struct_sizes(lib):
    assert lib.bh8_abi_version() == abi.ABI_VERSION
    assert C.sizeof(abi.Camera) == 13 * 8 + 8
    assert C.sizeof(abi.Object) == 16 + (15 + 9 + 4) * 8
    assert C.sizeof(abi.Scene) == 16
    assert C.sizeof(abi.Params) == 32
    assert C.sizeof(abi.Stats) == 13 * 8
    assert lib.bh8_pixel_bytes(abi.PIXEL_BGR8) == 3 and lib.bh8_pixel_bytes(abi.PIXEL_RGBA8) == 4


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    rc = lib.bh8_create(C.byref(ctx), None, 1)
    assert rc == abi.ENODEVICE and not ctx.value
    assert b"no CPU fallback" in lib.bh8_last_error(None)


def test_product_package_does_not_touch_the_oracle():
    """Neither the package, nor the public headers, nor the driver apps may import, include, link or load
    anything under oracle/ (only tests/, smoke() and bench.py's CPU-baseline legs do)."""
    walk = []
    for top in ("blackhole_8_b200", "include", "apps"):
        walk += list(os.walk(os.path.join(ROOT, top)))
    for dirpath, _, files in walk:
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                bad = re.findall(r"import\s+\S*oracle|from\s+\S*oracle|libbh8_oracle|bh8_oracle_|"
                                 r"#include\s+\"[^\"]*oracle|CDLL\([^)]*oracle", text)
                assert not bad, (os.path.join(dirpath, f), bad)


def test_header_is_plain_c_and_struct_sizes_match_the_ctypes_mirror(tmp_path):
    """include/bh8.h must be consumable by a C compiler (the boundary is a C ABI, not a C++ one) and the
    ctypes mirror of every POD must have the size the C compiler gives it."""
    import subprocess
    src = tmp_path / "abi_sizes.c"
    src.write_text('#include <stdio.h>\n#include "bh8.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bh8_camera), sizeof(bh8_object), sizeof(bh8_scene),\n'
                   '         sizeof(bh8_params), sizeof(bh8_stats), sizeof(bh8_action), sizeof(bh8_basis));\n  return 0;\n}\n')
    exe = tmp_path / "abi_sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    mirror = [C.sizeof(t) for t in (abi.Camera, abi.Object, abi.Scene, abi.Params, abi.Stats, abi.Action, abi.Basis)]
    assert sizes == mirror
