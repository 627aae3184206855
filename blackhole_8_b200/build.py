"""Build libbh8.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbh8.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libbh8.so cannot be built (there is no CPU fallback)")
    return exe


def sources():
    return [os.path.join(SRC, n) for n in sorted(os.listdir(SRC))] + [os.path.join(ROOT, "include", "bh8.h")]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, out=None, defines=()):
    """Compile blackhole_8_b200/csrc/bh8_lib.cu -> blackhole_8_b200/libbh8.so (or `out`, with extra
    -D defines, for tuning variants selected at run time through $BH8_LIB_PATH)."""
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + \
        ["-I" + os.path.join(ROOT, "include"), "-I" + SRC, "-o", out, os.path.join(SRC, "bh8_lib.cu"), "-lnvjpeg"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
