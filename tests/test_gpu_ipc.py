"""The torchrun data plane between TWO PROCESSES (bh8_ipc_export / bh8_ipc_import / bh8_ipc_close): one frame,
interleaved 16-row stripes, the second process stores its stripes straight into the first process's frame
buffer through a CUDA IPC mapping -- what `bench.py --gpus N` does between ranks (measure_stripes_8k).  Both
processes may share one GPU (CUDA IPC works between processes on the same device), so this runs on a one-GPU
box; with two GPUs the peer process takes the second one and the stores cross NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from blackhole_8_b200 import abi

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRIPE = 16

PEER = r"""
import os, sys
sys.path.insert(0, os.path.join({root!r}, "tests"))
sys.path.insert(0, {root!r})
import oracle_lib as O
from blackhole_8_b200 import abi
from blackhole_8_b200.renderer import Renderer
name, dev, handle = sys.argv[1], int(sys.argv[2]), bytes.fromhex(sys.argv[3])
snap = O.load_golden(name)["snap"]
r = Renderer((dev,))
r.set_textures(snap, O.load_texture)
buf = r.ipc_import(handle)
r.render_device(snap, buf, pixel_format=abi.PIXEL_RGBA8, stripe_rows={stripe}, shard_index=1, shard_count=2)
r.sync()
r.ipc_close(buf)
r.close()
print("peer done")
"""


@pytest.mark.parametrize("name", ["cfg1_640x360", "cfg1_odd_333x187"])
def test_two_processes_stripe_one_frame_through_ipc(name):
    import torch
    from blackhole_8_b200.renderer import Renderer
    snap = O.load_golden(name)["snap"]
    W, H = snap.width, snap.height
    r = Renderer((0,))
    try:
        r.set_textures(snap, O.load_texture)
        want = r.render(snap, pixel_format=abi.PIXEL_RGBA8)["pixels"][0].copy()
        nbytes = W * H * 4
        buf = r.frame_alloc(nbytes)
        r.memset_d(buf, 0, nbytes)
        r.sync()
        handle = r.ipc_export(buf)
        peer_dev = 1 if torch.cuda.device_count() >= 2 else 0
        # this process's stripes (0, 2, 4, ...) while the peer starts up
        r.render_device(snap, buf, pixel_format=abi.PIXEL_RGBA8, stripe_rows=STRIPE, shard_index=0, shard_count=2)
        r.sync()
        half = np.empty((H, W, 4), np.uint8)
        r.memcpy_d2h(half, buf)
        rows = np.arange(H)
        mine = (rows // STRIPE) % 2 == 0
        assert np.array_equal(half[mine], want[mine])
        assert not half[~mine].any()  # the peer's stripes are still untouched
        out = subprocess.run([sys.executable, "-c", PEER.format(root=ROOT, stripe=STRIPE), name, str(peer_dev),
                              handle.hex()], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "peer done" in out.stdout, out.stderr[-2000:]
        got = np.empty((H, W, 4), np.uint8)
        r.memcpy_d2h(got, buf)
        assert np.array_equal(got, want)
        r.frame_free(buf)
    finally:
        r.close()
