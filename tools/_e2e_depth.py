import sys, time, os
sys.path.insert(0, '.')
import numpy as np
import bench
from blackhole_8_b200 import abi
from blackhole_8_b200.renderer import Renderer
base = bench.load_snapshot(); H, W = base.height, base.width
seq = bench.frame_sequence("cfg1_spin", 240)
def run(n_ctx, n=3000):
    rs = [Renderer((0,)) for _ in range(n_ctx)]
    for r in rs: r.set_textures(base, bench.load_texture)
    pins = [[r.pinned((H, W, 3)) for _ in range(2)] for r in rs]
    def go(n):
        t0 = time.perf_counter(); pend = []
        for i in range(n):
            c = i % n_ctx; r = rs[c]
            tk = r.submit(seq[i % 240], pins[c][(i // n_ctx) & 1].array, pixel_format=abi.PIXEL_BGR8)
            pend.append((r, tk))
            if len(pend) > n_ctx: 
                rr, t = pend.pop(0); rr.wait(t)
        for rr, t in pend: rr.wait(t)
        return time.perf_counter() - t0
    go(50); dt = min(go(n) for _ in range(3))
    print("contexts %d (frames in flight %d): %.0f Mrays/s" % (n_ctx, n_ctx + 1 if n_ctx > 1 else 2, n * H * W / dt / 1e6))
    for r in rs: r.close()
run(1); run(2); run(3)
