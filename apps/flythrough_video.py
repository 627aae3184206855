#!/usr/bin/env python
"""BASELINE configs[3] end to end on N GPUs: the 240-frame camera fly-through, frames sharded round-robin
over one process per GPU, every frame traced AND JPEG-encoded on its GPU (the frame sink), the per-GPU
Motion-JPEG files merged into ONE video.avi with the frames in the order the reference's single loop
writes them (blackhole_solution_test.cc:71-72,334,346-407).

  python apps/flythrough_video.py --out video.avi [--frames 240] [--width 1920 --height 1080]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
         apps/flythrough_video.py --out video.avi

The camera / disc animation is a device-side script (bh8_script_*) when --script is given (every rank
replays the whole script on its GPU and draws its own frames from it), else host snapshots
(tests/golden/states, made by the reference's classes).  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--quality", type=int, default=95)
    ap.add_argument("--script", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import bench
    from blackhole_8_b200 import abi, sharding
    from blackhole_8_b200.renderer import Renderer, Script, VideoSink, merge_video_parts
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # barriers only: nothing is exchanged but files on disk

    seq = [s.with_resolution(args.width, args.height) for s in bench.frame_sequence("cfg3_flythrough", args.frames)]
    r = Renderer((local_rank,))
    r.set_textures(seq[0], bench.load_texture)
    part = "%s.part%d.avi" % (args.out, rank)
    mine = sharding.frames_of(len(seq), rank, world)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    sink = VideoSink(r, part, args.width, args.height, fps=29, quality=args.quality)
    if args.script:
        disc = [o.kind for o in seq[0].objects].index(abi.KIND_ANNULUS)
        sc = Script(r, seq[0], abi.reference_script("cfg3_flythrough", len(seq), disc), len(seq),
                    pixel_format=abi.PIXEL_BGR8)
        d_frame = r.frame_alloc(args.width * args.height * 3)
        for k in mine:
            sc.render(k, d_frame)
            r.sync()
            sink.write_device(d_frame)
        sc.close()
    else:
        for k in mine:
            sink.submit(seq[k])
    sink.close()
    if dist:
        dist.barrier()
    render_s = time.perf_counter() - t0
    if rank == 0:
        parts = ["%s.part%d.avi" % (args.out, q) for q in range(world)]
        t1 = time.perf_counter()
        frames, size = merge_video_parts(parts, args.out)
        merge_s = time.perf_counter() - t1
        for p in parts:
            os.remove(p)
        line = {"frames": frames, "file_bytes": size, "n_gpus": world, "render_encode_s": render_s, "merge_s": merge_s,
                "frames_per_s": frames / render_s, "animation": "device script" if args.script else "host snapshots"}
        try:  # read the file back the way a consumer of the reference's video.avi would
            import cv2
            import numpy as np
            cap = cv2.VideoCapture(args.out)
            n, worst = 0, 99.0
            check = {0, frames // 2, frames - 1}
            while True:
                ok, fr = cap.read()
                if not ok:
                    break
                if n in check:
                    ref = r.render(seq[n], pixel_format=abi.PIXEL_BGR8)["pixels"][0]
                    mse = float(np.mean((fr.astype(np.float64) - ref.astype(np.float64)) ** 2))
                    worst = min(worst, 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse))
                n += 1
            line.update(frames_read_back=n, min_psnr_db_checked_frames=worst)
        except ImportError:
            pass
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
