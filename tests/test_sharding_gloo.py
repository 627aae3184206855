"""N > 1 path on CPU: two ranks (torch.distributed, gloo) deal frames / row stripes exactly as the
GPU ranks do, each renders its share with the C oracle standing in for its GPU, and rank 0's gathered
buffer must equal the single-process result -- disjoint cover, right offsets, nothing reduced."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from blackhole_8_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "stripes":
            g = O.load_golden("cfg1_odd_333x187")
            snap = g["snap"]
            h, w = snap.height, snap.width
            mine = torch.zeros((h, w, 3), dtype=torch.uint8)
            for r0, r1 in sharding.stripe_rows_of(h, 16, rank, world):
                part = O.render(snap, rows=(r0, r1), threads=2)
                mine[r0:r1] = torch.from_numpy(part["bgr"][r0:r1])
            # the GPU ranks store into GPU 0's buffer; here the stripes travel by gather (no reduction)
            bufs = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
            dist.gather(mine, bufs, dst=0)
            if rank == 0:
                frame = torch.zeros_like(mine)
                for r, b in enumerate(bufs):
                    for r0, r1 in sharding.stripe_rows_of(h, 16, r, world):
                        frame[r0:r1] = b[r0:r1]
                np.save(out_path, frame.numpy())
        elif mode == "video":
            # one sink per rank on its own frames (the GPU ranks encode on the device; here cv2 stands in),
            # then rank 0 merges the per-rank files into ONE video in frame order (bh8_sink_merge)
            import cv2
            from blackhole_8_b200.renderer import VideoSink, merge_video_parts
            base = np.ascontiguousarray(O.load_golden("cfg1_odd_333x187")["bgr"])
            h, w = base.shape[:2]
            n_frames = 7
            part = out_path + ".part%d.avi" % rank
            with VideoSink(None, part, w, h, fps=29) as sink:
                for k in sharding.frames_of(n_frames, rank, world):
                    ok, enc = cv2.imencode(".jpg", np.roll(base, 9 * k, axis=1), [cv2.IMWRITE_JPEG_QUALITY, 95])
                    sink.append_jpeg(enc.tobytes())
            dist.barrier()
            if rank == 0:
                frames, _ = merge_video_parts([out_path + ".part%d.avi" % r for r in range(world)], out_path)
                assert frames == n_frames
        else:
            names = ["cfg3_frame60_480x270", "cfg3_frame180_480x270", "cfg5_480x270"]
            frames = []
            for k in sharding.frames_of(len(names), rank, world):
                frames.append((k, O.digest(O.render(O.load_golden(names[k])["snap"], threads=2)["bgr"])))
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(frames, gathered, dst=0)
            if rank == 0:
                flat = dict(kv for part in gathered for kv in part)
                np.save(out_path, np.array([flat[k] for k in range(len(names))]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_row_stripes_of_two_ranks_tile_the_frame(tmp_path):
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), "stripes", out), nprocs=2, join=True)
    g = O.load_golden("cfg1_odd_333x187")
    assert np.array_equal(np.load(out), g["bgr"])


def test_frames_dealt_round_robin_cover_the_flythrough(tmp_path):
    out = str(tmp_path / "digests.npy")
    mp.spawn(_worker, args=(2, _free_port(), "frames", out), nprocs=2, join=True)
    names = ["cfg3_frame60_480x270", "cfg3_frame180_480x270", "cfg5_480x270"]
    assert list(np.load(out)) == [O.load_golden(n)["digest"]["bgr"] for n in names]


def test_per_rank_videos_merge_into_one_file_in_frame_order(tmp_path):
    import cv2
    out = str(tmp_path / "video.avi")
    mp.spawn(_worker, args=(2, _free_port(), "video", out), nprocs=2, join=True)
    base = np.ascontiguousarray(O.load_golden("cfg1_odd_333x187")["bgr"])
    cap = cv2.VideoCapture(out)
    k = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        want = np.roll(base, 9 * k, axis=1)
        mse = np.mean((frame.astype(float) - want.astype(float)) ** 2)
        assert 10 * np.log10(255.0 ** 2 / mse) > 30.0, "frame %d is not frame %d of the job" % (k, k)
        k += 1
    assert k == 7


def test_partitions_are_disjoint_covers():
    for world in (1, 2, 3, 4, 8):
        for h in (8, 187, 1080, 4320):
            for stripe in (8, 16, 64):
                rows = sorted(r for rank in range(world) for r in sharding.stripe_rows_of(h, stripe, rank, world))
                assert rows[0][0] == 0 and rows[-1][1] == h
                assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert sorted(k for r in range(world) for k in sharding.frames_of(240, r, world)) == list(range(240))
    with pytest.raises(ValueError):
        sharding.stripe_rows_of(100, 12, 0, 2)
    offs = {sharding.ring_slot_offset(s, r, 4, 3, 100) for s in range(3) for r in range(4)}
    assert len(offs) == 12 and max(offs) == 1100
