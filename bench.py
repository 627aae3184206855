#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 geodesic renderer (driver contract).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU renderer, host cores

Workload (BASELINE.json configs[1]): single Schwarzschild black hole + thin accretion disc
(acc_disc.png) + textured background rectangle, 1920x1080, nstep 20 -- the scene snapshot the
reference's own classes produced (tests/golden/cfg1_1920x1080.json).  One "step" = one frame.
The frames are consecutive frames of the reference's own frame loop for that scene (the disc spins
pi/180 per frame, blackhole_solution_test.cc:407).  At N > 1 (one process per GPU, torchrun) step i
of the job renders frames i*N .. i*N+N-1 of the SAME sequence, rank r owning frame i*N+r -- weak
scaling, no collective on the data path; the frames are gathered on GPU 0 by the kernels' own stores
into its IPC-mapped frame ring (NVLink peer memory), and rank 0 checks the ring afterwards
(`gather_ok`).  Every N > 1 line also carries `stripes_8k`: BASELINE configs[4], one 7680x4320
nstep-200 frame split into interleaved row stripes over the ranks (strong scaling), checked
against a single-GPU render.

The JSON line: value = Mrays/s with everything resident on the GPU (CUDA events around each
frame's kernel, L2 flushed between frames outside the event pair); e2e = the same metric through
the public host-buffer calls (bh8_submit / bh8_wait: snapshot in, the reference's CV_8UC3 frame read
back into pinned host memory, wall clock, passes of at least 0.3 s) with the read-back ceiling of the
box measured beside it (`e2e.d2h_only`); roofline = FP64 flops the kernel executes / kernel time
against the DFMA peak measured in the same run (the algorithmic count of SURVEY.md 8(d) is kept
beside it); cpu_baseline = the reference's CPU renderer on this box's host cores, same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "cfg1_1920x1080"
WORKLOAD_TEXT = ("cfg1: 1 Schwarzschild BH + accretion disc (acc_disc.png) + background rectangle, "
                 "1920x1080, nstep 20, one frame per step")
# FP64 flops one geodesic update EXECUTES in the kernel (lane_advance, bh8_ray.cuh): u += delta (1);
# G = fma(u*u, fma(2M, u, -1), 1/b^2) (1 + 2 + 2); rsqrt correction x*y, fma, y*e, fma, fma (1 + 2 + 1 + 2 + 2);
# s = dphi_prev + dphi (1); phi = fma(s, du/2, phi) (2)
FLOPS_EXECUTED_PER_UPDATE = 17


def workload_config(sequence="cfg1_spin"):
    """`config` of BOTH arms (this repo's CUDA path and --impl reference): what is rendered, nothing about
    how -- the two dicts are equal, so the driver can tell that both arms ran the same job."""
    return {"workload": WORKLOAD_TEXT, "resolution": [1920, 1080], "nstep": 20, "rays_per_frame": 1920 * 1080,
            "frames_per_step_per_gpu": 1,
            "frame_sequence": "%s: consecutive frames of the reference's frame loop (the disc spins pi/180 per "
                              "frame, blackhole_solution_test.cc:407)" % sequence,
            "pixel_format_e2e": "BGR8 (the reference's CV_8UC3 frame)",
            "l2": "CUDA arm: flushed between steps by a 256 MiB memset outside the per-step CUDA-event pair; "
                  "reference arm: host caches, nothing to flush"}
FLOPS_PER_STEP_BASE = 57 + 8 + 10 + 36  # 2-object scene, rsqrt 8, rcp 10, sincos 36
FLOPS_PER_EXTRA_OBJECT = 9
FLOPS_SETUP = 385
FLOPS_TERMINAL = 130


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its
    version banner on stdout at init), so file descriptor 1 is pointed at stderr for the whole run and
    the result goes to the saved descriptor, unbuffered, the moment it exists."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_snapshot(name=WORKLOAD):
    from blackhole_8_b200 import abi
    return abi.SceneSnapshot.from_json(os.path.join(ROOT, "tests", "golden", name + ".json"))


def load_texture(name):
    import hashlib
    import cv2
    with open(os.path.join(ROOT, "resource", "MANIFEST.json")) as f:
        ent = json.load(f)[name]
    img = cv2.imread(os.path.join(ROOT, "resource", ent["file"]), cv2.IMREAD_COLOR)
    if img is None or hashlib.md5(img.tobytes()).hexdigest() != ent["md5_bgr"]:
        raise RuntimeError("texture %s does not match resource/MANIFEST.json" % name)
    return np.ascontiguousarray(img)


def frame_sequence(which, n_frames=240):
    """Consecutive frames as the reference's frame loop produces them (blackhole_solution_test.cc:346-407),
    snapshotted from the reference's own classes by tools/make_flythrough.py:
      "cfg1_spin"        configs[1]: fixed camera, the disc spins RotateZ(pi/180) after every frame (:407)
      "cfg3_flythrough"  configs[3]: 'w' MoveX(+10) for frames 0-119, then 'L' RotateZ(pi/180) + 'd'
                         MoveY(+10), plus the disc spin."""
    from blackhole_8_b200 import abi
    with open(os.path.join(ROOT, "tests", "golden", "states", which + ".json")) as f:
        fly = json.load(f)
    out = []
    for fr in fly["frames"][:n_frames]:
        d = json.loads(json.dumps(fly["base"]))
        d["camera"].update(fr["camera"])
        for o, v in zip(d["objects"], fr["v"]):
            o["v"] = v
        out.append(abi.SceneSnapshot.from_dict(d))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._nv = self._h = None
        try:  # NVML is opened on the caller's thread, so the first sample exists before the timed region starts
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def sample(self):
        nv, h = self._nv, self._h
        if nv is None:
            return
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = get(h)
            for bit, nm in names.items():
                if bits & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            time.sleep(0.002)

    def stop(self):
        self.sample()  # one more while the GPU is still warm: a region of a few ms is never left without samples
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_cpu(snap_name, frames, threads, frame=0):
    """The reference's CPU renderer for `frames` frames of the workload on `threads` host threads.
    Prefers oracle/_ref/ref_render (the reference's own classes, compiled in the build container
    from /root/reference; kind 'reference'); else the plain-C port (kind 'port')."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_render")
    tex = os.path.join(ROOT, "build", "textures")
    snap = load_snapshot(snap_name)
    rays = snap.width * snap.height
    if os.path.exists(ref):
        if not os.path.isdir(tex) or not os.listdir(tex):
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "decode_textures.py"), tex], check=True,
                           stdout=subprocess.DEVNULL)
        cfg = int(snap.meta.get("cfg", 1))
        out = subprocess.run([ref, "--cfg", str(cfg), "--width", str(snap.width), "--height", str(snap.height),
                              "--threads", str(threads), "--repeat", str(frames), "--frame", str(frame), "--texdir", tex],
                             check=True, capture_output=True, text=True).stdout
        run = json.loads(out)["run"]
        return {"kind": "reference", "ms_per_frame": run["mean_ms"], "best_ms": run["best_ms"],
                "rays": rays, "steps": run["steps"], "cores": threads}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O  # the oracle: only ever used as the measured CPU baseline / checker
    t0 = time.perf_counter()
    steps = 0
    for _ in range(frames):
        steps = O.render(snap, threads=threads)["result"].steps
    ms = (time.perf_counter() - t0) * 1e3 / frames
    return {"kind": "port", "ms_per_frame": ms, "best_ms": ms, "rays": rays, "steps": int(steps), "cores": threads}


def measure_sink(r, my_frames, W, H, world=1, rank=0, frames=120, quality=95):
    """SURVEY 8f-2, reported beside the headline: frames/s of the GPU frame sink (bh8_sink_render: trace the
    frame, draw the reference's HUD text into it and JPEG-encode it on the device with this repo's own kernels
    (csrc/bh8_jpeg.cuh), only the bitstream comes to the host) -- what the
    reference's `out_capture.write(frame)` becomes -- next to OpenCV's own JPEG encoder on one host core,
    which is what cv::VideoWriter(MJPG) spends per frame.  One sink per rank / GPU on that rank's frames;
    whole-job rate = world * frames / (max over ranks of the wall time).  Not part of value / e2e."""
    import torch
    import torch.distributed as dist
    err, out = None, {}
    try:
        import cv2
        from blackhole_8_b200 import abi
        from blackhole_8_b200.renderer import VideoSink
        sink = VideoSink(r, None, W, H, fps=29, quality=quality)
        sink.hud(abi.reference_hud(my_frames[0].camera))  # blackhole_solution_test.cc:309-326, drawn before the encode
        for i in range(5):
            sink.render(my_frames[i % len(my_frames)])
        before = sink.stats()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(frames):
            sink.render(my_frames[i % len(my_frames)])
        dt = time.perf_counter() - t0
        st = sink.stats()
        jpg = sink.last_jpeg()
        # pipelined form (bh8_sink_submit): frame k+1 is traced while frame k is encoded / read back
        for i in range(5):
            sink.submit(my_frames[i % len(my_frames)])
        sink.flush()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(frames):
            sink.submit(my_frames[i % len(my_frames)])
        sink.flush()
        dt_pipe = time.perf_counter() - t0
        sink.close()
        n = st["frames"] - before["frames"]
        out = {"dt": dt, "dt_pipe": dt_pipe, "n": n, "jpeg_bytes": st["jpeg_bytes"] - before["jpeg_bytes"],
               "encode_ms": st["encode_ms"] - before["encode_ms"]}
        if rank == 0:
            frame = r.render(my_frames[(frames - 1) % len(my_frames)], pixel_format=abi.PIXEL_BGR8)["pixels"][0]
            frame = np.ascontiguousarray(frame)
            for text, x, y in abi.reference_hud(my_frames[0].camera):
                cv2.putText(frame, text, (x, y), cv2.FONT_HERSHEY_PLAIN, 1, (0, 255, 0), 1)
            dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
            mse = float(np.mean((dec.astype(np.float64) - frame.astype(np.float64)) ** 2))
            n_cpu = 8
            t0 = time.perf_counter()
            for _ in range(n_cpu):
                ok, enc = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, quality])
            cpu_dt = time.perf_counter() - t0
            ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
            mse_cpu = float(np.mean((ref.astype(np.float64) - frame.astype(np.float64)) ** 2))
            out.update(psnr=10.0 * np.log10(255.0 ** 2 / mse) if mse > 0 else 99.0,
                       cpu={"frames_per_s": n_cpu / cpu_dt, "cores": 1, "bytes_per_frame": len(enc),
                            "psnr_db": 10.0 * np.log10(255.0 ** 2 / mse_cpu) if mse_cpu > 0 else 99.0})
    except Exception as e:  # the sink is an extra: never fail the headline line over it
        err = "%s: %s" % (type(e).__name__, e)
    dt_max, dt_pipe_max = out.get("dt", 0.0), out.get("dt_pipe", 0.0)
    ok_all = 0.0 if err else 1.0
    if world > 1:  # every rank takes part, also after a failure
        t = torch.tensor([dt_max, -ok_all, dt_pipe_max], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_max, ok_all, dt_pipe_max = float(t[0].item()), -float(t[1].item()), float(t[2].item())
    if rank != 0:
        return None
    if err or ok_all < 1.0:
        return {"unavailable": err or "a rank failed"}
    n = out["n"]
    return {"what": "bh8_sink_render: render + HUD text + baseline JPEG encode (own kernels, 4:2:0, quality %d, restart "
                    "interval 2 MCUs) on the GPU, bitstream to host; one sink per GPU" % quality,
            "frames_per_s": world * n / dt_max, "Mrays_per_s": world * n * W * H / dt_max / 1e6,
            "pipelined_bh8_sink_submit": {"frames_per_s": world * n / dt_pipe_max,
                                          "Mrays_per_s": world * n * W * H / dt_pipe_max / 1e6,
                                          "what": "two frames in flight: frame k+1 traced while frame k is encoded"},
            "jpeg_bytes_per_frame": out["jpeg_bytes"] / n, "raw_bytes_per_frame": W * H * 3,
            "encode_device_ms_per_frame": out["encode_ms"] / n, "psnr_db": out["psnr"],
            "host_opencv_imencode": out["cpu"]}


def measure_script(r, seq, which, d_frame, flags=0):
    """SURVEY 8f-3, reported beside the headline (N = 1): the frame sequence drawn from a device-resident
    script (bh8_script_*: the reference's per-frame Move*/Rotate* calls replayed on the GPU, frame constants
    built there, nothing per frame from the host) next to the same frames drawn from host snapshots
    (bh8_render_device), both back to back on one stream, wall clock around a final sync."""
    try:
        from blackhole_8_b200 import abi
        from blackhole_8_b200.renderer import Script
        n = len(seq)
        disc = [o.kind for o in seq[0].objects].index(abi.KIND_ANNULUS)
        t0 = time.perf_counter()
        sc = Script(r, seq[0], abi.reference_script(which, n, disc), n, flags=flags)
        create_s = time.perf_counter() - t0
        cam, objs = sc.state(n - 1)
        same = (list(cam.pos) == list(seq[n - 1].camera.pos) and
                all([list(a) for a in o.v] == [list(a) for a in q.v] for o, q in zip(objs, seq[n - 1].objects)))
        out = {}
        for name, draw in (("script", lambda k: sc.render(k, d_frame)),
                           ("host_snapshots", lambda k: r.render_device(seq[k], d_frame, flags=flags))):
            for k in range(10):
                draw(k)
            r.sync()
            t0 = time.perf_counter()
            for k in range(n):
                draw(k)
            r.sync()
            out[name] = n / (time.perf_counter() - t0)
        # the same frames with ONE host launch: the range captured into a CUDA graph
        fb = seq[0].width * seq[0].height * 4
        big = r.frame_alloc(n * fb)
        sc.render_range(0, n, big, fb)  # capture + first run
        r.sync()
        t0 = time.perf_counter()
        sc.render_range(0, n, big, fb)
        host_s = time.perf_counter() - t0  # what the launch costs the host
        r.sync()
        out["graph"] = n / (time.perf_counter() - t0)
        r.frame_free(big)
        sc.close()
        return {"what": "bh8_script_render: %d frames of '%s' animated and set up on the GPU, no per-frame host data"
                        % (n, which), "frames_per_s": out["script"], "host_snapshot_frames_per_s": out["host_snapshots"],
                "one_graph_launch": {"frames_per_s": out["graph"], "host_ms_for_all_frames": host_s * 1e3},
                "create_ms": create_s * 1e3, "last_state_equals_reference": bool(same)}
    except Exception as e:  # an extra: never fail the headline line over it
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference_frames(first_frame, n, threads, cfg=1, width=1920, height=1080):
    """Frames first_frame .. first_frame+n-1 of the workload's sequence through the reference's own classes
    (oracle/_ref/ref_render replays the frame loop's disc spin up to the frame it is asked for); one process
    per frame, the time is the pixel loop's own (what the reference prints as 'Took N ms')."""
    ms, steps = [], 0
    for k in range(n):
        r = run_reference_cpu(WORKLOAD, 1, threads, frame=first_frame + k)
        ms.append(r["ms_per_frame"])
        steps = r["steps"]
        kind = r["kind"]
    return {"kind": kind, "ms": ms, "steps": steps, "rays": width * height}


def bench_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    warmup = max(args.warmup, 3)
    steps = max(1, min(args.steps, 300))  # each step is one whole frame (~0.1-0.6 s on the host cores)
    run_reference_frames(0, warmup, threads)
    r = run_reference_frames(warmup, steps, threads)
    ms = sum(r["ms"]) / steps
    mrays = r["rays"] / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "threads": threads,
        "note": "the reference's pixel loop is single-threaded as shipped; here its rows are spread over all host "
                "threads with OpenMP (bit-identical output), so this baseline is the reference at its best on this box",
        "frames_per_s": 1e3 / ms, "gsteps_per_s": r["steps"] / (ms * 1e-3) / 1e9,
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": r["kind"],
                         "sample": "%d whole 1920x1080 frames of the workload's sequence (frames %d..%d)"
                                   % (steps, warmup, warmup + steps - 1)},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def allreduce_max(vals, world):
    if world == 1:
        return [float(v) for v in vals]
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def share_buffer(r, nbytes, rank, world):
    """A frame buffer on GPU 0 every rank's kernels can store into: rank 0 allocates and exports it through
    CUDA IPC, the others map it (NVLink peer memory)."""
    if world == 1:
        return r.frame_alloc(nbytes)
    import torch.distributed as dist
    handle = [None]
    if rank == 0:
        buf = r.frame_alloc(nbytes)
        handle[0] = r.ipc_export(buf)
    dist.broadcast_object_list(handle, src=0)
    if rank != 0:
        buf = r.ipc_import(handle[0])
    return buf


def measure_stripes_8k(r, base, rank, world, flags, steps=3, full=False):
    """BASELINE configs[4]: ONE 7680x4320 frame at nstep 200, sharded WITHIN the frame: interleaved
    16-row stripes dealt round-robin to the ranks (sharding.stripe_rows_of), every rank's kernel
    storing its stripes straight into GPU 0's frame buffer (IPC mapping, NVLink).  Strong scaling:
    the work is fixed, time = max over ranks of the kernel time.  Rank 0 then renders the whole frame
    alone (the N = 1 time in the same run) and compares the gathered frame with it byte for byte."""
    import torch
    from blackhole_8_b200 import abi
    W, H, nstep, stripe = 7680, 4320, 200, 16
    snap = base.with_resolution(W, H)
    frame_bytes = W * H * 4
    buf = share_buffer(r, frame_bytes, rank, world)
    flush_bytes = 256 << 20
    flush = r.frame_alloc(flush_bytes)
    kw = dict(nstep=nstep, flags=flags, stripe_rows=stripe, shard_index=rank, shard_count=world)
    if world == 1:
        kw.update(stripe_rows=0, shard_index=0, shard_count=0)
    r.render_device(snap, buf, **dict(kw, flags=flags | abi.FLAG_STATS))
    r.sync()
    st = r.read_stats()
    cnt = [st.rays, st.steps, st.warps, st.update_slots, st.resolve_passes, st.exact_tests]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor(cnt, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        cnt = t.tolist()
    launches0 = r.launches
    r.render_device(snap, buf, **kw)
    r.sync()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.sample()
    sampler.start()
    barrier(world)
    torch.cuda.synchronize()
    kernel_ms = 0.0
    for i in range(steps):
        r.memset_d(flush, i & 0xFF, flush_bytes)
        r.timer_begin()
        r.render_device(snap, buf, **kw)
        kernel_ms += r.timer_end_ms()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = r.launches - launches0 - 1
    # end to end: all ranks render their stripes, then rank 0 reads the gathered frame back
    pinned = r.pinned((H, W, 4)) if rank == 0 else None
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        r.render_device(snap, buf, **kw)
        r.sync()
        barrier(world)
        if rank == 0:
            r.memcpy_d2h(pinned.array, buf)
    e2e_s = time.perf_counter() - t0
    kernel_ms, e2e_s = allreduce_max([kernel_ms, e2e_s], world)
    out = None
    if rank == 0:
        ms = kernel_ms / steps
        rec = {"what": "cfg4: one 7680x4320 frame, cfg1 scene, nstep 200; interleaved %d-row stripes over the ranks, "
                       "stores into GPU 0's IPC-mapped frame (NVLink)" % stripe,
               "n_gpus": world, "steps": steps, "ms": ms, "Mrays_per_s": W * H / (ms * 1e-3) / 1e6,
               "steps_per_ray": cnt[1] / cnt[0], "gsteps_per_s": cnt[1] / (ms * 1e-3) / 1e9,
               "schedule": {"update_slots_per_warp": cnt[3] / cnt[2], "slot_utilisation": cnt[1] / (32.0 * cnt[3]),
                            "resolve_passes_per_warp": cnt[4] / cnt[2], "exact_tests_per_ray": cnt[5] / cnt[0]},
               "e2e_ms": e2e_s * 1e3 / steps, "e2e_what": "stripes on all ranks, barrier, rank 0 copies the 133 MB frame "
                                                          "to pinned host memory",
               "gpu_launches": launches, "clocks": clocks}
        if world > 1:
            # the same frame on rank 0 alone: the N = 1 time of this run, and the picture the gather must equal
            alone = r.frame_alloc(frame_bytes)
            kw1 = dict(nstep=nstep, flags=flags)
            r.render_device(snap, alone, **kw1)
            r.sync()
            ms1 = 0.0
            for i in range(steps):
                r.memset_d(flush, i & 0xFF, flush_bytes)
                r.timer_begin()
                r.render_device(snap, alone, **kw1)
                ms1 += r.timer_end_ms()
            ms1 /= steps
            ref = np.empty((H, W, 4), np.uint8)
            r.memcpy_d2h(ref, alone)
            same = bool(np.array_equal(ref, pinned.array))
            bands = {}
            for y0, y1 in ((2152, 2168), (304, 320)):  # the hole's silhouette; a band of plain sky/disc
                bands["%d-%d" % (y0, y1)] = bool(np.array_equal(ref[y0:y1], pinned.array[y0:y1]))
            rec.update(n1_ms=ms1, speedup_vs_n1=ms1 / ms, gather_equals_single_gpu_frame=same, row_bands_equal=bands)
            r.frame_free(alone)
        out = rec
        pinned.free()
    barrier(world)
    r.frame_free(flush)
    if world == 1 or rank == 0:
        r.frame_free(buf)
    else:
        r.ipc_close(buf)
    return out


def kernel_rate(r, frames, d_frame, flags, steps, l2_flush, warmup=3):
    """Mrays/s of `steps` frames drawn from `frames` (cyclic) with CUDA events per frame, L2 flushed
    between frames, plus the device counters of exactly those frames (one untimed BH8_FLAG_STATS pass)."""
    from blackhole_8_b200 import abi
    flush, flush_bytes = l2_flush
    for i in range(steps):
        r.render_device(frames[i % len(frames)], d_frame, flags=flags | abi.FLAG_STATS)
    r.sync()
    st = r.read_stats()
    for i in range(warmup):
        r.render_device(frames[i % len(frames)], d_frame, flags=flags)
    r.sync()
    ms = 0.0
    for i in range(steps):
        r.memset_d(flush, i & 0xFF, flush_bytes)
        r.timer_begin()
        r.render_device(frames[i % len(frames)], d_frame, flags=flags)
        ms += r.timer_end_ms()
    ms /= steps
    rays = st.rays / steps
    out = {"ms_per_frame": ms, "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "frames": steps,
           "steps_per_ray": st.steps / st.rays, "gsteps_per_s": st.steps / steps / (ms * 1e-3) / 1e9,
           "class_mix": {n: st.class_count[k] / st.rays for k, n in enumerate(("background", "horizon", "disc", "object"))}}
    sched = st.schedule()
    if sched:
        out["schedule"] = sched
    return out


def measure_other_workloads(r, base, flags, l2_flush, W, H):
    """The other BASELINE configs at N = 1 (each parity-tested in tests/test_gpu_parity.py): kernel-only
    Mrays/s, geodesic steps per ray and the warp-schedule counters.  Not part of value / e2e."""
    out = {}
    d_frame = r.frame_alloc(W * H * 4)
    try:
        for name, build, steps in (
                ("cfg2_objects_1080p", lambda: [load_snapshot("cfg2_1920x1080")], 20),
                ("cfg3_flythrough_1080p", lambda: frame_sequence("cfg3_flythrough", 240), 240),
                ("cfg10_flat_1080p", lambda: [load_snapshot("cfg10_flat_800x450").with_resolution(W, H)], 20)):
            try:
                frames = build()
                r.set_textures(frames[0], load_texture)
                out[name] = kernel_rate(r, frames, d_frame, flags, steps, l2_flush)
            except Exception as e:
                out[name] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        r.set_textures(base, load_texture)
    finally:
        r.frame_free(d_frame)
    try:
        rec = measure_stripes_8k(r, base, 0, 1, flags)
        rec.pop("clocks", None)
        out["cfg4_8k_nstep200"] = rec
    except Exception as e:
        out["cfg4_8k_nstep200"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return out


def instruction_model():
    """Static instruction model of the render kernel (tools/instr_model.py -> profiles/*_instr_model.json,
    made from the SASS of the committed sources and, for the per-pass costs, the committed ncu capture).
    It is only used when it was made from exactly the kernel sources that are being run."""
    import glob
    import hashlib
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_instr_model.json")))
    if not files:
        return None
    import re
    src = os.path.join(ROOT, "blackhole_8_b200", "csrc")

    def digest(names):  # the sources as the compiler sees them: no comments, no blank lines (tools/instr_model.py)
        h = hashlib.sha256()
        for n in names:
            with open(os.path.join(src, n)) as f:
                text = f.read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            text = re.sub(r"//[^\n]*", "", text)
            h.update("\n".join(ln.strip() for ln in text.splitlines() if ln.strip()).encode())
        return h.hexdigest()

    model = None
    for path in reversed(files):  # the newest model made from the running sources, else the newest (marked stale)
        with open(path) as f:
            m = json.load(f)
        m["file"] = os.path.relpath(path, ROOT)
        m["matches_running_sources"] = digest(m.get("sources", [])) == m.get("sources_sha256")
        if model is None or m["matches_running_sources"]:
            model = m
        if m["matches_running_sources"]:
            break
    return model


def measure_e2e(args, r, my_frames, rank, world, H, W, flags):
    """The metric through the public host-buffer calls, with the reference's frame format (CV_8UC3 = BGR8,
    blackhole_solution_test.cc:117,214): bh8_submit / bh8_wait, two frames in flight per GPU, every frame's
    snapshot in and its pixels out to pinned host memory inside the timed region.  A pass renders enough
    frames to last >= 0.35 s whatever --steps says; three passes, the median counts.  Beside it the
    read-back ceiling of the box: the same copies into the same buffers without kernels."""
    from blackhole_8_b200 import abi
    fmt = abi.PIXEL_BGR8
    frame_bytes = H * W * 3
    pinned = [r.pinned((H, W, 3)) for _ in range(2)]
    n_seq = len(my_frames)

    def run(n):
        t0 = time.perf_counter()
        prev = None
        for i in range(n):
            tk = r.submit(my_frames[i % n_seq], pinned[i & 1].array, pixel_format=fmt, flags=flags)
            if prev is not None:
                r.wait(prev)
            prev = tk
        r.wait(prev)
        return time.perf_counter() - t0

    run(8)
    barrier(world)
    cal = run(40)
    n = max(args.steps, int(np.ceil(0.35 / (cal / 40))))
    n = int(allreduce_max([n], world)[0])
    passes = []
    for _ in range(3):
        barrier(world)
        dt = run(n)
        passes.append(allreduce_max([dt], world)[0])
    e2e_s = sorted(passes)[1]
    # the frame the last pass left in pinned memory must be the frame the device-resident path draws
    last = (n - 1) % n_seq
    d_chk = r.frame_alloc(frame_bytes)
    r.render_device(my_frames[last], d_chk, pixel_format=fmt, flags=flags)
    r.sync()
    chk = np.empty((H, W, 3), np.uint8)
    r.memcpy_d2h(chk, d_chk)
    r.frame_free(d_chk)
    frame_ok = bool(np.array_equal(chk, pinned[(n - 1) & 1].array)) and bool(chk.any())
    # synchronous form, one call per frame (reported beside the headline)
    out = {"pixels": pinned[0].array.reshape(1, H, W, 3)}
    ns = max(20, n // 8)
    for i in range(3):
        r.render(my_frames[i], pixel_format=fmt, out=out, flags=flags)
    barrier(world)
    t0 = time.perf_counter()
    for i in range(ns):
        r.render(my_frames[i % n_seq], pixel_format=fmt, out=out, flags=flags)
    sync_s = allreduce_max([time.perf_counter() - t0], world)[0]
    # read-back ceiling: the same copies, same buffers, no kernels; all ranks at once
    r.measure_d2h(pinned[0], pinned[1], frame_bytes, 8)
    barrier(world)
    d2h_s = allreduce_max([r.measure_d2h(pinned[0], pinned[1], frame_bytes, n)], world)[0]
    wc = None
    try:
        pw = [r.pinned((H, W, 3), write_combined=True) for _ in range(2)]
        r.measure_d2h(pw[0], pw[1], frame_bytes, 8)
        barrier(world)
        wc_s = allreduce_max([r.measure_d2h(pw[0], pw[1], frame_bytes, n)], world)[0]
        wc = world * n * frame_bytes / wc_s / 1e9
        for b in pw:
            b.free()
    except Exception:
        barrier(world)
        allreduce_max([0.0], world)
    ok_all = allreduce_max([0.0 if frame_ok else 1.0], world)[0] == 0.0
    rays = H * W
    rate = world * n * rays / e2e_s / 1e6
    ceiling = world * n * rays / d2h_s / 1e6
    return {"value": rate, "unit": "Mrays/s",
            "h2d_bytes_per_step": int(r.lib.bh8_launch_param_bytes()) * world, "d2h_bytes_per_step": frame_bytes * world,
            "pixel_format": "BGR8", "frames_per_pass": n * world, "frames_per_s": world * n / e2e_s,
            "passes_s": passes, "d2h_GBps": world * n * frame_bytes / e2e_s / 1e9, "frame_ok": ok_all,
            "what": "bh8_submit()/bh8_wait(): snapshot -> kernel parameters, BGR8 frame (the reference's CV_8UC3) read back "
                    "into pinned host memory, two frames in flight per GPU, each on its own stream; wall clock, max over "
                    "ranks, median of 3 passes of >= 0.35 s",
            "d2h_only": {"what": "the same read-backs into the same pinned buffers with no kernel in between, all ranks "
                                 "at once (bh8_measure_d2h): the host-side ceiling of this box for this frame format",
                         "GBps": world * n * frame_bytes / d2h_s / 1e9, "Mrays_per_s_ceiling": ceiling,
                         "e2e_over_ceiling": rate / ceiling, "write_combined_GBps": wc},
            "synchronous_bh8_render": {"value": world * ns * rays / sync_s / 1e6, "unit": "Mrays/s",
                                       "frames_per_s": world * ns / sync_s}}


def check_gather(r, ring, seq, args, rank, world, ring_slots, frame_bytes, H, W, flags):
    """N > 1: what the other ranks' kernels stored into GPU 0's frame ring over NVLink must be, byte for byte,
    what rank 0 draws for the same frame index.  Checks the frames of the last ring_slots timed steps
    (earlier ones were overwritten); every rank takes part in the barriers, rank 0 returns the verdict."""
    from blackhole_8_b200 import sharding
    r.sync()
    barrier(world)
    res = None
    if rank == 0:
        got = np.empty((H, W, 4), np.uint8)
        want = np.empty((H, W, 4), np.uint8)
        scratch = r.frame_alloc(frame_bytes)
        checked, bad = 0, []
        for i in range(max(0, args.steps - ring_slots), args.steps):
            for q in range(world):
                k = q + (args.warmup + i) * world  # the frame rank q drew at timed step i (sharding.frames_of)
                r.render_device(seq[k % len(seq)], scratch, flags=flags)
                r.sync()
                r.memcpy_d2h(want, scratch)
                r.memcpy_d2h(got, ring + sharding.ring_slot_offset(i, q, world, ring_slots, frame_bytes))
                checked += 1
                if not np.array_equal(got, want) or not want.any():
                    bad.append([i, q])
        r.frame_free(scratch)
        res = {"gather_ok": not bad, "frames_checked": checked, "mismatching_step_rank": bad[:8]}
    barrier(world)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batching", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the legs that are reported beside the headline (sink, script, other workloads, 8K stripes)")
    ap.add_argument("--kernel-only", action="store_true",
                    help="tuning runs: skip the e2e / sink / script / CPU legs (their keys are null; not a driver line)")
    ap.add_argument("--workload", default="cfg1_spin",
                    choices=["cfg1_spin", "cfg3_flythrough", "cfg1_static", "cfg2_static", "cfg10_flat", "cfg4_8k"],
                    help="frame sequence: configs[1] with the disc spinning frame to frame as in the "
                         "reference's loop (default), the configs[3] fly-through, or configs[1] frame 0 only")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        bench_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from blackhole_8_b200 import abi
    from blackhole_8_b200.build import build
    from blackhole_8_b200.renderer import Renderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU fallback")
    build()
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    numa_cpus = None
    if world > 1:
        from blackhole_8_b200.sharding import bind_to_gpu_numa
        if not os.environ.get("BH8_NO_NUMA_BIND"):
            numa_cpus = bind_to_gpu_numa(local_rank)  # before any pinned allocation (first touch)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    base = load_snapshot()
    H, W = base.height, base.width
    rays = H * W
    r = Renderer((local_rank,))
    r.set_textures(base, load_texture)
    flags = abi.FLAG_NO_BATCHING if args.no_batching else 0
    if args.workload == "cfg4_8k":
        rec = measure_stripes_8k(r, base, rank, world, flags, steps=max(3, min(args.steps, 20)))
        if rank == 0:
            emit(dict(rec, metric="Mrays/s", value=rec["Mrays_per_s"], unit="Mrays/s", scaling="strong",
                      ms_per_step=rec["ms"], higher_is_better=True, dtype="f64", data="synthetic"))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # Frames are whole units of work: step i of the job renders frames i*world .. i*world+world-1 of
    # the sequence and rank r owns frame i*world + r (sharding.frames_of) -- the same workload at every
    # N, no data-path collective.
    from blackhole_8_b200 import sharding
    if args.workload == "cfg1_static":
        seq = [base]
    elif args.workload == "cfg2_static":  # configs[2]: the hole among ordinary textured objects and a chess floor
        base = load_snapshot("cfg2_1920x1080")
        r.set_textures(base, load_texture)
        seq = [base]
    elif args.workload == "cfg10_flat":  # the ray_tracer_test.cc scene (linear tracer) at 1080p
        base = load_snapshot("cfg10_flat_800x450").with_resolution(W, H)
        r.set_textures(base, load_texture)
        seq = [base]
    else:
        seq = frame_sequence(args.workload, 240)
    my_frames = [seq[k % len(seq)] for k in
                 sharding.frames_of((args.steps + args.warmup + 3) * world, rank, world)]

    # ---- device-resident frame ring on GPU 0 (peer-mapped into the other ranks) ----------------
    frame_bytes = rays * 4
    ring_slots = 4
    ring = share_buffer(r, frame_bytes * ring_slots * world, rank, world)
    flush_bytes = 256 << 20
    local_flush = r.frame_alloc(flush_bytes)  # each rank flushes its own GPU's L2

    def slot_ptr(step):
        return ring + sharding.ring_slot_offset(step, rank, world, ring_slots, frame_bytes)

    # untimed pass with the device counters on: geodesic steps / class mix / warp schedule of exactly the
    # frames the timed region renders (they feed steps/s and the flop counts)
    for i in range(args.steps):
        r.render_device(my_frames[args.warmup + i], slot_ptr(i), flags=flags | abi.FLAG_STATS)
    r.sync()
    st_raw = r.read_stats()
    cnt = torch.tensor([st_raw.rays, st_raw.steps] + list(st_raw.class_count) +
                       [st_raw.warps, st_raw.update_slots, st_raw.resolve_passes, st_raw.exact_tests],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    cnt = [float(x) / args.steps for x in cnt.tolist()]  # per step, summed over ranks

    class st:
        rays, steps, class_count = cnt[0], cnt[1], cnt[2:6]
        warps, update_slots, resolve_passes, exact_tests = cnt[6:10]
    launches0 = r.launches

    for i in range(args.warmup):
        r.render_device(my_frames[i], slot_ptr(i), flags=flags)
    r.sync()

    sampler = ClockSampler(local_rank)
    sampler.sample()
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    kernel_ms = 0.0
    for i in range(args.steps):
        snap = my_frames[args.warmup + i]
        r.memset_d(local_flush, i & 0xFF, flush_bytes)  # L2 flush, outside the event pair
        r.timer_begin()
        r.render_device(snap, slot_ptr(i), flags=flags)
        kernel_ms += r.timer_end_ms()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = r.launches - launches0 - args.warmup

    kernel_ms = allreduce_max([kernel_ms], world)[0]
    ms_per_step = kernel_ms / args.steps
    total_rays = rays * world * args.steps
    value = total_rays / (kernel_ms * 1e-3) / 1e6
    if args.kernel_only:  # tuning aid, not the driver's line
        if rank == 0:
            emit({"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                  "ms_per_step": ms_per_step, "kernel_only": True, "clocks": clocks, "gpu_launches": launches,
                  "lib": os.environ.get("BH8_LIB_PATH", "default")})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    gather = check_gather(r, ring, seq, args, rank, world, ring_slots, frame_bytes, H, W, flags) if world > 1 else None

    e2e = measure_e2e(args, r, my_frames, rank, world, H, W, flags)
    e2e["host_placement"] = ("rank 0 bound to %d CPUs next to its GPU (NVML affinity), every rank likewise"
                             % len(numa_cpus)) if numa_cpus else "as launched"

    sink = script = others = stripes = None
    if not args.no_extras:
        sink = measure_sink(r, my_frames, W, H, world, rank)  # every rank: one sink per GPU
        if world == 1 and args.workload in ("cfg1_spin", "cfg3_flythrough"):
            script = measure_script(r, seq, args.workload, ring, flags)
        if world == 1:
            others = measure_other_workloads(r, base, flags, (local_flush, flush_bytes), W, H)
        else:
            try:
                stripes = measure_stripes_8k(r, base, rank, world, flags)
            except Exception as e:  # an extra: never fail the headline line over it (all ranks fail alike)
                stripes = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline ------------------------------------------------------------------------------------
    # FP64-pipe bound (no tensor cores, no HBM traffic to speak of).  peak = DFMA chain measured now.
    peak, _ = r.measure_fp64_peak()
    step64, step32 = r.measure_stepping(False), r.measure_stepping(True)
    sec = ms_per_step * 1e-3
    # (a) what the hardware executes: every update slot runs the update for all 32 lanes of the warp
    stepping_flops = FLOPS_EXECUTED_PER_UPDATE * 32.0 * st.update_slots / world  # per launch
    model = instruction_model()
    fixed_flops = instr_per_warp = None
    if model and model["matches_running_sources"]:
        m = model["per_warp"]
        w_launch = st.warps / world
        passes = st.resolve_passes / world
        slots = st.update_slots / world
        fixed_flops = 32.0 * m.get("fp64_lane_fraction", 1.0) * (m["fp64_flops_fixed_per_lane"] * w_launch +
                                                                 m["fp64_flops_per_pass_per_lane"] * passes)
        instr_per_warp = (m["instr_fixed"] * w_launch + m["instr_per_update_slot"] * slots +
                          m["instr_per_pass"] * passes) / w_launch
    executed = stepping_flops + (fixed_flops or 0.0)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    clock_hz = (clocks.get("sm_mhz") or 1965) * 1e6
    hw = {"stepping_frac": (st.steps / world / sec) / step64,
          "stepping_frac_what": "geodesic updates/s of the renderer / updates/s of the bare update chain (bh8_measure_stepping), "
                                "both measured in this run",
          "renderer_updates_per_s": st.steps / world / sec, "bare_chain_updates_per_s": step64,
          "update_slot_utilisation": st.steps / (32.0 * st.update_slots),
          "stepping_fp64_flops_per_launch": stepping_flops}
    if instr_per_warp:
        issue_s = instr_per_warp * (st.warps / world) / (sm_count * 4 * clock_hz)
        hw.update(instr_per_warp_model=instr_per_warp, issue_bound_ms=issue_s * 1e3, issue_bound_frac=issue_s / sec,
                  issue_bound_what="modelled warp instructions / (SMs x 4 issue slots x SM clock) over the measured kernel time",
                  fixed_fp64_flops_per_launch=fixed_flops,
                  model={"file": model["file"], "from_committed_capture": True, "capture_git": model.get("git"),
                         "ncu_check": model.get("ncu_check")})
    elif model:
        hw["model"] = {"file": model["file"], "stale": "kernel sources changed since the model was made; not used"}
    # (b) the algorithmic count of SURVEY 8(d), kept for continuity: NOT a bound on this kernel, which skips
    # the per-step sincos / 1/u / object tests behind conservative filters
    n_extra = max(0, base.scene.n_obj - 2)
    hits = st.rays - st.class_count[0]
    flops_alg = ((FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra) * st.steps + FLOPS_SETUP * st.rays +
                 FLOPS_TERMINAL * hits) / world
    roofline = {"bound": "fp64", "achieved": executed / sec / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                "frac": executed / sec / peak, "traffic": None,
                "achieved_what": "FP64 flops the kernel EXECUTES per launch (17 per lane-update x 32 lanes x update slots, "
                                 "counted by the device; + the static per-ray / per-pass FP64 count of the instruction "
                                 "model when it matches the sources) / kernel time: <= peak by construction",
                "peak_source": "DFMA chain measured in this run (bh8_measure_fp64_peak: two register sources + an "
                               "immediate, the pipe's own rate -- a DFMA with three register sources is bound by "
                               "operand reads, profiles/r02_fp64_operand_mix.txt); MEASURED_PEAKS.json has no FP64 "
                               "vector figure",
                "hw": hw,
                "algorithmic": {"flops_per_launch": flops_alg, "TFLOPs": flops_alg / sec / 1e12,
                                "over_peak": flops_alg / sec / peak,
                                "convention": "SURVEY 8(d) W_sm100: %d flops/geodesic step, 385/ray setup, 130/terminal hit; "
                                              "exceeds what the kernel executes, so it is not a bound"
                                              % (FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra)},
                "precision_study": {"fp64_updates_per_s": step64, "fp32_updates_per_s": step32,
                                    "fp32_over_fp64": step32 / step64,
                                    "note": "bare geodesic update chain, no hit logic; FP32 parity cost measured in "
                                            "tests/test_precision_study.py"}}
    cpu = None
    if not args.no_cpu_baseline:
        threads = host_threads()
        c = run_reference_frames(args.warmup, 3, threads)
        c1 = run_reference_cpu(WORKLOAD, 1, 1)
        ms = sum(c["ms"]) / len(c["ms"])
        cpu = {"value": c["rays"] / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "cores": threads,
               "kind": c["kind"], "sample": "3 whole 1920x1080 frames of the workload's sequence, all host threads (OpenMP rows)",
               "ms_per_frame": ms,
               "single_thread_as_shipped": {"value": c1["rays"] / (c1["ms_per_frame"] * 1e-3) / 1e6,
                                            "ms_per_frame": c1["ms_per_frame"], "cores": 1}}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload),
        "workload_stats": {"rays_per_step": rays * world, "steps_per_ray": st.steps / st.rays,
                           "geodesic_steps_per_step": st.steps,
                           "class_mix": {"background": st.class_count[0] / st.rays, "horizon": st.class_count[1] / st.rays,
                                         "disc": st.class_count[2] / st.rays, "object": st.class_count[3] / st.rays},
                           "schedule": {"update_slots_per_warp": st.update_slots / st.warps,
                                        "slot_utilisation": st.steps / (32.0 * st.update_slots),
                                        "resolve_passes_per_warp": st.resolve_passes / st.warps,
                                        "exact_tests_per_ray": st.exact_tests / st.rays},
                           "batched_resolve": not args.no_batching},
        "frames_per_s": world * 1e3 / ms_per_step,
        "gsteps_per_s": st.steps / (ms_per_step * 1e-3) / 1e9,
        "wall_s_timed_region": wall_s,
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if world > 1:
        line["gather"] = dict(gather, what="rank r renders frame i*N+r of the sequence at step i into GPU 0's IPC-mapped "
                                          "frame ring (NVLink peer stores, no collective); rank 0 re-draws the frames of the "
                                          "last ring slots and compares the ring byte for byte")
        line["gather_ok"] = gather["gather_ok"]
        line["stripes_8k"] = stripes
    if sink:
        line["sink"] = sink
    if script:
        line["script"] = script
    if others:
        line["other_workloads"] = others
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if world > 1 and not gather["gather_ok"]:
        raise SystemExit("gather check failed: %r" % (gather,))


if __name__ == "__main__":
    main()
