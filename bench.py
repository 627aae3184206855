#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 geodesic renderer (driver contract).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU renderer, host cores

Workload (BASELINE.json configs[1]): single Schwarzschild black hole + thin accretion disc
(acc_disc.png) + textured background rectangle, 1920x1080, nstep 20 -- the scene snapshot the
reference's own classes produced (tests/golden/cfg1_1920x1080.json).  One "step" = one frame.
At N > 1 (one process per GPU, torchrun) every rank renders its own frame of the 240-frame camera
fly-through (configs[3]) per step -- weak scaling, no collective on the data path; frames are
gathered on GPU 0 by the kernels' own stores into its IPC-mapped frame ring (NVLink peer memory).

The JSON line: value = Mrays/s with everything resident on the GPU (CUDA events around each
frame's kernel, L2 flushed between frames outside the event pair); e2e = the same metric through
the public host-buffer call (bh8_render: snapshot in, frame read back into pinned host memory,
wall clock); roofline = algorithmic FP64 flops (SURVEY.md 8(d) convention W_sm100) / kernel time
against the DFMA peak measured in the same run; cpu_baseline = the reference's CPU renderer on
this box's host cores for the same frame.
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
FLOPS_PER_STEP_BASE = 57 + 8 + 10 + 36  # 2-object scene, rsqrt 8, rcp 10, sincos 36
FLOPS_PER_EXTRA_OBJECT = 9
FLOPS_SETUP = 385
FLOPS_TERMINAL = 130


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


def flythrough_snapshots(n_frames=240):
    """Frames of BASELINE configs[3]: the per-frame camera / disc states the reference's own classes
    produce when its frame loop is replayed (tools/make_flythrough.py -> tests/golden/states/cfg3_flythrough.json;
    'w' MoveX(+10) for frames 0-119, then 'L' RotateZ(pi/180) + 'd' MoveY(+10), disc RotateZ(pi/180)
    every frame: blackhole_solution_test.cc:346-407)."""
    from blackhole_8_b200 import abi
    with open(os.path.join(ROOT, "tests", "golden", "states", "cfg3_flythrough.json")) as f:
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

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                        nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    bits = get(h)
                    for bit, nm in names.items():
                        if bits & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
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


def run_reference_cpu(snap_name, frames, threads):
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
                              "--threads", str(threads), "--repeat", str(frames), "--texdir", tex],
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


def bench_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    for _ in range(min(args.warmup, 1)):
        run_reference_cpu(WORKLOAD, 1, threads)
    steps = max(1, min(args.steps, 20))  # each step is one whole frame (~0.3-1 s on the host cores)
    r = run_reference_cpu(WORKLOAD, steps, threads)
    mrays = r["rays"] / (r["ms_per_frame"] * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_frame"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg1: 1 Schwarzschild BH + accretion disc (acc_disc.png) + background rectangle, "
                               "1920x1080, nstep 20, one frame per step", "threads": threads,
                   "note": "reference pixel loop is single-threaded as shipped; rows are spread over all host "
                           "threads with OpenMP (bit-identical output)"},
        "frames_per_s": 1e3 / r["ms_per_frame"], "gsteps_per_s": r["steps"] / (r["ms_per_frame"] * 1e-3) / 1e9,
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": r["kind"],
                         "sample": "%d whole 1920x1080 frames of the workload" % steps},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batching", action="store_true")
    args = ap.parse_args()
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    base = load_snapshot()
    H, W = base.height, base.width
    rays = H * W
    r = Renderer((local_rank,))
    r.set_textures(base, load_texture)
    flags = abi.FLAG_NO_BATCHING if args.no_batching else 0

    # frames this rank renders: N == 1 -> the configs[1] frame every step; N > 1 -> fly-through frames
    if world > 1:
        from blackhole_8_b200 import sharding
        fly = flythrough_snapshots(240)
        # step i of the job renders frames i*world .. i*world+world-1; rank r owns frame i*world + r
        my_frames = [fly[k % 240] for k in sharding.frames_of((args.steps + args.warmup) * world, rank, world)]
    else:
        my_frames = [base] * (args.steps + args.warmup)

    # ---- device-resident frame ring on GPU 0 (peer-mapped into the other ranks) ----------------
    frame_bytes = rays * 4
    ring_slots = 4
    if world > 1:
        handle = [None]
        if rank == 0:
            ring = r.frame_alloc(frame_bytes * ring_slots * world)
            handle[0] = r.ipc_export(ring)
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ring = r.ipc_import(handle[0])
    else:
        ring = r.frame_alloc(frame_bytes * ring_slots)
    flush_bytes = 256 << 20
    local_flush = r.frame_alloc(flush_bytes)  # each rank flushes its own GPU's L2

    def slot_ptr(step):
        from blackhole_8_b200 import sharding
        return ring + sharding.ring_slot_offset(step, rank, world, ring_slots, frame_bytes)

    # one untimed launch with counters: steps / class mix of this rank's first frame
    r.render_device(my_frames[0], slot_ptr(0), flags=flags | abi.FLAG_STATS)
    r.sync()
    st = r.read_stats()
    launches0 = r.launches

    for i in range(args.warmup):
        r.render_device(my_frames[i], slot_ptr(i), flags=flags)
    r.sync()

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    kernel_ms = 0.0
    total_steps_geo = 0
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

    if world > 1:
        t = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = float(t.item())
    ms_per_step = kernel_ms / args.steps
    total_rays = rays * world * args.steps
    value = total_rays / (kernel_ms * 1e-3) / 1e6

    # ---- end to end through the public host-buffer calls --------------------------------------------
    # (a) bh8_render: one synchronous call per frame.  (b) bh8_submit / bh8_wait: the streaming form
    # of the same call, two frames in flight, so the read-back of frame k overlaps the kernel of
    # frame k+1.  Both copy every frame's snapshot in and its RGBA8 pixels out to pinned host memory
    # inside the timed region; (b) is the headline e2e number, (a) is reported beside it.
    pinned = [r.pinned((1, H, W, 4)) for _ in range(2)]
    e2e_steps = min(args.steps, 200)
    out = {"pixels": pinned[0].array}
    for i in range(3):
        r.render(my_frames[i], out=out, flags=flags)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r.render(my_frames[args.warmup + i], out=out, flags=flags)
    sync_s = time.perf_counter() - t0

    for i in range(3):
        r.wait(r.submit(my_frames[i], pinned[i & 1].array, flags=flags))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    prev = None
    for i in range(e2e_steps):
        tk = r.submit(my_frames[args.warmup + i], pinned[i & 1].array, flags=flags)
        if prev is not None:
            r.wait(prev)
        prev = tk
    r.wait(prev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, sync_s = float(t[0].item()), float(t[1].item())
    e2e_value = rays * world * e2e_steps / e2e_s / 1e6
    e2e_sync_value = rays * world * e2e_steps / sync_s / 1e6
    frame_ok = bool(pinned[0].array[0, :, :, 3].min() == 255 and pinned[1].array[0, :, :, 3].min() == 255)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic FP64 flops / kernel time vs the DFMA peak measured now ----------
    n_extra = max(0, base.scene.n_obj - 2)
    hits = int(st.rays - st.class_count[0])
    flops_frame = (FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra) * st.steps + FLOPS_SETUP * st.rays + \
        FLOPS_TERMINAL * hits
    peak, _ = r.measure_fp64_peak()
    achieved = flops_frame / (ms_per_step * 1e-3)
    roofline = {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": "DFMA chain measured in this run (bh8_measure_fp64_peak); MEASURED_PEAKS.json has "
                               "no FP64 vector figure",
                "algorithmic_flops_per_launch": flops_frame,
                "convention": "SURVEY 8(d) W_sm100: %d flops/geodesic step, 385/ray setup, 130/terminal hit"
                              % (FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra)}

    cpu = None
    if not args.no_cpu_baseline:
        threads = host_threads()
        c = run_reference_cpu(WORKLOAD, 3, threads)
        c1 = run_reference_cpu(WORKLOAD, 1, 1)
        cpu = {"value": c["rays"] / (c["ms_per_frame"] * 1e-3) / 1e6, "unit": "Mrays/s", "cores": c["cores"],
               "kind": c["kind"], "sample": "3 whole 1920x1080 frames of the workload, all host threads (OpenMP rows)",
               "ms_per_frame": c["ms_per_frame"],
               "single_thread_as_shipped": {"value": c1["rays"] / (c1["ms_per_frame"] * 1e-3) / 1e6,
                                            "ms_per_frame": c1["ms_per_frame"], "cores": 1}}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg1: 1 Schwarzschild BH + accretion disc (acc_disc.png) + background rectangle, "
                               "1920x1080, nstep 20, one frame per step" +
                               ("; N>1: every rank renders its own fly-through frame (cfg3) per step into GPU 0's "
                                "IPC-mapped frame ring" if world > 1 else ""),
                   "rays_per_step": rays * world, "steps_per_ray": st.steps / st.rays,
                   "class_mix": {"background": st.class_count[0] / st.rays, "horizon": st.class_count[1] / st.rays,
                                 "disc": st.class_count[2] / st.rays, "object": st.class_count[3] / st.rays},
                   "l2": "flushed between steps by a 256 MiB memset outside the per-step CUDA-event pair",
                   "batched_resolve": not args.no_batching},
        "frames_per_s": world * 1e3 / ms_per_step,
        "gsteps_per_s": st.steps * world / (ms_per_step * 1e-3) / 1e9,
        "wall_s_timed_region": wall_s,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 8192 * world,
                "d2h_bytes_per_step": frame_bytes * world, "steps": e2e_steps,
                "frames_per_s": world * e2e_steps / e2e_s, "frame_ok": frame_ok,
                "what": "bh8_submit()/bh8_wait(): snapshot -> kernel parameters, RGBA8 frame read back into pinned "
                        "host memory, two frames in flight (copy of frame k overlaps kernel of frame k+1), wall clock",
                "synchronous_bh8_render": {"value": e2e_sync_value, "unit": "Mrays/s",
                                           "frames_per_s": world * e2e_steps / sync_s}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
