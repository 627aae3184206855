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


def ncu_summary():
    """Hardware view of the same kernel from the newest committed ncu capture (profiles/*_ncu_render_kernel_summary.txt):
    FP64 pipe utilisation, issue-slot utilisation, lanes per instruction, registers.  Static context for the
    live roofline numbers, never a substitute for them."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_render_kernel_summary.txt")))
    if not files:
        return None
    want = {"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_active_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_active_pct",
            "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
            "smsp__sass_average_branch_targets_threads_uniform.pct": "uniform_branch_targets_pct",
            "launch__registers_per_thread": "registers_per_thread",
            "sm__warps_active.avg.per_cycle_active": "warps_per_sm",
            "gpu__time_duration.sum": "kernel_us_under_ncu",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed": "dfma_per_cycle",
            "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed": "dmul_per_cycle",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed": "dadd_per_cycle",
            "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained": "fp64_lanes_per_cycle_peak"}
    out = {"source": os.path.relpath(files[-1], ROOT)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = 0.0
    for line in open(files[-1]):
        m = re.match(r"(\S+) \[(.*)\] = ([-0-9.e+]+)", line)
        if not m:
            continue
        if m.group(1) in want:
            out[want[m.group(1)]] = float(m.group(3))
        if m.group(1) in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += float(m.group(3)) * scale.get(m.group(2), 1.0)
    out["dram_bytes_per_launch"] = dram
    if "dfma_per_cycle" in out and out.get("fp64_lanes_per_cycle_peak"):
        flops = 2 * out.pop("dfma_per_cycle") + out.pop("dmul_per_cycle", 0.0) + out.pop("dadd_per_cycle", 0.0)
        out["executed_fp64_flop_frac_of_peak"] = flops / (2 * out.pop("fp64_lanes_per_cycle_peak"))
    return out


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
    emit(line)


def bench_8k_stripes(args, r, base, rank, world, flags):
    """BASELINE configs[4]: ONE 7680x4320 frame at nstep 200, sharded WITHIN the frame: interleaved
    16-row stripes dealt round-robin to the ranks (sharding.stripe_rows_of), every rank's kernel
    storing its stripes straight into GPU 0's frame buffer (IPC mapping, NVLink).  Strong scaling:
    the work per step is fixed, so value = rays of one frame / (max over ranks of the kernel time)."""
    import torch
    import torch.distributed as dist
    from blackhole_8_b200 import abi
    W, H, nstep, stripe = 7680, 4320, 200, 16
    snap = base.with_resolution(W, H)
    frame_bytes = W * H * 4
    handle = [None]
    if rank == 0:
        buf = r.frame_alloc(frame_bytes)
        handle[0] = r.ipc_export(buf) if world > 1 else None
    if world > 1:
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            buf = r.ipc_import(handle[0])
    flush_bytes = 256 << 20
    flush = r.frame_alloc(flush_bytes)
    kw = dict(nstep=nstep, flags=flags, stripe_rows=stripe, shard_index=rank, shard_count=world)
    steps = max(3, min(args.steps, 20))
    r.render_device(snap, buf, **dict(kw, flags=flags | abi.FLAG_STATS))
    r.sync()
    st = r.read_stats()
    cnt = torch.tensor([st.rays, st.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    launches0 = r.launches
    for _ in range(3):
        r.render_device(snap, buf, **kw)
    r.sync()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernel_ms = 0.0
    for i in range(steps):
        r.memset_d(flush, i & 0xFF, flush_bytes)
        r.timer_begin()
        r.render_device(snap, buf, **kw)
        kernel_ms += r.timer_end_ms()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = r.launches - launches0 - 3
    # end to end: all ranks render their stripes, then rank 0 reads the gathered frame back
    pinned = r.pinned((H, W, 4)) if rank == 0 else None
    e2e_steps = 3
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r.render_device(snap, buf, **kw)
        r.sync()
        if world > 1:
            dist.barrier()
        if rank == 0:
            r.memcpy_d2h(pinned.array, buf)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([kernel_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kernel_ms, e2e_s = float(t[0].item()), float(t[1].item())
    if rank != 0:
        return
    ms = kernel_ms / steps
    rays = W * H
    peak, _ = r.measure_fp64_peak()
    n_extra = max(0, base.scene.n_obj - 2)
    flops = (FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra) * cnt[1].item() + FLOPS_SETUP * cnt[0].item()
    emit(({
        "metric": "Mrays/s", "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
        "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: one 7680x4320 frame, cfg1 scene, nstep 200, interleaved %d-row stripes over the "
                               "ranks, stores into GPU 0's IPC-mapped frame" % stripe,
                   "rays_per_step": rays, "steps_per_ray": cnt[1].item() / cnt[0].item(),
                   "l2": "flushed between steps by a 256 MiB memset outside the per-step CUDA-event pair"},
        "frames_per_s": 1e3 / ms, "gsteps_per_s": cnt[1].item() / (ms * 1e-3) / 1e9, "clocks": clocks,
        "e2e": {"value": rays * e2e_steps / e2e_s / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": int(r.lib.bh8_launch_param_bytes()) * world, "d2h_bytes_per_step": frame_bytes,
                "what": "all ranks render their stripes into GPU 0, barrier, rank 0 copies the 133 MB frame to pinned host memory"},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "achieved": flops / world / (ms * 1e-3) / 1e12, "peak": peak / 1e12,
                     "unit": "TFLOP/s", "frac": flops / world / (ms * 1e-3) / peak, "traffic": None,
                     "note": "per GPU: algorithmic flops of its stripes / kernel time"},
        "cpu_baseline": None,
    }))


def measure_sink(r, my_frames, W, H, world=1, rank=0, frames=120, quality=95):
    """SURVEY 8f-2, reported beside the headline: frames/s of the GPU frame sink (bh8_sink_render: trace the
    frame, JPEG-encode it on the device with nvJPEG, only the bitstream comes to the host) -- what the
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
    return {"what": "bh8_sink_render: render + nvJPEG encode (4:2:0, quality %d) on the GPU, bitstream to host; "
                    "one sink per GPU" % quality,
            "frames_per_s": world * n / dt_max, "Mrays_per_s": world * n * W * H / dt_max / 1e6,
            "pipelined_bh8_sink_submit": {"frames_per_s": world * n / dt_pipe_max,
                                          "Mrays_per_s": world * n * W * H / dt_pipe_max / 1e6,
                                          "what": "two frames in flight: frame k+1 traced while frame k is encoded"},
            "jpeg_bytes_per_frame": out["jpeg_bytes"] / n, "raw_bytes_per_frame": W * H * 3,
            "nvjpeg_device_ms_per_frame": out["encode_ms"] / n, "psnr_db": out["psnr"],
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batching", action="store_true")
    ap.add_argument("--kernel-only", action="store_true",
                    help="tuning runs: skip the e2e / sink / script / CPU legs (their keys are null; not a driver line)")
    ap.add_argument("--workload", default="cfg1_spin",
                    choices=["cfg1_spin", "cfg3_flythrough", "cfg1_static", "cfg10_flat", "cfg4_8k"],
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
        bench_8k_stripes(args, r, base, rank, world, flags)
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

    # untimed pass with the device counters on: geodesic steps / class mix of exactly the frames the
    # timed region renders (they feed steps/s and the algorithmic flop count)
    for i in range(args.steps):
        r.render_device(my_frames[args.warmup + i], slot_ptr(i), flags=flags | abi.FLAG_STATS)
    r.sync()
    st_raw = r.read_stats()
    cnt = torch.tensor([st_raw.rays, st_raw.steps] + list(st_raw.class_count), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    cnt = [float(x) / args.steps for x in cnt.tolist()]  # per step, summed over ranks

    class _St:
        rays, steps, class_count = cnt[0], cnt[1], cnt[2:6]
    st = _St
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
    if args.kernel_only:  # tuning aid, not the driver's line
        if rank == 0:
            emit({"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                  "ms_per_step": ms_per_step, "kernel_only": True, "clocks": clocks, "gpu_launches": launches,
                  "lib": os.environ.get("BH8_LIB_PATH", "default")})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

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
    # three passes over the same e2e_steps frames, the median pass counts (a pass lasts tens of ms, so
    # a single one is at the mercy of one scheduling hiccup on the host)
    passes = []
    for _ in range(3):
        t0 = time.perf_counter()
        prev = None
        for i in range(e2e_steps):
            tk = r.submit(my_frames[args.warmup + i], pinned[i & 1].array, flags=flags)
            if prev is not None:
                r.wait(prev)
            prev = tk
        r.wait(prev)
        passes.append(time.perf_counter() - t0)
    e2e_s = sorted(passes)[1]
    if world > 1:
        t = torch.tensor([e2e_s, sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, sync_s = float(t[0].item()), float(t[1].item())
    e2e_value = rays * world * e2e_steps / e2e_s / 1e6
    e2e_sync_value = rays * world * e2e_steps / sync_s / 1e6
    frame_ok = bool(pinned[0].array[0, :, :, 3].min() == 255 and pinned[1].array[0, :, :, 3].min() == 255)

    sink = measure_sink(r, my_frames, W, H, world, rank)  # every rank: one sink per GPU
    script = None
    if world == 1 and args.workload in ("cfg1_spin", "cfg3_flythrough"):
        script = measure_script(r, seq, args.workload, ring, flags)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic FP64 flops / kernel time vs the DFMA peak measured now ----------
    n_extra = max(0, base.scene.n_obj - 2)
    hits = st.rays - st.class_count[0]
    # per step of the whole job (all ranks); one launch = one rank's frame = 1/world of it
    flops_frame = ((FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra) * st.steps + FLOPS_SETUP * st.rays +
                   FLOPS_TERMINAL * hits) / world
    peak, _ = r.measure_fp64_peak()
    achieved = flops_frame / (ms_per_step * 1e-3)
    roofline = {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": "DFMA chain measured in this run (bh8_measure_fp64_peak); MEASURED_PEAKS.json has "
                               "no FP64 vector figure",
                "algorithmic_flops_per_launch": flops_frame,
                "convention": "SURVEY 8(d) W_sm100: %d flops/geodesic step, 385/ray setup, 130/terminal hit"
                              % (FLOPS_PER_STEP_BASE + FLOPS_PER_EXTRA_OBJECT * n_extra)}

    # precision study: the bare update chain in FP64 (as shipped) and FP32, same launch shape
    step64, step32 = r.measure_stepping(False), r.measure_stepping(True)
    roofline["precision_study"] = {
        "fp64_updates_per_s": step64, "fp32_updates_per_s": step32, "fp32_over_fp64": step32 / step64,
        "renderer_updates_per_s": st.steps / world / (ms_per_step * 1e-3),
        "note": "bare geodesic update chain, no hit logic; FP32 parity cost measured in tests/test_precision_study.py"}
    ncu = ncu_summary()
    if ncu:
        roofline["traffic"] = ncu.pop("dram_bytes_per_launch", None)  # bytes/launch from the ncu capture
        roofline["pipe_utilisation"] = ncu
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
                               "; frame sequence '%s' (the reference's frame loop: disc spins pi/180 per frame)" % args.workload +
                               ("; rank r renders frame i*N+r of the sequence at step i into GPU 0's IPC-mapped frame "
                                "ring (NVLink peer stores, no collective)" if world > 1 else ""),
                   "rays_per_step": rays * world, "steps_per_ray": st.steps / st.rays,
                   "geodesic_steps_per_step": st.steps,
                   "class_mix": {"background": st.class_count[0] / st.rays, "horizon": st.class_count[1] / st.rays,
                                 "disc": st.class_count[2] / st.rays, "object": st.class_count[3] / st.rays},
                   "l2": "flushed between steps by a 256 MiB memset outside the per-step CUDA-event pair",
                   "batched_resolve": not args.no_batching},
        "frames_per_s": world * 1e3 / ms_per_step,
        "gsteps_per_s": st.steps / (ms_per_step * 1e-3) / 1e9,
        "wall_s_timed_region": wall_s,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(r.lib.bh8_launch_param_bytes()) * world,
                "d2h_bytes_per_step": frame_bytes * world, "steps": e2e_steps,
                "frames_per_s": world * e2e_steps / e2e_s, "frame_ok": frame_ok,
                "host_placement": ("rank 0 bound to %d CPUs next to its GPU (NVML affinity), every rank likewise"
                                   % len(numa_cpus)) if numa_cpus else "as launched",
                "what": "bh8_submit()/bh8_wait(): snapshot -> kernel parameters, RGBA8 frame read back into pinned "
                        "host memory, two frames in flight, each on its own stream (copy of frame k and the tail of its kernel "
                        "overlap kernel k+1), wall clock, median of 3 passes over the same frames",
                "passes_s": passes,
                "synchronous_bh8_render": {"value": e2e_sync_value, "unit": "Mrays/s",
                                           "frames_per_s": world * e2e_steps / sync_s}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if sink:
        line["sink"] = sink
    if script:
        line["script"] = script
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
