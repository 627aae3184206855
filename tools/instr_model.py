#!/usr/bin/env python3
"""Instruction model of bh8_render_kernel<1, false, 2> from one ncu capture and the SASS of the same build.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       ncu -i X.ncu-rep --page raw --csv > raw.csv
       python tools/instr_model.py src.csv raw.csv blackhole_8_b200/libbh8.so <warps> <update_slots_per_warp> \
              <resolve_passes_per_warp> profiles/rNN_instr_model.json

Every SASS instruction of the kernel is put into one of five phases by the inline chain nvdisasm -gi records
for it (innermost frame first):
    pass   inside lane_resolve / lane_exact          -- runs once per resolve pass of a warp
    rare   inside lane_update_rare                   -- filter (2), leases, events
    lean   inside lane_update otherwise, and the stepping loop's votes (trace_patch's loop, warp_decide)
                                                     -- runs once per update slot
    setup  inside lane_setup                         -- once per warp
    other  prologue, mailbox, store                  -- once per warp
With the executed counts of the capture this gives, per warp,
    instr = fixed + instr_per_update_slot x update_slots + instr_per_pass x resolve_passes
(fixed = setup + rare + other at the captured workload) and the same for executed FP64 flops (DFMA 2, DMUL /
DADD 1, per lane).  bench.py feeds the device-counted update slots and resolve passes of ITS run into the
model: the line's `issue_bound_frac` and the fixed part of the executed flops come from here, marked as taken
from a committed capture, and are used only while the kernel sources still hash to `sources_sha256` (comments
and blank lines do not count)."""
import collections
import csv
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from sass_static_table import function_ranges  # noqa: E402

SOURCES = ["bh8_ray.cuh", "bh8_kernel.cuh", "bh8_warp.cuh", "bh8_frame.h"]
TAG = "bh8_render_kernelILi1ELb0ELi2E"
FLOPS = {"DFMA": 2, "DMUL": 1, "DADD": 1}


def sources_digest(src_dir, names):
    """sha256 of the kernel sources as the compiler sees them: comments, blank lines and indentation removed, so
    that rewording a comment does not orphan the model (bench.py carries the same function)."""
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(src_dir, n)) as f:
            text = f.read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        h.update("\n".join(ln.strip() for ln in text.splitlines() if ln.strip()).encode())
    return h.hexdigest()


def chains(so):
    """[(address, opcode, [(file, line), ...innermost first])] of the kernel's instructions."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    on, cur, out = False, [], []
    fresh = True
    for line in txt.splitlines():
        if line.startswith("\t.section\t.text."):
            on = TAG in line
            continue
        if line.startswith("\t.section"):
            on = False
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if fresh:
                cur, fresh = [], False
            ref = (os.path.basename(m.group(1)), int(m.group(2)))
            if ref not in cur:
                cur.append(ref)
            if m.group(3):
                ref = (os.path.basename(m.group(3)), int(m.group(4)))
                if ref not in cur:
                    cur.append(ref)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            op = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip()).split()[0].split(".")[0]
            out.append((int(m.group(1), 16), op, list(cur)))
            fresh = True
    return out


def phase_of(chain, ranges):
    fns = set()
    for f, ln in chain:
        name = None
        for first, fn in ranges.get(f, []):
            if first <= ln:
                name = fn
        fns.add(name or f)
    if fns & {"lane_resolve", "lane_exact"}:
        return "pass"
    if "lane_update_rare" in fns:
        return "rare"
    if fns & {"lane_update", "lane_advance"}:
        return "lean"
    if "lane_setup" in fns:
        return "setup"
    if "trace_patch" in fns or "warp_decide" in fns or "bh8_warp.cuh" in fns or "sm_80_rt.hpp" in fns:
        return "lean"  # the votes and the loop around the updates
    return "other"


def main():
    src_csv, raw_csv, so = sys.argv[1:4]
    warps, slots, passes = (float(x) for x in sys.argv[4:7])
    out_path = sys.argv[7]
    rows = list(csv.reader(open(src_csv)))
    hdr, executed = None, {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or not r or r[0] != "" or len(r) < 8:
            continue
        extra = len(r) - len(hdr)
        try:
            executed[int(r[2], 16)] = int(r[hdr.index("Instructions Executed") + extra])
        except (ValueError, IndexError):
            continue
    addrs = sorted(executed)
    static = chains(so)
    if len(addrs) != len(static):
        sys.exit("capture has %d instructions, %s has %d: not the same build" % (len(addrs), so, len(static)))
    src = os.path.join(ROOT, "blackhole_8_b200", "csrc")
    ranges = {f: function_ranges(os.path.join(src, f)) for f in ("bh8_ray.cuh", "bh8_kernel.cuh", "bh8_warp.cuh")}
    instr, flops = collections.Counter(), collections.Counter()
    for a, (_, op, chain) in zip(addrs, static):
        ph = phase_of(chain, ranges)
        instr[ph] += executed[a]
        flops[ph] += executed[a] * FLOPS.get(op, 0)  # per lane: a warp instruction is 32 lane-operations at most
    per = {k: v / warps for k, v in instr.items()}
    fl = {k: v / warps for k, v in flops.items()}
    raw = list(csv.reader(open(raw_csv)))
    col = {h: k for k, h in enumerate(raw[0])}

    def metric(name):
        return float(raw[2][col[name]]) if name in col else None

    # executed FP64 flops at thread level: the per-cycle rates of the raw page x elapsed cycles
    thread_flops = None
    names = ["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % x for x in ("dfma", "dmul", "dadd")]
    cycles = metric("sm__cycles_elapsed.avg") or metric("smsp__cycles_elapsed.avg")
    if all(n in col for n in names) and cycles:
        thread_flops = (2 * metric(names[0]) + metric(names[1]) + metric(names[2])) * cycles
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    fixed = per.get("setup", 0) + per.get("rare", 0) + per.get("other", 0)
    model = {
        "what": "instruction model of bh8_render_kernel<1,false,2> (tools/instr_model.py), calibrated on configs[1] 1920x1080",
        "git": git, "sources": SOURCES, "sources_sha256": sources_digest(src, SOURCES),
        "captured": {"warps": warps, "update_slots_per_warp": slots, "resolve_passes_per_warp": passes,
                     "instr_per_warp_by_phase": per, "fp64_flops_per_lane_per_warp_by_phase": fl,
                     "instr_per_warp": sum(per.values())},
        "per_warp": {"instr_fixed": fixed, "instr_per_update_slot": per.get("lean", 0) / slots,
                     "instr_per_pass": per.get("pass", 0) / passes,
                     "fp64_flops_fixed_per_lane": fl.get("setup", 0) + fl.get("rare", 0) + fl.get("other", 0),
                     "fp64_flops_per_update_slot_per_lane": fl.get("lean", 0) / slots,
                     "fp64_flops_per_pass_per_lane": fl.get("pass", 0) / passes,
                     # lanes that really executed an FP64 operation / 32 (predication, divergence), from the capture
                     "fp64_lane_fraction": (thread_flops / (32.0 * sum(flops.values()))) if thread_flops else 1.0},
        "ncu_check": {"inst_executed": metric("smsp__inst_executed.sum"), "model_inst_executed": sum(instr.values()),
                      "thread_fp64_flops_executed": thread_flops,
                      "model_fp64_flops_upper_bound_32_lanes": 32.0 * sum(flops.values()),
                      "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "fp64_pipe_active_pct": metric("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                      "kernel_us_under_ncu": metric("gpu__time_duration.sum")},
    }
    with open(out_path, "w") as f:
        json.dump(model, f, indent=1)
    print(json.dumps(model["captured"]["instr_per_warp_by_phase"], indent=1))
    print(json.dumps(model["per_warp"], indent=1))
    print(json.dumps(model["ncu_check"], indent=1))


if __name__ == "__main__":
    main()
