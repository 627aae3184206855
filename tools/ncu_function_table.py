#!/usr/bin/env python3
"""Executed warp instructions per source function of bh8_render_kernel<NN>, from an ncu capture.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_function_table.py src.csv libbh8.so <warps_launched> [NN] [top_lines]
ncu lists an inlined instruction under every source file of its call chain, so summing its source
page per file counts such instructions more than once.  Here every SASS address counts once and is
attributed to the INNERMOST source line, taken from the line table of the same binary
(nvdisasm -g, joined by instruction index)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_static_table import function_ranges  # noqa: E402


def innermost_lines(so, tag):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    on, cur, out = False, ("?", 0), []
    for line in txt.splitlines():
        if line.startswith("\t.section\t.text."):
            on = tag in line
            continue
        if line.startswith("\t.section"):
            on = False
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    src_csv, so, warps = sys.argv[1], sys.argv[2], float(sys.argv[3])
    nn = sys.argv[4] if len(sys.argv) > 4 else "1"
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    tag = "bh8_render_kernelILi%sELb0ELi2E" % nn
    rows = list(csv.reader(open(src_csv)))
    hdr, executed, samples = None, {}, {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or not r or r[0] != "" or len(r) < 8:
            continue
        extra = len(r) - len(hdr)
        try:
            a = int(r[2], 16)
            executed[a] = int(r[hdr.index("Instructions Executed") + extra])
            samples[a] = int(r[hdr.index("# Samples") + extra])
        except (ValueError, IndexError):
            continue
    addrs = sorted(executed)
    static = innermost_lines(so, tag)
    if len(addrs) != len(static):
        sys.exit("capture has %d instructions, %s has %d in %s: not the same build" % (len(addrs), so, len(static), tag))
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackhole_8_b200", "csrc")
    ranges = {f: function_ranges(os.path.join(src, f)) for f in ("bh8_ray.cuh", "bh8_kernel.cuh")}
    by_fn, by_line = collections.Counter(), collections.Counter()
    smp_fn = collections.Counter()
    pipes = collections.defaultdict(collections.Counter)
    for a, (_, (f, ln), text) in zip(addrs, static):
        name = f
        for first, fn in ranges.get(f, []):
            if first <= ln:
                name = f + ":" + fn
        by_fn[name] += executed[a]
        smp_fn[name] += samples[a]
        by_line[(f, ln)] += executed[a]
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]
        pipes[name][op] += executed[a]
    total, nsmp = sum(executed.values()), max(1, sum(samples.values()))
    print("warp instructions per launched warp: %.1f (%d SASS instructions, each counted once)" % (total / warps, len(addrs)))
    for name, c in by_fn.most_common():
        if c == 0:
            continue
        topops = ", ".join("%s %.0f" % (k, v / warps) for k, v in pipes[name].most_common(4))
        print("  %-36s %8.1f instr/warp %5.1f%%  samples %5.1f%%  (%s)" %
              (name, c / warps, 100.0 * c / total, 100.0 * smp_fn[name] / nsmp, topops))
    print("top lines:")
    for (f, ln), c in by_line.most_common(top):
        print("  %s:%-5d %8.1f" % (f, ln, c / warps))


if __name__ == "__main__":
    main()
