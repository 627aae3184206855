#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_source_table.py src.csv <warps_launched> [top_n]
Prints executed warp instructions per launched warp and stall samples, by file and by line.

NOTE: ncu lists an inlined instruction under every source file of its call chain, so the per-file and
per-function sums here count such instructions more than once (the total comes out ~15 % high for this
kernel).  For instruction budgets use tools/ncu_function_table.py, which counts every SASS address once
and attributes it to its innermost source line; this tool stays for the per-line stall samples."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    warps = float(sys.argv[2])
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if not hdr or not cur or not r or r[0] == "":
            continue
        extra = len(r) - len(hdr)  # commas inside the source text split the cell
        try:
            line = int(r[0])
            ie = int(r[hdr.index("Instructions Executed") + extra])
            smp = int(r[hdr.index("# Samples") + extra])
        except (ValueError, IndexError):
            continue
        src = ",".join(r[1:2 + extra])[:90]
        a = agg.setdefault((cur, line), [0, 0, src])
        a[0] += ie
        a[1] += smp
    total = sum(a[0] for a in agg.values())
    samples = sum(a[1] for a in agg.values())
    print("warp instructions per launched warp: %.1f   stall samples: %d" % (total / warps, samples))
    byfile = {}
    for (f, _), a in agg.items():
        b = byfile.setdefault(f, [0, 0])
        b[0] += a[0]
        b[1] += a[1]
    for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  %-20s %9.1f instr/warp  %5.1f%% of samples" % (f, b[0] / warps, 100.0 * b[1] / max(1, samples)))
    # per enclosing function (same crude ranges as tools/sass_static_table.py)
    import os
    import re
    src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackhole_8_b200", "csrc")
    ranges = {}
    for fn in ("bh8_ray.cuh", "bh8_kernel.cuh"):
        out = []
        for n, text in enumerate(open(os.path.join(src_dir, fn)), 1):
            m = re.match(r"\s*(?:BH8_HD|__global__|__device__|static __device__|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
            if m and not text.strip().startswith("//"):
                out.append((n, m.group(1)))
        ranges[fn] = out
    byfn = {}
    for (f, line), a in agg.items():
        name = f
        for first, fn in ranges.get(f, []):
            if first <= line:
                name = f + ":" + fn
        b = byfn.setdefault(name, [0, 0])
        b[0] += a[0]
        b[1] += a[1]
    print("by function:")
    for f, b in sorted(byfn.items(), key=lambda kv: -kv[1][0])[:30]:
        print("  %-40s %9.1f instr/warp  %5.1f%% of samples" % (f, b[0] / warps, 100.0 * b[1] / max(1, samples)))
    print("by line:")
    for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print("%s:%-4d %9.1f %5.1f%%  %s" % (f, line, a[0] / warps, 100.0 * a[1] / max(1, samples), a[2]))


if __name__ == "__main__":
    main()
