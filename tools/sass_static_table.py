#!/usr/bin/env python3
"""Static SASS instruction count per source function / line of bh8_render_kernel<NN> (no GPU needed).

usage: python tools/sass_static_table.py libbh8.so [NN] [top_n]
Disassembles the built library with line info (nvdisasm -g), attributes every instruction of the
kernel to the source line nvcc recorded for it and sums per enclosing function of bh8_ray.cuh /
bh8_kernel.cuh.  Straight-line per-ray code (setup, shading, store) executes once per warp, so its
static count is its dynamic cost; loops (stepping, exact test) need the ncu source page
(tools/ncu_source_table.py) for trip counts."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def function_ranges(path):
    """[(first_line, name)] of the function definitions in a source file (crude: BH8_HD / __global__ lines)."""
    out = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"\s*(?:BH8_HD|__global__|__device__|static __device__|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
        if m and not line.strip().startswith("//"):
            out.append((n, m.group(1)))
    return out


def main():
    so = sys.argv[1]
    nn = sys.argv[2] if len(sys.argv) > 2 else "1"
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    tag = "bh8_render_kernelILi%sELb0ELi2E" % nn if not nn.startswith("-") else "bh8_render_kernelILin%sELb0ELi2E" % nn[1:]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    on, cur = False, ("?", 0)
    per_line = collections.Counter()
    mix = collections.defaultdict(collections.Counter)
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
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            per_line[cur] += 1
            mix[cur][m.group(1).split(".")[0]] += 1
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackhole_8_b200", "csrc")
    ranges = {f: function_ranges(os.path.join(src, f)) for f in ("bh8_ray.cuh", "bh8_kernel.cuh")}
    per_fn = collections.Counter()
    fn_mix = collections.defaultdict(collections.Counter)
    for (f, ln), c in per_line.items():
        name = f
        for first, fn in ranges.get(f, []):
            if first <= ln:
                name = f + ":" + fn
        per_fn[name] += c
        fn_mix[name].update(mix[(f, ln)])
    total = sum(per_line.values())
    print("static SASS instructions in %s: %d" % (tag, total))
    for name, c in per_fn.most_common():
        top = ", ".join("%s %d" % kv for kv in fn_mix[name].most_common(6))
        print("  %-44s %6d  (%s)" % (name, c, top))
    print("top lines:")
    for (f, ln), c in per_line.most_common(top_n):
        print("  %s:%-5d %5d" % (f, ln, c))


if __name__ == "__main__":
    main()
