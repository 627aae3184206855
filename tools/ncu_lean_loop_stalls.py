#!/usr/bin/env python3
"""Per-instruction stall samples of the render kernel's stepping loop, from an ncu capture.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_lean_loop_stalls.py src.csv > profiles/rNN_lean_loop_stalls.txt
The loop is found as the REDUX vote and the two address ranges that execute as often as it does and hold the
update's FP64 instructions (the two unrolled copies of lane_update's plain step)."""
import csv
import sys

KEYS = ["stall_wait", "stall_math", "stall_short_sb", "stall_dispatch", "stall_no_inst", "stall_branch_resolving",
        "stall_not_selected", "stall_selected"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, seen = None, {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] == "" and r[2].startswith("0x"):
            a = int(r[2], 16)
            if a not in seen:
                seen[a] = dict(zip(hdr, r))
    addrs = sorted(seen)
    base = addrs[0]
    total = sum(int(seen[a]["# Samples"] or 0) for a in addrs)
    redux = [a for a in addrs if seen[a]["Source"].strip().startswith("REDUX")]
    rounds = max(int(seen[a]["Instructions Executed"]) for a in redux)
    hot = [a for a in addrs if int(seen[a]["Instructions Executed"] or 0) >= rounds]
    print("kernel: %d stall samples over %d SASS instructions; the stepping loop = the %d instructions executed at "
          "least once per round of votes (%d rounds)" % (total, len(addrs), len(hot), rounds))
    print("%-8s %-46s %7s  " % ("offset", "instruction", "samples") + " ".join(k[6:].rjust(9)[:9] for k in KEYS))
    loop = 0
    for a in hot:
        d = seen[a]
        n = int(d["# Samples"] or 0)
        loop += n
        print("0x%04x   %-46s %7d  " % (a - base, d["Source"].strip()[:46], n) +
              " ".join((d[k] or "0").rjust(9) for k in KEYS))
    print("stepping loop: %d samples = %.1f %% of the kernel's" % (loop, 100.0 * loop / total))
    for k in KEYS + ["stall_long_sb"]:
        print("  kernel total %-24s %6d  (%.1f %%)" % (k, sum(int(seen[a][k] or 0) for a in addrs),
                                                      100.0 * sum(int(seen[a][k] or 0) for a in addrs) / total))


if __name__ == "__main__":
    main()
