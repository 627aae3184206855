#!/usr/bin/env python3
"""Length of the stepping loop's fast path in the render kernel's SASS (no GPU needed).

usage: python tools/sass_lean_path.py libbh8.so [NN] [-v]
Finds the shortest instruction cycle through the stepping loop's first warp vote in
bh8_render_kernel<NN> that avoids everything only the slow paths do (shared/local memory traffic, S2R, FP32 MUFU) -- the path of a warp whose lanes all
take plain steps -- and prints its length and its instruction mix."""
import collections
import re
import subprocess
import sys

SLOW = ("STS", "LDS", "S2R", "S2UR", "MUFU.SIN", "MUFU.COS", "MUFU.RCP", "LDL", "STL", "CALL", "F2F", "LDG", "STG")


def load(so, nn):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = ("bh8_render_kernelILi%sELb0ELi2E" % nn) in line
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if on and m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def main():
    so = sys.argv[1]
    nn = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else "1"
    verbose = "-v" in sys.argv
    ins = load(so, nn)
    index = {a: k for k, (a, _) in enumerate(ins)}
    votes = [k for k, (_, t) in enumerate(ins) if "VOTE" in t or "REDUX" in t]
    head = votes[0]  # the shortest cycle through the stepping loop's vote is one lean loop iteration

    def succ(k):
        t = ins[k][1]
        if "BRA.DIV" in t:  # taken only by a diverged warp
            return [k + 1]
        m = re.search(r"\bBRA(?:\.U)?(?:\.ANY)?\b.*?(0x[0-9a-f]+)$", t)
        if m:
            tgt = index.get(int(m.group(1), 16))
            out = [tgt] if tgt is not None else []
            # conditional: a guard predicate (@P0 BRA) or a uniform-predicate operand (BRA.U !UP0, target)
            if t.startswith("@") or re.search(r"BRA\.U(?:\.ANY)?\s+!?UP\d", t):
                out.append(k + 1)
            return out
        if t.startswith("EXIT") or t.startswith("RET"):
            return []
        return [k + 1]

    # Every simple cycle through the vote that stays clear of slow-path instructions; the plain-step
    # path is the one that tests the step index against the plain range (an unsigned ISETP on span).
    cycles = []

    def walk(k, path, seen):
        if len(cycles) >= 64 or len(path) > 400:
            return
        for n in succ(k):
            if n is None or n >= len(ins):
                continue
            if n == head:
                cycles.append(list(path))
                continue
            body = ins[n][1].split(None, 1)[-1] if ins[n][1].startswith("@") else ins[n][1]
            if n in seen or any(body.startswith(x) for x in SLOW):
                continue
            seen.add(n)
            path.append(n)
            walk(n, path, seen)
            path.pop()
            seen.discard(n)

    sys.setrecursionlimit(10000)
    for head in votes[:4]:  # the stepping loop's vote is the first one that lies on a cycle of plain steps
        walk(head, [head], {head})
        if cycles:
            break
    assert cycles, "no cycle through the first votes"

    def is_plain(path):
        n_range = sum(1 for k in path if re.match(r"(@!?U?P\d\s+)?ISETP\.(LT|GE)\.U32\.(OR|AND)", ins[k][1])
                      and "0x" not in ins[k][1])
        n_upd = sum(1 for k in path if "MUFU.RSQ64H" in ins[k][1])
        return n_upd > 0 and n_range >= n_upd

    plain = [c for c in cycles if is_plain(c)] or cycles
    path = min(plain, key=len)
    print("cycles found: %d (lengths %s), plain-step candidates: %d" %
          (len(cycles), sorted(len(c) for c in cycles)[:8], len(plain)))
    ops = collections.Counter()
    for k in path:
        t = ins[k][1]
        t = t.split(None, 1)[-1] if t.startswith("@") else t
        ops[t.split()[0].split(".")[0]] += 1
    n_upd = sum(1 for k in path if "MUFU.RSQ64H" in ins[k][1])
    print("%s <%s>: vote at 0x%x, %d instructions per lean loop iteration, %d updates -> %.1f per update"
          % (so, nn, ins[head][0], len(path), n_upd, len(path) / max(1, n_upd)))
    print("  mix:", dict(ops.most_common()))
    if verbose:
        for k in path:
            print("    /*%04x*/ %s" % ins[k])


if __name__ == "__main__":
    main()
