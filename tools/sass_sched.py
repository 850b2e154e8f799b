#!/usr/bin/env python
"""Static schedule of a kernel's SASS: decodes the control words (stall count, yield, scoreboard
set / wait) from `cuobjdump -sass` and prints, per opcode class, the instruction count and the
sum of the encoded stall cycles -- the issue time of ONE warp if no scoreboard wait ever blocks.
Usage: tools/sass_sched.py <file.sass> [--dump START END]   (run here, no GPU needed)"""
import re
import sys
from collections import Counter

ins = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
hi = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")


def parse(path):
    out = []
    lines = open(path).read().splitlines()
    i = 0
    while i < len(lines):
        m = ins.match(lines[i])
        if m and i + 1 < len(lines):
            h = hi.match(lines[i + 1])
            if h:
                w = int(h.group(1), 16)
                out.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=(w >> 41) & 0xf,
                                yld=(w >> 45) & 1, wbar=(w >> 46) & 7, rbar=(w >> 49) & 7, wait=(w >> 52) & 0x3f))
                i += 2
                continue
        i += 1
    return out


def opclass(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]


if __name__ == "__main__":
    prog = parse(sys.argv[1])
    if "--dump" in sys.argv:
        a, b = int(sys.argv[-2], 0), int(sys.argv[-1], 0)
        for p in prog:
            if a <= p["addr"] < b:
                print(f"{p['addr']:06x} st={p['stall']:2d} y={p['yld']} wb={p['wbar']} rb={p['rbar']} wt={p['wait']:02x}  {p['text']}")
        sys.exit(0)
    cnt, st = Counter(), Counter()
    for p in prog:
        c = opclass(p["text"])
        cnt[c] += 1
        st[c] += p["stall"]
    tot = sum(cnt.values())
    print(f"{tot} instructions, stall sum {sum(st.values())} ({sum(st.values()) / tot:.2f} per instruction)")
    for c, n in cnt.most_common(25):
        print(f"  {c:10s} {n:6d}  stall {st[c]:7d}  ({st[c] / n:.2f})")


def stage_report(prog, mark="MUFU"):
    """Stall sums between consecutive marker instructions (one marker per unrolled stage)."""
    idx = [i for i, p in enumerate(prog) if p["text"].startswith(mark)]
    out = []
    for a, b in zip(idx[:-1], idx[1:]):
        out.append((b - a, sum(p["stall"] for p in prog[a:b])))
    return out
