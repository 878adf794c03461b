#!/usr/bin/env python
"""SASS opcode mix of a kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.

usage: python profiles/ncu_opmix.py sass.csv [n_clips] [frames_per_clip] [split_addr_hex]
Prints executed warp-instructions per opcode (total and per frame).  With split_addr_hex (offset from the kernel's
first instruction) the table is printed twice: instructions below / at-or-above that offset (e.g. the R-warp branch
and the F-warp branch of fbank_ws_kernel).
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
n_clips = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 498
split = int(sys.argv[4], 16) if len(sys.argv) > 4 else None

rows = list(csv.reader(open(path)))
hdr = None
insts = []
for r in rows:
    if not r:
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if hdr and r[0].startswith("0x"):
        d = dict(zip(hdr, r))
        txt = r[1].strip()
        toks = txt.split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0] if toks else "?"
        base = op.split(".")[0]
        if base in ("LDS", "STS", "LDG", "STG"):     # keep the access width
            w = [x for x in op.split(".") if x in ("64", "128", "U8", "U16")]
            base = base + ("." + w[0] if w else "")
        insts.append((int(r[0], 16), base, float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0),
                      float(d.get("L1 Wavefronts Shared", 0) or 0), txt))
a0 = insts[0][0]


def table(sel, title):
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for a, op, n, s, wf, _ in sel:
        g = agg[op]
        g[0] += n; g[1] += s; g[2] += wf; g[3] += 1
    tot = sum(g[0] for g in agg.values()) or 1
    tots = sum(g[1] for g in agg.values()) or 1
    print(f"== {title}: {tot:.4e} warp-instr = {tot / n_clips / frames:.1f} per frame, {len(sel)} static instructions")
    print(f"{'opcode':12s} {'static':>6s} {'inst%':>7s} {'winst/frame':>12s} {'stall-smp%':>10s} {'smemWF/frame':>13s}")
    for op, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"{op:12s} {g[3]:6d} {100 * g[0] / tot:7.2f} {g[0] / n_clips / frames:12.2f} {100 * g[1] / tots:10.2f} {g[2] / n_clips / frames:13.1f}")


if split is None:
    table(insts, "kernel")
else:
    table([i for i in insts if i[0] - a0 < split], f"offset < {split:#x}")
    table([i for i in insts if i[0] - a0 >= split], f"offset >= {split:#x}")
