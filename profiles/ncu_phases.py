#!/usr/bin/env python
"""Phase-level breakdown of an `ncu --page source --csv --print-source cuda,sass` dump of
fbank_fast_kernel: the kernel source carries `// [phase: NAME]` markers; every source line is
charged to the last marker above it (other files are charged to their file name).

usage: python profiles/ncu_phases.py src.csv path/to/fbank_fast.cuh [n_clips] [frames_per_clip]
"""
import csv
import re
import sys

src_csv, kernel_src = sys.argv[1], sys.argv[2]
n_clips = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 498
marks = []
for i, line in enumerate(open(kernel_src), 1):
    m = re.search(r"\[phase:\s*([^\]]+)\]", line)
    if m:
        marks.append((i, m.group(1).strip()))


def phase_of(fname, line):
    if not kernel_src.endswith(fname):
        return "file:" + fname
    name = "preamble"
    for l, n in marks:
        if l <= line:
            name = n
    return name


def num(d, k):
    try:
        return float(d.get(k, 0) or 0)
    except ValueError:
        return 0.0


rows = list(csv.reader(open(src_csv)))
cur, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        a = agg.setdefault(phase_of(cur, int(r[0])), [0.0, 0.0, 0.0, 0.0])
        a[0] += num(d, "Instructions Executed")
        a[1] += num(d, "# Samples")
        a[2] += num(d, "L1 Wavefronts Shared")
        a[3] += num(d, "L1 Wavefronts Shared Excessive")
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instr {ti:.4e}; per frame {ti / n_clips / frames:.1f}")
print(f"{'phase':34s} {'inst%':>7s} {'stall-smp%':>10s} {'winst/frame':>12s} {'smemWF/frame':>13s} {'excess/frame':>13s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} {100 * a[0] / ti:7.2f} {100 * a[1] / ts:10.2f} {a[0] / n_clips / frames:12.1f} "
          f"{a[2] / n_clips / frames:13.1f} {a[3] / n_clips / frames:13.1f}")
