#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` dump by stall samples.
usage: python profiles/ncu_top.py src.csv [n_lines] [n_clips] [frames_per_clip]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
n_clips = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 498
cur, hdr, out = None, None, []


def num(d, k):
    try:
        return float(d.get(k, 0) or 0)
    except ValueError:
        return 0.0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        out.append((cur, int(r[0]), r[1].strip(), num(d, "Instructions Executed"), num(d, "# Samples"),
                    num(d, "L1 Wavefronts Shared"), num(d, "stall_long_sb"), num(d, "stall_barrier"), num(d, "stall_short_sb"),
                    num(d, "stall_wait"), num(d, "stall_mio"), num(d, "stall_not_selected"), num(d, "stall_math"),
                    num(d, "stall_no_inst"), num(d, "stall_dispatch")))
ti = sum(o[3] for o in out) or 1
ts = sum(o[4] for o in out) or 1
print(f"total warp-instr/frame {ti / n_clips / frames:.1f}  smem wavefronts/frame {sum(o[5] for o in out) / n_clips / frames:.1f}  samples {ts:.0f}")
tot = [sum(o[i] for o in out) for i in range(6, 15)]
names = ["long_sb", "barrier", "short_sb", "wait", "mio", "not_sel", "math", "no_inst", "dispatch"]
print("stall totals: " + "  ".join(f"{n} {100 * t / ts:.1f}%" for n, t in zip(names, tot)))
print("file:line            inst%  smp%  wf/fr |  long   bar short  wait   mio notsel  math  source")
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{o[0][:14]}:{o[1]:4d} {100 * o[3] / ti:5.1f} {100 * o[4] / ts:5.1f} {o[5] / n_clips / frames:6.1f} | "
          f"{o[6]:5.0f} {o[7]:5.0f} {o[8]:5.0f} {o[9]:5.0f} {o[10]:5.0f} {o[11]:5.0f} {o[12]:5.0f}  {o[2][:64]}")
