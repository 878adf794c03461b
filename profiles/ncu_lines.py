#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.

usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python profiles/ncu_lines.py src.csv [top_n]
Prints, per source line: warp instructions executed, stall samples, shared wavefronts.
"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "Function Name", "File Name") and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, 0) or 0)
            except ValueError:
                return 0.0
        out.append((cur_file, int(r[0]), r[1].strip()[:90], num("Instructions Executed"), num("# Samples"),
                    num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive"), num("stall_long_sb"), num("stall_short_sb"), num("stall_barrier"), num("stall_wait"), num("stall_mio")))
tot_i = sum(o[3] for o in out) or 1
tot_s = sum(o[4] for o in out) or 1
print(f"total warp-instr {tot_i:.3e}  samples {tot_s:.0f}")
print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s} {'smemWF':>10s} {'excess':>9s} {'long':>6s} {'short':>6s} {'bar':>5s} {'wait':>5s} {'mio':>5s}  source")
for o in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{o[0]+':'+str(o[1]):28s} {100*o[3]/tot_i:6.2f} {100*o[4]/tot_s:6.2f} {o[5]:10.3e} {o[6]:9.2e} {o[7]:6.0f} {o[8]:6.0f} {o[9]:5.0f} {o[10]:5.0f} {o[11]:5.0f}  {o[2]}")
