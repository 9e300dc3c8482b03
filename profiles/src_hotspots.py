#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python profiles/src_hotspots.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try:
                return float(d.get(k, 0) or 0)
            except ValueError:
                return 0.0
        lines.append((cur_file, int(r[0]), r[1].strip()[:90], f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"),
                      f("stall_barrier"), f("stall_wait"), f("stall_short_sb"), f("stall_branch_resolving"), f("stall_long_sb")))
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s:.0f}  total warp instructions {tot_i:.3g}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'thr/inst':>8s} {'barr':>6s} {'wait':>6s} {'ssb':>6s} {'br':>6s} {'lsb':>6s}  source")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    act = l[5] / l[4] if l[4] else 0
    print(f"{l[0][:20]+':'+str(l[1]):28s} {100*l[3]/tot_s:6.2f} {100*l[4]/tot_i:6.2f} {act:8.1f} {l[6]:6.0f} {l[7]:6.0f} {l[8]:6.0f} {l[9]:6.0f} {l[10]:6.0f}  {l[2]}")
