#!/usr/bin/env python
"""Aggregate an ncu `--page source --print-source cuda,sass --csv` export of distance_wq_kernel into kernel regions
(function boundaries are read from the current rvs_distance_wq.cuh). usage: python profiles/src_regions.py src.csv"""
import collections
import csv
import os
import re
import sys

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rpmplv2_b200", "csrc", "rvs_distance_wq.cuh")
marks = []
for n, line in enumerate(open(SRC), 1):
    m = re.match(r"__device__ __forceinline__ \w[\w<>]* (\w+)", line) or re.match(r"template <class R> __device__ __forceinline__ R (\w+)", line)
    if m:
        marks.append((n, m.group(1)))
    elif "__global__" in line:
        marks.append((n, "prologue"))
    elif "inflated link boxes" in line:
        marks.append((n, "refresh"))
    elif "for (int j = j0; j < jend" in line:
        marks.append((n, "scan"))
    elif "per-configuration epilogue" in line:
        marks.append((n, "epilogue"))


def region(f, line):
    if f != "rvs_distance_wq.cuh":
        return f
    name = "header"
    for n, m in marks:
        if line >= n:
            name = m
    return name


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
cur, hdr = None, None
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit() and len(r) > 8:
        k = region(cur, int(r[0]))
        inst[k] += num(r[7]); samp[k] += num(r[6]); thr[k] += num(r[8])
ti, ts = sum(inst.values()), sum(samp.values())
print(f"warp instructions {ti:.3g}, samples {ts:.0f}")
for k, v in inst.most_common():
    print(f"{k:30s} inst {100 * v / ti:5.1f}%  samples {100 * samp[k] / ts:5.1f}%  active threads/inst {thr[k] / max(v, 1):5.1f}")
