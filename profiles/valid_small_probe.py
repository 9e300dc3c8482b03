#!/usr/bin/env python
"""profiles/valid_small_probe.py — host-call latency of planner-sized validity batches against 1000 boxes (xArm6)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpmplv2_b200 import engine, robots, synthetic
rb = robots.xarm6()
eng = engine.Engine(rb)
eng.set_obstacles(*synthetic.sweep_obstacles(1000, "D1", r_min=0.3))
def med(fn, n=40):
    fn(); t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(t))
for E in (1, 64, 512, 3200):
    q = synthetic.configs(rb, E, seed=5); q1, q2 = synthetic.edges(rb, E, 0.1)
    print("E", E, "isValid %.1f us" % med(lambda: eng.is_valid(q)), "isValidEdges (x10 checks) %.1f us" % med(lambda: eng.is_valid_edges(q1, q2)), flush=True)
