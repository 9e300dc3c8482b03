#!/usr/bin/env python
"""profiles/edge_latency_probe.py — host-call latency of the edge tests on xArm6 (scenario1: ground + 4 pillars): isValid(q1,q2) alone, the
self-collision edge test alone, and both in one call (rvs_check_edges), for planner-sized batches and two edge lengths."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from rpmplv2_b200 import engine, synthetic

def med(fn, n=200):
    fn(); t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(t))

rb, mn, mx, g, tol = scenes.xarm6_scenario1()
eng = engine.Engine(rb); eng.set_obstacles(mn, mx, g, tol)
for length in (0.1, 1.0):
    for E in (1, 10, 50, 500, 2000):
        q1, q2 = synthetic.edges(rb, E, length, seed=5)
        hits = int(eng.self_collision_edges(q1, q2)[0].sum())
        print(f"edge length {length} E {E:5d}  is_valid_edges {med(lambda: eng.is_valid_edges(q1, q2)):7.1f} us   self_collision_edges {med(lambda: eng.self_collision_edges(q1, q2)):7.1f} us   "
              f"check_edges {med(lambda: eng.check_edges(q1, q2)):7.1f} us   self hits {hits}", flush=True)
