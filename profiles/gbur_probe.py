#!/usr/bin/env python
"""profiles/gbur_probe.py — host-call latency of the four engine calls of one DRGBT iteration (spatial_22dof, 6 objects) and of xarm6 M=51."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from rpmplv2_b200 import drgbt, engine, robots, synthetic
_horizon = synthetic.horizon_states

def med(fn, n=60):
    fn(); t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(t))

for name in ("spatial22", "xarm6"):
    if name == "spatial22":
        rb = robots.spatial(22); sc = scenes.dynamic_scene(rb, 5, 21, r_lo=1.0, r_hi=2.0, size=0.66)
        half = np.float32(0.5) * sc["dim"]; mn, mx, g = sc["pos"] - half, sc["pos"] + half, sc["is_ground"]
    else:
        rb, mn, mx, g, _ = scenes.xarm6_scene(51, "D2", 0.45)
    tol = np.full(len(g), 1000.0, np.float32)
    eng = engine.Engine(rb); eng.set_obstacles(mn, mx, g, tol)
    cand = synthetic.configs(rb, 300, seed=0xD6)
    q_cur = cand[int(np.argmax(eng.distance(cand)[0]))]
    sp = drgbt.default_spine_params()
    eng.store_reserve(4)
    for H in (1, 7, 52, 220):
        hor = _horizon(rb, q_cur, H, 11)
        slots = np.zeros(H, np.int32)
        t_root = med(lambda: eng.store_distance(q_cur[None], np.array([0], np.int32)))
        t_sp = med(lambda: eng.store_extend_spines(slots, hor, sp, False))
        ql, nl, _, st = eng.store_extend_spines(slots, hor, sp, False)
        qr = ql[np.arange(H), np.maximum(nl, 1) - 1]
        t_d = med(lambda: eng.distance(qr, robots.DIST_PROFILE))
        t_full = med(lambda: eng.distance(qr, robots.DIST_FULL))
        print(name, "H", H, "store_distance(1) %.1f us" % t_root, "store_extend_spines %.1f us" % t_sp, "distance PROFILE %.1f us" % t_d, "distance FULL %.1f us" % t_full, flush=True)
    eng.close()
